/*
 * discrete_queue_multi.cpp -- the discrete_queue model (quokkasim_examples/src/bin/discrete_queue.rs:31-107) as ONE job
 * on every GPU of the box from a single host process, without Python: the C++ host mirror over qk_multi_* of the C ABI.
 * The library shards the replications over the devices, and the summary it prints is reduced over all of them inside
 * the library (one ncclAllGather of packed exact partials + one ncclAllReduce of the histograms).
 *
 *   g++ -std=c++17 -Iinclude examples/discrete_queue_multi.cpp -o discrete_queue_multi \
 *       -Lquokkasim_b200 -l:libquokka_b200.so -Wl,-rpath,$PWD/quokkasim_b200
 *   ./discrete_queue_multi [--replications N] [--events K] [--gpus G] [--steps S]
 *
 * BASELINE config C2: --replications 1048576 --events 10000 (the defaults).  Prints one line per step (device-timed,
 * maximum over the GPUs) and a key = value summary that is the same whatever the number of GPUs.
 */
#include <cstdlib>
#include <iostream>

#include "quokka_b200.hpp"

using namespace quokka;

int main(int argc, char** argv) {
  uint64_t n_replications = 1048576, n_events = 10000;
  int gpus = 0, steps = 1;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    if (a == "--replications" && i + 1 < argc) n_replications = std::strtoull(argv[++i], nullptr, 10);
    else if (a == "--events" && i + 1 < argc) n_events = std::strtoull(argv[++i], nullptr, 10);
    else if (a == "--gpus" && i + 1 < argc) gpus = std::atoi(argv[++i]);
    else if (a == "--steps" && i + 1 < argc) steps = std::atoi(argv[++i]);
  }
  try {
    DistributionFactory df(1234);
    df.next_seed = 0;
    auto source = ComponentModel::StringSource(
        DiscreteSource::new_().with_name("Source").with_process_time_distr(Distribution::Constant(3.)), Mailbox::new_());
    auto queue_1 = ComponentModel::StringStock(
        DiscreteStock::new_().with_name("Queue1").with_low_capacity(0).with_max_capacity(10).with_initial_resource({}),
        Mailbox::new_());
    auto process_1 = ComponentModel::StringProcess(
        DiscreteProcess::new_().with_name("Process1").with_process_time_distr(df.create(DistributionConfig::Triangular(1., 10., 6.))),
        Mailbox::new_());
    auto queue_2 = ComponentModel::StringStock(
        DiscreteStock::new_().with_name("Queue2").with_low_capacity(0).with_max_capacity(10).with_initial_resource({}),
        Mailbox::new_());
    auto process_2 = ComponentModel::StringProcess(
        DiscreteProcess::new_().with_name("Process2").with_process_time_distr(df.create(DistributionConfig::Triangular(1., 10., 6.))),
        Mailbox::new_());
    auto queue_3 = ComponentModel::StringStock(
        DiscreteStock::new_().with_name("Queue3").with_low_capacity(0).with_max_capacity(10).with_initial_resource({}),
        Mailbox::new_());
    auto sink = ComponentModel::StringSink(
        DiscreteSink::new_().with_name("Sink").with_process_time_distr(Distribution::Constant(1.)), Mailbox::new_());
    connect_components(source, queue_1);
    connect_components(queue_1, process_1);
    connect_components(process_1, queue_2);
    connect_components(queue_2, process_2);
    connect_components(process_2, queue_3);
    connect_components(queue_3, sink);

    auto sim_builder = BatchedSimInit::new_();
    for (auto* c : {&source, &queue_1, &process_1, &queue_2, &process_2, &queue_3, &sink})
      sim_builder = register_component(sim_builder, *c);

    BatchOptions opt;
    opt.n_replications = n_replications;
    opt.base_seed = 1234;
    if (gpus <= 0) {
      opt.all_devices = true;
    } else {
      for (int d = 0; d < gpus; ++d) opt.devices.push_back(d);
      if (gpus == 1) opt.device = 0;
    }
    qk_batch_summary total;
    std::vector<qk_comp_summary> st;
    for (int s = 0; s < steps; ++s) {
      auto simu = sim_builder.init(MonotonicTime::EPOCH, opt); /* Model::init of every replication is part of the step */
      simu.step_events(n_events);
      st = simu.stats(&total);
      const float ms = simu.last_step_ms();
      std::cout << "step " << s << ": gpus " << simu.n_shards() << " kernel_ms " << ms << " events_per_s "
                << (double)total.n_events / (ms * 1e-3) << "\n";
    }
    /* the summary: identical for 1, 2, 4, 8 GPUs (exact integer reduction) */
    std::cout << "replications = " << total.n_reps << "\nevents = " << total.n_events << "\nfaulted = " << total.n_faulted
              << "\nclock_ns_min = " << total.min_now_ns << "\nclock_ns_max = " << total.max_now_ns << "\n";
    const char* names[] = {"Source", "Queue1", "Process1", "Queue2", "Process2", "Queue3", "Sink"};
    for (size_t c = 0; c < st.size(); ++c) {
      uint64_t hsum = 0;
      for (int k = 0; k < QK_HIST_BINS; ++k) hsum += st[c].hist_time[k] * (uint64_t)(k + 1);
      std::cout << names[c] << ": sum_count = " << st[c].sum_count << " sum_time_ns_lo = " << st[c].sum_time_lo
                << " min_time = " << st[c].min_time << " max_time = " << st[c].max_time << " sumsq_time_ns2 = ["
                << st[c].sumsq_time_ns2[0] << "," << st[c].sumsq_time_ns2[1] << "," << st[c].sumsq_time_ns2[2]
                << "] sumsq_count = " << st[c].sumsq_count_lo << " hist_time_checksum = " << hsum << "\n";
    }
    return total.n_faulted == 0 ? 0 : 4;
  } catch (const Error& e) {
    std::cerr << "error (" << e.status << "): " << e.what() << "\n";
    return 1;
  }
}

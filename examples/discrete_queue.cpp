/*
 * discrete_queue.cpp -- quokkasim_examples/src/bin/discrete_queue.rs (reference lines 31-123) on the batched executor,
 * written against the C++ host mirror include/quokka_b200.hpp.  The model-building code reads like the Rust it
 * restates; what differs is `BatchedSimInit` where the reference has `SimInit`, and that one run advances N
 * independent replications (Philox streams seeded from (base_seed, replication)).
 *
 *   g++ -std=c++17 -Iinclude examples/discrete_queue.cpp -o discrete_queue \
 *       -Lquokkasim_b200 -l:libquokka_b200.so -Wl,-rpath,$PWD/quokkasim_b200
 *   ./discrete_queue [--replications N] [--out DIR] [--kat0]
 *
 * --kat0 drives the two process-time distributions from the pre-drawn tapes of the hand-derived KAT-0 trace
 * (P1 = [4, 5, ...], P2 = [2.5, ...]) and stops at 7.000000002 s, so that the CSV files can be compared
 * with tests/golden/kat0_*.csv.
 */
#include <cstdlib>
#include <iostream>

#include "quokka_b200.hpp"

using namespace quokka;

int main(int argc, char** argv) {
  uint64_t n_replications = 1024;
  std::string out_dir;
  bool kat0 = false;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    if (a == "--replications" && i + 1 < argc) n_replications = std::strtoull(argv[++i], nullptr, 10);
    else if (a == "--out" && i + 1 < argc) out_dir = argv[++i];
    else if (a == "--kat0") kat0 = true;
  }
  try {
    DistributionFactory df(1234);
    df.next_seed = 0;

    auto source = ComponentModel::StringSource(
        DiscreteSource::new_().with_name("Source").with_process_time_distr(Distribution::Constant(3.)), Mailbox::new_());
    auto queue_1 = ComponentModel::StringStock(
        DiscreteStock::new_().with_name("Queue1").with_low_capacity(0).with_max_capacity(10).with_initial_resource({}),
        Mailbox::new_());
    const Distribution d1 = df.create(DistributionConfig::Triangular(1., 10., 6.));
    auto process_1 = ComponentModel::StringProcess(DiscreteProcess::new_().with_name("Process1").with_process_time_distr(d1),
                                                   Mailbox::new_());
    auto queue_2 = ComponentModel::StringStock(
        DiscreteStock::new_().with_name("Queue2").with_low_capacity(0).with_max_capacity(10).with_initial_resource({}),
        Mailbox::new_());
    const Distribution d2 = df.create(DistributionConfig::Triangular(1., 10., 6.));
    auto process_par = ComponentModel::StringProcess(DiscreteProcess::new_().with_name("Process2").with_process_time_distr(d2),
                                                     Mailbox::new_());
    auto queue_3 = ComponentModel::StringStock(
        DiscreteStock::new_().with_name("Queue3").with_low_capacity(0).with_max_capacity(10).with_initial_resource({}),
        Mailbox::new_());
    auto sink = ComponentModel::StringSink(
        DiscreteSink::new_().with_name("Sink").with_process_time_distr(Distribution::Constant(1.)), Mailbox::new_());

    connect_components(source, queue_1);
    connect_components(queue_1, process_1);
    connect_components(process_1, queue_2);
    connect_components(queue_2, process_par);
    connect_components(process_par, queue_3);
    connect_components(queue_3, sink);
    try { /* discrete_queue.rs:14-20: an undefined pair is an Err, not a panic */
      connect_components(source, sink);
      return 2;
    } catch (const Error& e) {
      if (std::string(e.what()) != "No component connection defined from StringSource to StringSink (n=None)") return 3;
    }

    auto queue_logger = ComponentLogger::StringStockLogger("QueueLogger");
    auto process_logger = ComponentLogger::StringProcessLogger("ProcessLogger");
    connect_logger(queue_logger, queue_1);
    connect_logger(queue_logger, queue_2);
    connect_logger(queue_logger, queue_3);
    connect_logger(process_logger, source);
    connect_logger(process_logger, process_1);
    connect_logger(process_logger, process_par);
    connect_logger(process_logger, sink);

    auto sim_builder = BatchedSimInit::new_();
    sim_builder = register_component(sim_builder, source);
    sim_builder = register_component(sim_builder, queue_1);
    sim_builder = register_component(sim_builder, process_1);
    sim_builder = register_component(sim_builder, queue_2);
    sim_builder = register_component(sim_builder, process_par);
    sim_builder = register_component(sim_builder, queue_3);
    sim_builder = register_component(sim_builder, sink);

    BatchOptions opt;
    opt.n_replications = n_replications;
    opt.base_seed = 1234;
    opt.log_capacity = out_dir.empty() ? 0 : 4096;
    auto simu = sim_builder.init(MonotonicTime::EPOCH, opt);
    if (kat0) {
      std::vector<double> p1, p2;
      for (uint64_t r = 0; r < n_replications; ++r)
        for (int i = 0; i < 22; ++i) {
          p1.push_back(i == 0 ? 4.0 : (i == 1 ? 5.0 : 6.0));
          p2.push_back(i == 0 ? 2.5 : 6.0);
        }
      simu.set_tape(d1, p1, 22);
      simu.set_tape(d2, p2, 22);
      simu.step_until(7000000002ull);
    } else {
      simu.step_until(MonotonicTime::EPOCH + Duration::from_secs(200));
    }

    const qk_batch_summary s = simu.summary();
    std::cout << "replications " << s.n_reps << " events " << s.n_events << " faulted " << s.n_faulted << " clock_ns "
              << s.max_now_ns << "\n";
    const auto sunk = simu.comp_stats(sink, 0, n_replications);
    uint64_t items = 0;
    for (const auto& x : sunk) items += x.count;
    std::cout << "items through the sink, all replications: " << items << "\n";
    if (!out_dir.empty()) {
      std::cout << queue_logger.write_csv(out_dir, simu, n_replications - 1) << "\n";
      std::cout << process_logger.write_csv(out_dir, simu, n_replications - 1) << "\n";
    }
    return s.n_faulted == 0 ? 0 : 4;
  } catch (const Error& e) {
    std::cerr << "error (" << e.status << "): " << e.what() << "\n";
    return 1;
  }
}

/*
 * container_haulage.cpp -- a payload-true truck cycle on the batched executor, written against the C++ host mirror
 * (include/quokka_b200.hpp): a fixed fleet of Vector3Containers circulates
 *   ReadyToLoad -> ContainerLoadingProcess (ore from a Vector3Stock) -> Loaded -> Vector3ContainerParallelProcess (haul)
 *   -> AtDump -> ContainerUnloadingProcess (ore into a Vector3Stock) -> Emptied -> Vector3ContainerProcess (return)
 * with a breakdown mode on the loader and an environment that stops the shift for 200 s.  Same model, parameter for
 * parameter, as quokkasim_b200/workloads.py::container_haulage (the tests compare the two hosts).
 *
 *   g++ -std=c++17 -Iinclude examples/container_haulage.cpp -o container_haulage \
 *       -Lquokkasim_b200 -l:libquokka_b200.so -Wl,-rpath,$PWD/quokkasim_b200
 *   ./container_haulage [--replications N] [--events E] [--out DIR]     (--out: CSV files of the last replication)
 */
#include <cstdlib>
#include <iomanip>
#include <iostream>

#include "quokka_b200.hpp"

using namespace quokka;

int main(int argc, char** argv) {
  uint64_t n_replications = 256, n_events = 3000;
  std::string out_dir;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    if (a == "--replications" && i + 1 < argc) n_replications = std::strtoull(argv[++i], nullptr, 10);
    else if (a == "--events" && i + 1 < argc) n_events = std::strtoull(argv[++i], nullptr, 10);
    else if (a == "--out" && i + 1 < argc) out_dir = argv[++i];
  }
  try {
    const int n_trucks = 4;
    DistributionFactory df(777);
    df.next_seed = 0;

    auto ore = ComponentModel::Vector3Stock(VectorStock::new_().with_name("Ore").with_code("ORE").with_low_capacity_f(0.0)
                                                .with_max_capacity_f(1e9).with_initial_resource(Vector3(60000.0, 30000.0, 10000.0)),
                                            Mailbox::new_());
    std::vector<Vector3Container> trucks;
    for (int i = 0; i < n_trucks; ++i) {
      char id[16];
      std::snprintf(id, sizeof(id), "Truck_%02d", i);
      const double cap = 80.0 + 10.0 * (i % 4);
      trucks.push_back(i % 3 ? Vector3Container(id, cap) : Vector3Container(id, Vector3(1.0, 0.5, 0.25), cap));
    }
    auto ready = ComponentModel::Vector3ContainerStock(
        DiscreteStock::new_().with_name("ReadyToLoad").with_code("RTL").with_low_capacity(0).with_max_capacity(n_trucks + 2)
            .with_initial_resource(trucks), Mailbox::new_());
    const Distribution t_load = df.create(DistributionConfig::Triangular(20.0, 60.0, 30.0));
    const Distribution r_load = df.create(DistributionConfig::Uniform(0.8, 1.0));
    const Distribution brk_every = df.create(DistributionConfig::Uniform(300.0, 600.0));
    const Distribution brk_fix = df.create(DistributionConfig::Uniform(20.0, 60.0));
    auto load = ComponentModel::Vector3ContainerLoadProcess(
        ContainerLoadingProcess::new_().with_name("Loader").with_code("LD").with_max_process_count(2)
            .with_process_time_distr(t_load).with_process_capacity_ratio_distr(r_load)
            .with_delay_mode(DelayMode{"Breakdown", brk_every, brk_fix}), Mailbox::new_());
    auto loaded = ComponentModel::Vector3ContainerStock(
        DiscreteStock::new_().with_name("Loaded").with_code("LDD").with_max_capacity(n_trucks + 2), Mailbox::new_());
    auto haul = ComponentModel::Vector3ContainerParallelProcess(
        DiscreteParallelProcess::new_().with_name("Haul").with_code("HL")
            .with_process_time_distr(df.create(DistributionConfig::Triangular(100.0, 200.0, 140.0))), Mailbox::new_());
    auto at_dump = ComponentModel::Vector3ContainerStock(
        DiscreteStock::new_().with_name("AtDump").with_code("ATD").with_max_capacity(n_trucks + 2), Mailbox::new_());
    const Distribution t_dump = df.create(DistributionConfig::Uniform(15.0, 25.0));
    const Distribution r_dump = df.create(DistributionConfig::Uniform(0.85, 0.98));
    auto unload = ComponentModel::Vector3ContainerUnloadProcess(
        ContainerUnloadingProcess::new_().with_name("Dump").with_code("DMP").with_max_process_count(1)
            .with_process_time_distr(t_dump).with_process_quantity_ratio_distr(r_dump), Mailbox::new_());
    auto dumped = ComponentModel::Vector3Stock(
        VectorStock::new_().with_name("Dumped").with_code("DPD").with_low_capacity_f(0.0).with_max_capacity_f(1e9), Mailbox::new_());
    auto emptied = ComponentModel::Vector3ContainerStock(
        DiscreteStock::new_().with_name("Emptied").with_code("EMP").with_max_capacity(n_trucks + 2), Mailbox::new_());
    auto ret = ComponentModel::Vector3ContainerProcess(
        DiscreteProcess::new_().with_name("Return").with_code("RET")
            .with_process_time_distr(df.create(DistributionConfig::Triangular(30.0, 70.0, 45.0))), Mailbox::new_());
    auto env = ComponentModel::BasicEnvironment(BasicEnvironment::new_().with_name("Shift").with_code("ENV"), Mailbox::new_());

    connect_components(ore, load);
    connect_components(ready, load);
    connect_components(load, loaded);
    connect_components(loaded, haul);
    connect_components(haul, at_dump);
    connect_components(at_dump, unload);
    connect_components(unload, emptied);
    connect_components(unload, dumped);
    connect_components(emptied, ret);
    connect_components(ret, ready);
    connect_components(env, load);
    connect_components(env, unload);
    connect_components(env, ret);
    try { /* core.rs:1149-1172 has no (BasicEnvironment, Vector3ContainerParallelProcess) arm */
      connect_components(env, haul);
      return 2;
    } catch (const Error&) {
    }

    /* DiscreteStockLogger<Vector3Container> / DiscreteProcessLogger<Vector3Container> (core.rs:1292-1293, 1330-1334) */
    auto stock_logger = ComponentLogger::Vector3ContainerStockLogger("ContainerStockLogger");
    auto process_logger = ComponentLogger::Vector3ContainerProcessLogger("ContainerProcessLogger");
    for (auto* c : {&ready, &loaded, &at_dump, &emptied}) connect_logger(stock_logger, *c);
    for (auto* c : {&load, &haul, &unload, &ret}) connect_logger(process_logger, *c);

    auto sim_builder = BatchedSimInit::new_();
    for (auto* c : {&ore, &ready, &load, &loaded, &haul, &at_dump, &unload, &dumped, &emptied, &ret, &env})
      sim_builder = register_component(sim_builder, *c);
    BatchOptions opt;
    opt.n_replications = n_replications;
    opt.base_seed = 42;
    opt.log_capacity = out_dir.empty() ? 0 : 65536;
    auto simu = sim_builder.init(MonotonicTime::EPOCH, opt);
    create_scheduled_event(simu, Duration::from_secs(1500), ScheduledEventConfig::SetEnvironmentState(BasicEnvironmentState::Stopped), env);
    create_scheduled_event(simu, Duration::from_secs(1700), ScheduledEventConfig::SetEnvironmentState(BasicEnvironmentState::Normal), env);
    simu.step_events(n_events);

    const qk_batch_summary s = simu.summary();
    std::cout << "replications " << s.n_reps << " events " << s.n_events << " faulted " << s.n_faulted << " clock_ns "
              << s.max_now_ns << "\n";
    /* mass balance of the last replication: ore + dumped + carried by the fleet is constant */
    const uint64_t r = n_replications - 1;
    double total = simu.read_vector(ore, r).total() + simu.read_vector(dumped, r).total();
    size_t fleet = 0;
    for (auto* c : {&ready, &load, &loaded, &haul, &at_dump, &unload, &emptied, &ret})
      for (auto& k : simu.read_containers(*c, r)) {
        fleet += 1;
        if (k.second.has_resource) total += k.second.resource.total();
      }
    std::cout << std::setprecision(17) << "fleet " << fleet << " dumped " << simu.read_vector(dumped, r).total() << " total " << total
              << "\n";
    if (!out_dir.empty()) {
      std::cout << stock_logger.write_csv(out_dir, simu, r) << "\n";
      std::cout << process_logger.write_csv(out_dir, simu, r) << "\n";
    }
    return (s.n_faulted == 0 && fleet == (size_t)n_trucks && std::abs(total - 100003.5) < 1e-4) ? 0 : 4;
  } catch (const Error& e) {
    std::cerr << "error (" << e.status << "): " << e.what() << "\n";
    return 1;
  }
}

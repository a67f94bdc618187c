//! quokkasim_examples/src/bin/discrete_queue.rs handed to the B200 executor: the model-building part is the
//! reference's, with `SimInit` -> `BatchedSimInit` and the enum constructors taking the inert Mailbox.
use std::time::Duration;

use quokkasim_b200::*;

fn main() -> Result<(), Box<dyn std::error::Error>> {
    let mut df = DistributionFactory { base_seed: 1234, next_seed: 0 };

    let mut source = ComponentModel::StringSource_(
        DiscreteSource::new().with_name("Source").with_process_time_distr(Distribution::Constant(3.)),
        Mailbox::new(),
    );
    let queue = |name: &str| {
        ComponentModel::StringStock_(
            DiscreteStock::new().with_name(name).with_low_capacity(0).with_max_capacity(10).with_initial_resource(ItemDeque::default()),
            Mailbox::new(),
        )
    };
    let mut queue_1 = queue("Queue1");
    let mut process_1 = ComponentModel::StringProcess_(
        DiscreteProcess::new()
            .with_name("Process1")
            .with_process_time_distr(df.create(DistributionConfig::Triangular { min: 1., max: 10., mode: 6. })?),
        Mailbox::new(),
    );
    let mut queue_2 = queue("Queue2");
    let mut process_2 = ComponentModel::StringProcess_(
        DiscreteProcess::new()
            .with_name("Process2")
            .with_process_time_distr(df.create(DistributionConfig::Triangular { min: 1., max: 10., mode: 6. })?),
        Mailbox::new(),
    );
    let mut queue_3 = queue("Queue3");
    let mut sink = ComponentModel::StringSink_(
        DiscreteSink::new().with_name("Sink").with_process_time_distr(Distribution::Constant(1.)),
        Mailbox::new(),
    );

    connect_components!(&mut source, &mut queue_1)?;
    connect_components!(&mut queue_1, &mut process_1)?;
    connect_components!(&mut process_1, &mut queue_2)?;
    connect_components!(&mut queue_2, &mut process_2)?;
    connect_components!(&mut process_2, &mut queue_3)?;
    connect_components!(&mut queue_3, &mut sink)?;

    let mut queue_logger = ComponentLogger::StringStockLogger(DiscreteStockLogger::new("QueueLogger"));
    let mut process_logger = ComponentLogger::StringProcessLogger(DiscreteProcessLogger::new("ProcessLogger"));
    for q in [&mut queue_1, &mut queue_2, &mut queue_3] {
        connect_logger!(&mut queue_logger, q)?;
    }
    for p in [&mut source, &mut process_1, &mut process_2, &mut sink] {
        connect_logger!(&mut process_logger, p)?;
    }

    let mut sim_builder = BatchedSimInit::new();
    sim_builder = register_component!(sim_builder, source);
    sim_builder = register_component!(sim_builder, queue_1);
    sim_builder = register_component!(sim_builder, process_1);
    sim_builder = register_component!(sim_builder, queue_2);
    sim_builder = register_component!(sim_builder, process_2);
    sim_builder = register_component!(sim_builder, queue_3);
    sim_builder = register_component!(sim_builder, sink);

    // one million replications with seeds 1234 + (0 .. 2^20), newest 64 log records kept per replication
    let opt = BatchOptions { n_replications: 1 << 20, base_seed: 1234, log_capacity: 64, ..Default::default() };
    let mut simu = sim_builder.init(MonotonicTime::EPOCH, opt)?;
    simu.step_until(MonotonicTime::EPOCH + Duration::from_secs(200))?;

    let (status, now_ns, n_events) = simu.status(0)?;
    println!("replication 0: status {} now {} ns, {} events", status, now_ns, n_events);
    let sunk = simu.comp_stats(&sink, 0)?;
    println!("replication 0: sink finished {} items, busy {} ns", sunk.count, sunk.time_ns);
    for r in simu.read_log(0)?.iter().rev().take(5) {
        println!("{:>14} ns  {}  <- {}  kind {}", r.time_ns, simu.event_id(r.comp, r.seq), simu.event_id(r.src_comp, r.src_seq), r.kind);
    }
    Ok(())
}

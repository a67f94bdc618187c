"""Synthetic workloads of the named shapes (BASELINE.json `configs`): the reference's example topologies
rebuilt with the mirrored builder API.  Used by bench.py and by the tests (no oracle involvement here).
"""
from . import api as qk


class Model:
    """A built model: registered ComponentModels in registration order + scheduled env events."""

    def __init__(self):
        self.components = []
        self.ext = []  # (t_ns, env ComponentModel, state)
        self.random_dists = []  # Distribution objects that need tapes in tape mode
        self.by_name = {}

    def reg(self, cm):
        self.components.append(cm)
        self.by_name[cm.component.element_name] = cm
        return cm

    def sim(self, lib, n_reps, t0=0, base_seed=0, rep_offset=0, log_capacity=0, threads_per_block=0, stream=None, flags=0):
        si = qk.BatchedSimInit.new()
        for c in self.components:
            si = qk.register_component(si, c)
        sim, sched = si.init(t0, n_replications=n_reps, base_seed=base_seed, rep_offset=rep_offset,
                             log_capacity=log_capacity, threads_per_block=threads_per_block, stream=stream,
                             lib=lib, flags=flags)
        for t, env, state in self.ext:
            qk.create_scheduled_event(sched, t, qk.ScheduledEventConfig.SetEnvironmentState(state), env.get_address())
        return sim


def _log_all(m):
    sl = qk.ComponentLogger.StringStockLogger(qk.DiscreteStockLogger.new("StockLogger"))
    pl = qk.ComponentLogger.StringProcessLogger(qk.DiscreteProcessLogger.new("ProcessLogger"))
    el = qk.ComponentLogger.BasicEnvironmentLogger(qk.BasicEnvironmentLogger.new("EnvLogger"))
    for c in m.components:
        if c.variant == "StringStock":
            qk.connect_logger(sl, c)
        elif c.variant == "BasicEnvironment":
            qk.connect_logger(el, c)
        else:
            qk.connect_logger(pl, c)
    m.loggers = (sl, pl, el)


def discrete_queue(tri=(1.0, 10.0, 6.0), src_time=3.0, sink_time=1.0, cap=10):
    """quokkasim_examples/src/bin/discrete_queue.rs:31-107"""
    m = Model()
    df = qk.DistributionFactory(base_seed=1234, next_seed=0)
    source = m.reg(qk.ComponentModel.StringSource(
        qk.DiscreteSource.new().with_name("Source").with_process_time_distr(qk.Distribution.Constant(src_time)),
        qk.Mailbox.new()))
    q1 = m.reg(qk.ComponentModel.StringStock(
        qk.DiscreteStock.new().with_name("Queue1").with_low_capacity(0).with_max_capacity(cap)
        .with_initial_resource(qk.ItemDeque.default()), qk.Mailbox.new()))
    d1 = df.create(qk.DistributionConfig.Triangular(min=tri[0], max=tri[1], mode=tri[2]))
    p1 = m.reg(qk.ComponentModel.StringProcess(
        qk.DiscreteProcess.new().with_name("Process1").with_process_time_distr(d1), qk.Mailbox.new()))
    q2 = m.reg(qk.ComponentModel.StringStock(
        qk.DiscreteStock.new().with_name("Queue2").with_low_capacity(0).with_max_capacity(cap), qk.Mailbox.new()))
    d2 = df.create(qk.DistributionConfig.Triangular(min=tri[0], max=tri[1], mode=tri[2]))
    p2 = m.reg(qk.ComponentModel.StringProcess(
        qk.DiscreteProcess.new().with_name("Process2").with_process_time_distr(d2), qk.Mailbox.new()))
    q3 = m.reg(qk.ComponentModel.StringStock(
        qk.DiscreteStock.new().with_name("Queue3").with_low_capacity(0).with_max_capacity(cap), qk.Mailbox.new()))
    sink = m.reg(qk.ComponentModel.StringSink(
        qk.DiscreteSink.new().with_name("Sink").with_process_time_distr(qk.Distribution.Constant(sink_time)),
        qk.Mailbox.new()))
    for a, b in ((source, q1), (q1, p1), (p1, q2), (q2, p2), (p2, q3), (q3, sink)):
        qk.connect_components(a, b)
    m.random_dists = [d1, d2]
    _log_all(m)
    return m


def car_workshop():
    """quokkasim_examples/src/bin/car_workshop.rs: 3 hoists sharing both stocks, all Constant -> exact ties."""
    m = Model()
    arrivals = qk.ComponentModel.StringSource(
        qk.DiscreteSource.new().with_name("Arrivals").with_code("A")
        .with_item_factory(qk.StringItemFactory("Car", 0, 4))
        .with_process_time_distr(qk.Distribution.Constant(900.0)), qk.Mailbox.new())
    rts = qk.ComponentModel.StringStock(
        qk.DiscreteStock.new().with_name("Ready to Service").with_code("Q1").with_low_capacity(0).with_max_capacity(3),
        qk.Mailbox.new())
    hoists = [qk.ComponentModel.StringProcess(
        qk.DiscreteProcess.new().with_name("Car Hoist %d" % i).with_code("P%d" % i)
        .with_process_time_distr(qk.Distribution.Constant(1800.0)), qk.Mailbox.new()) for i in range(3)]
    rtd = qk.ComponentModel.StringStock(
        qk.DiscreteStock.new().with_name("Ready to Depart").with_code("Q2").with_low_capacity(0).with_max_capacity(3),
        qk.Mailbox.new())
    dep = qk.ComponentModel.StringSink(
        qk.DiscreteSink.new().with_name("Departures").with_code("D")
        .with_process_time_distr(qk.Distribution.Constant(1.0)), qk.Mailbox.new())
    qk.connect_components(arrivals, rts)
    for h in hoists:
        qk.connect_components(rts, h)
        qk.connect_components(h, rtd)
    qk.connect_components(rtd, dep)
    m.reg(arrivals)
    m.reg(rts)
    for h in hoists:
        m.reg(h)
    m.reg(rtd)
    m.reg(dep)
    _log_all(m)
    m.t0 = qk.MonotonicTime.try_from_date_time(2025, 7, 1, 8, 0, 0, 0)
    return m


def assembly_line(parallel=True, env_events=True, stream0=0):
    """quokkasim_examples/src/bin/assembly_line.rs: line with a DiscreteParallelProcess and an environment that
    stops the line at +3600 s and resumes it at +7200 s."""
    m = Model()
    df = qk.DistributionFactory(base_seed=12345, next_seed=stream0)
    source = qk.ComponentModel.StringSource(
        qk.DiscreteSource.new().with_name("Source").with_code("SRC")
        .with_process_time_distr(qk.Distribution.Constant(15.0 * 60)), qk.Mailbox.new())
    q1 = qk.ComponentModel.StringStock(
        qk.DiscreteStock.new().with_name("Queue1").with_code("Q1").with_low_capacity(0).with_max_capacity(10),
        qk.Mailbox.new())
    d1 = df.create(qk.DistributionConfig.Triangular(min=600.0, max=1500.0, mode=780.0))
    p1 = qk.ComponentModel.StringProcess(
        qk.DiscreteProcess.new().with_name("Process1").with_code("P1").with_process_time_distr(d1), qk.Mailbox.new())
    q2 = qk.ComponentModel.StringStock(
        qk.DiscreteStock.new().with_name("Queue2").with_code("Q2").with_low_capacity(0).with_max_capacity(10),
        qk.Mailbox.new())
    d2 = df.create(qk.DistributionConfig.Triangular(min=600.0, max=1500.0, mode=780.0))
    if parallel:
        p2 = qk.ComponentModel.StringParallelProcess(
            qk.DiscreteParallelProcess.new().with_name("Process2").with_code("P2").with_process_time_distr(d2),
            qk.Mailbox.new())
    else:
        p2 = qk.ComponentModel.StringProcess(
            qk.DiscreteProcess.new().with_name("Process2").with_code("P2").with_process_time_distr(d2),
            qk.Mailbox.new())
    q3 = qk.ComponentModel.StringStock(
        qk.DiscreteStock.new().with_name("Queue3").with_code("Q3").with_low_capacity(0).with_max_capacity(10),
        qk.Mailbox.new())
    sink = qk.ComponentModel.StringSink(
        qk.DiscreteSink.new().with_name("Sink").with_code("SNK")
        .with_process_time_distr(qk.Distribution.Constant(15.0 * 60)), qk.Mailbox.new())
    env = qk.ComponentModel.BasicEnvironment(
        qk.BasicEnvironment.new().with_name("EnvironmentController").with_code("ENV"), qk.Mailbox.new())
    for a, b in ((source, q1), (q1, p1), (p1, q2), (q2, p2), (p2, q3), (q3, sink)):
        qk.connect_components(a, b)
    for c in (source, p1, p2, sink):
        qk.connect_components(env, c)
    for c in (source, q1, p1, q2, p2, q3, sink, env):
        m.reg(c)
    m.t0 = qk.MonotonicTime.try_from_date_time(2025, 7, 1, 0, 0, 0, 0)
    if env_events:
        m.ext.append((m.t0 + qk.Duration.from_secs(3600), env, qk.BasicEnvironmentState.Stopped))
        m.ext.append((m.t0 + qk.Duration.from_secs(7200), env, qk.BasicEnvironmentState.Normal))
    m.random_dists = [d1, d2]
    _log_all(m)
    return m


def serial_line(n_proc=4, cap=10, breakdowns=True, exp=True, src_time=3.0, sink_time=1.0):
    """BASELINE config C4 in miniature: Source -> [Stock_i -> DiscreteProcess_i] x n -> Stock -> Sink, process time
    Tri(1,10,6); each process one DelayMode until_delay Exp(600 s) / until_fix Exp(30 s)."""
    m = Model()
    df = qk.DistributionFactory(base_seed=7, next_seed=0)
    stream = [1000]

    def exp_dist(mean):
        # the factory's Exponential arm skips next_seed += 1 (Q6); give each instance its own stream explicitly
        d = df.create(qk.DistributionConfig.Exponential(mean)) if exp else \
            df.create(qk.DistributionConfig.Uniform(0.5 * mean, 1.5 * mean))
        d.stream = stream[0]
        stream[0] += 1
        return d

    source = m.reg(qk.ComponentModel.StringSource(
        qk.DiscreteSource.new().with_name("Source").with_code("SRC")
        .with_process_time_distr(qk.Distribution.Constant(src_time)), qk.Mailbox.new()))
    prev = source
    for i in range(n_proc):
        q = m.reg(qk.ComponentModel.StringStock(
            qk.DiscreteStock.new().with_name("Stock%d" % i).with_code("S%d" % i).with_max_capacity(cap),
            qk.Mailbox.new()))
        d = df.create(qk.DistributionConfig.Triangular(min=1.0, max=10.0, mode=6.0))
        m.random_dists.append(d)
        proc = qk.DiscreteProcess.new().with_name("Process%d" % i).with_code("P%d" % i).with_process_time_distr(d)
        if breakdowns:
            ud, uf = exp_dist(600.0), exp_dist(30.0)
            m.random_dists += [ud, uf]
            proc = proc.with_delay_mode(qk.DelayModeChange.Add(qk.DelayMode("Breakdown", ud, uf)))
        p = m.reg(qk.ComponentModel.StringProcess(proc, qk.Mailbox.new()))
        qk.connect_components(prev, q)
        qk.connect_components(q, p)
        prev = p
    q = m.reg(qk.ComponentModel.StringStock(
        qk.DiscreteStock.new().with_name("Stock%d" % n_proc).with_code("S%d" % n_proc).with_max_capacity(cap),
        qk.Mailbox.new()))
    sink = m.reg(qk.ComponentModel.StringSink(
        qk.DiscreteSink.new().with_name("Sink").with_code("SNK")
        .with_process_time_distr(qk.Distribution.Constant(sink_time)), qk.Mailbox.new()))
    qk.connect_components(prev, q)
    qk.connect_components(q, sink)
    _log_all(m)
    return m


def haulage(n_src=2, per_queue=3, cap=5):
    """BASELINE config C5 in miniature: n Sources -> n shared load queues -> n*per_queue DiscreteProcesses
    (per_queue per load queue) -> n shared dump queues -> n Sinks. Shared stocks as car_workshop.rs:76-82."""
    m = Model()
    df = qk.DistributionFactory(base_seed=99, next_seed=0)
    procs_all = []
    for s in range(n_src):
        ds = df.create(qk.DistributionConfig.Uniform(1.0, 3.0))
        m.random_dists.append(ds)
        src = m.reg(qk.ComponentModel.StringSource(
            qk.DiscreteSource.new().with_name("Src%d" % s).with_code("SRC%d" % s)
            .with_item_factory(qk.StringItemFactory("Load%d" % s, 0, 5)).with_process_time_distr(ds),
            qk.Mailbox.new()))
        lq = m.reg(qk.ComponentModel.StringStock(
            qk.DiscreteStock.new().with_name("LoadQ%d" % s).with_code("LQ%d" % s).with_max_capacity(cap),
            qk.Mailbox.new()))
        qk.connect_components(src, lq)
        procs = []
        for k in range(per_queue):
            d = df.create(qk.DistributionConfig.Triangular(min=2.0, max=12.0, mode=5.0))
            m.random_dists.append(d)
            p = m.reg(qk.ComponentModel.StringProcess(
                qk.DiscreteProcess.new().with_name("Truck%d_%d" % (s, k)).with_code("T%d_%d" % (s, k))
                .with_process_time_distr(d), qk.Mailbox.new()))
            qk.connect_components(lq, p)
            procs.append(p)
        procs_all.append(procs)
    for s in range(n_src):
        dq = m.reg(qk.ComponentModel.StringStock(
            qk.DiscreteStock.new().with_name("DumpQ%d" % s).with_code("DQ%d" % s).with_max_capacity(cap),
            qk.Mailbox.new()))
        # the trucks of load queue s dump into dump queue (s+1) % n: a branching, re-merging network
        for p in procs_all[(s + 1) % n_src]:
            qk.connect_components(p, dq)
        dk = df.create(qk.DistributionConfig.Uniform(0.5, 1.5))
        m.random_dists.append(dk)
        snk = m.reg(qk.ComponentModel.StringSink(
            qk.DiscreteSink.new().with_name("Sink%d" % s).with_code("SNK%d" % s).with_process_time_distr(dk),
            qk.Mailbox.new()))
        qk.connect_components(dq, snk)
    _log_all(m)
    return m



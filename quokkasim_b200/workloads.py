"""Synthetic workloads of the named shapes (BASELINE.json `configs`): the reference's example topologies
rebuilt with the mirrored builder API.  Used by bench.py and by the tests (no oracle involvement here).
"""
from . import api as qk
from . import vector_api as _va

# quokkasim_examples/src/bin/iron_ore.rs:21-96 -- a user-defined resource: five masses, total() = fe + other_elements --
# and the ComponentModel / ComponentLogger variants its define_model_enums! block declares (:258-270)
IronOre = qk.define_vector_variants("IronOre", qk.define_resource(
    "IronOre", ("fe", "other_elements", "magnetite", "hematite", "limonite"), total=("fe", "other_elements")))


class Model:
    """A built model: registered ComponentModels in registration order + scheduled env events."""

    def __init__(self):
        self.components = []
        self.ext = []  # (t_ns, env ComponentModel, state)
        self.ext_low = []  # (t_ns, F64Stock ComponentModel, new low_capacity)
        self.ext_qty = []  # (t_ns, F64Process ComponentModel, Distribution instance made by df.create at schedule time)
        self.random_dists = []  # Distribution objects that need tapes in tape mode
        self.by_name = {}

    def reg(self, cm):
        self.components.append(cm)
        self.by_name[cm.component.element_name] = cm
        return cm

    def sim(self, lib, n_reps, t0=0, base_seed=0, rep_offset=0, log_capacity=0, threads_per_block=0, stream=None, flags=0,
            device=0):
        si = qk.BatchedSimInit.new()
        for c in self.components:
            si = qk.register_component(si, c)
        sim, sched = si.init(t0, n_replications=n_reps, base_seed=base_seed, device=device, rep_offset=rep_offset,
                             log_capacity=log_capacity, threads_per_block=threads_per_block, stream=stream,
                             lib=lib, flags=flags)
        for t, env, state in self.ext:
            qk.create_scheduled_event(sched, t, qk.ScheduledEventConfig.SetEnvironmentState(state), env.get_address())
        for t, stock, value in self.ext_low:
            qk.create_scheduled_event(sched, t, qk.ScheduledEventConfig.SetLowCapacity(value), stock.get_address())
        for t, proc, dist in self.ext_qty:
            qk.create_scheduled_event(sched, t, qk.ScheduledEventConfig.SetProcessQuantity(dist), proc.get_address(),
                                      qk.DistributionFactory(base_seed=0, next_seed=0))
        return sim


def _log_all(m):
    sl = qk.ComponentLogger.StringStockLogger(qk.DiscreteStockLogger.new("StockLogger"))
    pl = qk.ComponentLogger.StringProcessLogger(qk.DiscreteProcessLogger.new("ProcessLogger"))
    el = qk.ComponentLogger.BasicEnvironmentLogger(qk.BasicEnvironmentLogger.new("EnvLogger"))
    vsl = {p: qk.ComponentLogger(p + "StockLogger", qk.VectorStockLogger.new(p + "StockLogger"))
           for p in _va.PREFIXES}
    vpl = {p: qk.ComponentLogger(p + "ProcessLogger", qk.VectorProcessLogger.new(p + "ProcessLogger"))
           for p in _va.PREFIXES}
    csl = qk.ComponentLogger.Vector3ContainerStockLogger(qk.ContainerStockLogger.new("ContainerStockLogger"))
    cpl = qk.ComponentLogger.Vector3ContainerProcessLogger(qk.ContainerProcessLogger.new("ContainerProcessLogger"))
    for c in m.components:
        prefix = _va.PREFIX_OF.get(c.variant)
        if c.variant.startswith("Vector3Container"):
            if c.variant == "Vector3ContainerStock":
                qk.connect_logger(csl, c)
            elif c.variant in cpl.logger.ACCEPTS:
                qk.connect_logger(cpl, c)
            elif getattr(m, "log_unloggable", False):
                # the reference has no logger arm for Vector3ContainerSource / Sink (core.rs:1331-1334); parity tests
                # still want their records compared, so they switch logging on below the Python API
                c.component._logged_by = cpl.logger
                cpl.logger.components.append(c.component)
        elif prefix:
            qk.connect_logger((vsl if c.variant.endswith("Stock") else vpl)[prefix], c)
        elif c.variant == "StringStock":
            qk.connect_logger(sl, c)
        elif c.variant == "BasicEnvironment":
            qk.connect_logger(el, c)
        else:
            qk.connect_logger(pl, c)
    m.loggers = (sl, pl, el, vsl, vpl)
    m.container_loggers = (csl, cpl)


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




# ------------------------------------------------------------------------------------------- vector models
def source_sink():
    """quokkasim_examples/src/bin/source_sink.rs: Vector3 Source -> Stock -> Process -> Stock -> Sink, all Constant.
    Registration order follows the example (stocks first)."""
    m = Model()
    c = qk.Distribution.Constant
    source = qk.ComponentModel.Vector3Source(
        qk.VectorSource.new().with_name("Source").with_process_quantity_distr(c(1.0)).with_process_time_distr(c(1.0))
        .with_source_vector(qk.Vector3([1.0, 4.0, 5.0])), qk.Mailbox.new())
    stock_1 = qk.ComponentModel.Vector3Stock(
        qk.VectorStock.new().with_name("Stock 1").with_low_capacity(50.0).with_max_capacity(101.0)
        .with_initial_resource(qk.Vector3([0.0, 0.0, 0.0])), qk.Mailbox.new())
    process = qk.ComponentModel.Vector3Process(
        qk.VectorProcess.new().with_name("Process").with_process_quantity_distr(c(1.0)).with_process_time_distr(c(1.0)),
        qk.Mailbox.new())
    stock_2 = qk.ComponentModel.Vector3Stock(
        qk.VectorStock.new().with_name("Stock 2").with_low_capacity(50.0).with_max_capacity(101.0)
        .with_initial_resource(qk.Vector3([0.0, 0.0, 0.0])), qk.Mailbox.new())
    sink = qk.ComponentModel.Vector3Sink(
        qk.VectorSink.new().with_name("Sink").with_process_quantity_distr(c(1.0)).with_process_time_distr(c(2.0)),
        qk.Mailbox.new())
    qk.connect_components(source, stock_1)
    qk.connect_components(stock_1, process)
    qk.connect_components(process, stock_2)
    qk.connect_components(stock_2, sink)
    for x in (stock_1, stock_2, source, process, sink):
        m.reg(x)
    _log_all(m)
    return m


def material_blending():
    """quokkasim_examples/src/bin/material_blending.rs: 3 stockpiles, Combiner2 x2, Combiner1, Splitter2, stocks shared by
    several processes, feedback loop through the stacker."""
    m = Model()
    c = qk.Distribution.Constant

    def pile(name, code, low, maxc, init):
        return qk.ComponentModel.Vector3Stock(
            qk.VectorStock.new().with_name(name).with_code(code).with_low_capacity(low).with_max_capacity(maxc)
            .with_initial_resource(qk.Vector3(init)), qk.Mailbox.new())

    sp1 = pile("Stockpile 1", "SP1", 100.0, 10000.0, [8000.0, 2000.0, 0.0])
    sp2 = pile("Stockpile 2", "SP2", 100.0, 10000.0, [0.0, 6000.0, 2000.0])
    sp3 = pile("Stockpile 3", "SP3", 100.0, 10000.0, [5000.0, 5000.0, 0.0])
    osp1 = pile("Output Stockpile 1", "OSP1", 100.0, 15000.0, [0.0, 0.0, 0.0])
    osp2 = pile("Output Stockpile 2", "OSP2", 100.0, 15000.0, [0.0, 0.0, 0.0])
    rc1 = qk.ComponentModel.Vector3Combiner2(
        qk.VectorCombiner.new().with_name("Reclaimer 1").with_code("RC1").with_process_quantity_distr(c(100.0))
        .with_process_time_distr(c(30.0)), qk.Mailbox.new())
    rc2 = qk.ComponentModel.Vector3Combiner2(
        qk.VectorCombiner.new().with_name("Reclaimer 2").with_code("RC2").with_process_quantity_distr(c(100.0))
        .with_process_time_distr(c(30.0)), qk.Mailbox.new())
    rc3 = qk.ComponentModel.Vector3Combiner1(
        qk.VectorCombiner.new().with_name("Reclaimer 3").with_code("RC3").with_process_quantity_distr(c(100.0))
        .with_process_time_distr(c(60.0)), qk.Mailbox.new())
    stk = qk.ComponentModel.Vector3Splitter2(
        qk.VectorSplitter.new().with_name("Stacker").with_code("STK").with_process_quantity_distr(c(600.0))
        .with_process_time_distr(c(360.0)), qk.Mailbox.new())
    qk.connect_components(sp1, rc1, 0)
    qk.connect_components(sp2, rc1, 1)
    qk.connect_components(rc1, osp1)
    qk.connect_components(sp2, rc2, 0)
    qk.connect_components(sp3, rc2, 1)
    qk.connect_components(rc2, osp2)
    qk.connect_components(sp1, rc3, 0)
    qk.connect_components(rc3, osp1)
    qk.connect_components(osp1, stk)
    qk.connect_components(stk, sp2, 0)
    qk.connect_components(stk, sp3, 1)
    for x in (sp1, sp2, sp3, rc1, rc2, rc3, stk, osp1, osp2):
        m.reg(x)
    _log_all(m)
    return m


def iron_ore():
    """quokkasim_examples/src/bin/iron_ore.rs:308-372: IronOreStock -> IronOreProcess -> IronOreStock, quantity
    Uniform(2, 8) from DistributionFactory(123456789), 10 s per batch; the example steps until 300 s."""
    m = Model()
    df = qk.DistributionFactory.new(123456789)
    qd = df.create(qk.DistributionConfig.Uniform(2.0, 8.0))
    m.random_dists.append(qd)
    s1 = qk.ComponentModel.IronOreStock(
        qk.VectorStock.new().with_name("MyStock1").with_code("SP1").with_type("IronOreStock")
        .with_initial_resource(IronOre(fe=60.0, other_elements=40.0, magnetite=10.0, hematite=5.0, limonite=15.0))
        .with_low_capacity(10.0).with_max_capacity(100.0), qk.Mailbox.new())
    p1 = qk.ComponentModel.IronOreProcess(
        qk.VectorProcess.new().with_name("MyProcess1").with_code("P1").with_name("IronOreProcess")
        .with_process_quantity_distr(qd).with_process_time_distr(qk.Distribution.Constant(10.0)), qk.Mailbox.new())
    s2 = qk.ComponentModel.IronOreStock(
        qk.VectorStock.new().with_name("MyStock2").with_code("SP2").with_type("IronOreStock")
        .with_initial_resource(IronOre(fe=3.0, other_elements=2.0, magnetite=0.5, hematite=0.25, limonite=0.75))
        .with_low_capacity(10.0).with_max_capacity(100.0), qk.Mailbox.new())
    qk.connect_components(s1, p1)
    qk.connect_components(p1, s2)
    for x in (s1, p1, s2):
        m.reg(x)
    _log_all(m)
    m.mass = {"vector_stocks": [s1, s2], "vector_procs": [p1]}
    return m


def scheduled_event_f64():
    """quokkasim_examples/src/bin/scheduled_event.rs in spirit: F64 Stock -> Process -> Stock with
    SetLowCapacity(10) on the first stock at 60 s."""
    m = Model()
    c = qk.Distribution.Constant
    s1 = qk.ComponentModel.F64Stock(
        qk.VectorStock.new().with_name("Stock1").with_code("S1").with_low_capacity(50.0).with_max_capacity(101.0)
        .with_initial_resource(100.0), qk.Mailbox.new())
    p = qk.ComponentModel.F64Process(
        qk.VectorProcess.new().with_name("Process").with_code("P").with_process_quantity_distr(c(1.0))
        .with_process_time_distr(c(1.0)), qk.Mailbox.new())
    s2 = qk.ComponentModel.F64Stock(
        qk.VectorStock.new().with_name("Stock2").with_code("S2").with_low_capacity(50.0).with_max_capacity(101.0)
        .with_initial_resource(0.0), qk.Mailbox.new())
    qk.connect_components(s1, p)
    qk.connect_components(p, s2)
    for x in (s1, p, s2):
        m.reg(x)
    m.ext_low.append((qk.Duration.from_secs(60), s1, 10.0))
    _log_all(m)
    return m


def scheduled_process_quantity(random=True):
    """create_scheduled_event!(SetProcessQuantity(..)) on an F64Process (core.rs:1387-1391): F64 Source -> Stock ->
    Process -> Stock -> Sink; the process moves U(0.5, 1.5) per cycle until 20 s, Tri(2, 4, 3) from then on, and a
    Constant(0.25) from 45 s.  The instances come from the factory at schedule time (seeds 2 and 3 after the model's
    own two)."""
    m = Model()
    df = qk.DistributionFactory(base_seed=31, next_seed=0)
    c = qk.Distribution.Constant

    def dist(cfg, const):
        if not random:
            return c(const)
        d = df.create(cfg)
        m.random_dists.append(d)
        return d

    src = qk.ComponentModel.F64Source(
        qk.VectorSource.new().with_name("Source").with_code("SRC").with_source_vector(1.0)
        .with_process_quantity_distr(c(2.0)).with_process_time_distr(c(1.0)), qk.Mailbox.new())
    s1 = qk.ComponentModel.F64Stock(
        qk.VectorStock.new().with_name("S1").with_code("S1").with_low_capacity(1.0).with_max_capacity(200.0)
        .with_initial_resource(20.0), qk.Mailbox.new())
    p = qk.ComponentModel.F64Process(
        qk.VectorProcess.new().with_name("Process").with_code("P")
        .with_process_quantity_distr(dist(qk.DistributionConfig.Uniform(0.5, 1.5), 1.0))
        .with_process_time_distr(dist(qk.DistributionConfig.Uniform(0.5, 1.5), 1.0)), qk.Mailbox.new())
    s2 = qk.ComponentModel.F64Stock(
        qk.VectorStock.new().with_name("S2").with_code("S2").with_low_capacity(1.0).with_max_capacity(500.0)
        .with_initial_resource(0.0), qk.Mailbox.new())
    snk = qk.ComponentModel.F64Sink(
        qk.VectorSink.new().with_name("Sink").with_code("SNK").with_process_quantity_distr(c(1.0))
        .with_process_time_distr(c(2.0)), qk.Mailbox.new())
    qk.connect_components(src, s1)
    qk.connect_components(s1, p)
    qk.connect_components(p, s2)
    qk.connect_components(s2, snk)
    for x in (src, s1, p, s2, snk):
        m.reg(x)
    m.ext_qty.append((qk.Duration.from_secs(20), p, dist(qk.DistributionConfig.Triangular(min=2.0, max=4.0, mode=3.0), 3.0)))
    m.ext_qty.append((qk.Duration.from_secs(45), p, c(0.25)))
    _log_all(m)
    return m


def vector_flow(dim=3, random=True, breakdowns=True, env_events=True, sink_delay=False):
    """BASELINE config C3: Source(qty U(0.5,1.5), time U(0.5,1.5), vector [1,4,5]) -> Stock(low 50, max 101) -> Process
    -> Stock -> Sink, plus a Combiner2 / Splitter2 diamond as in material_blending.rs:128-141; optional breakdown
    modes (transport.rs) and an environment stop/resume (delay_testing.rs)."""
    m = Model()
    df = qk.DistributionFactory(base_seed=555, next_seed=0)
    P = {1: "F64", 3: "Vector3", 5: "IronOre"}[dim]
    if dim == 5:  # the same topology carrying the user-defined resource of iron_ore.rs (two of five fields counted)
        V = lambda v: IronOre(v[0], v[1] + v[2], 0.5 * v[0], 0.25 * v[1], 0.125 * v[2] + 0.0625)  # noqa: E731
    else:
        V = (lambda v: qk.Vector3(v)) if dim == 3 else (lambda v: float(sum(v)))

    def dist(lo, hi):
        if not random:
            return qk.Distribution.Constant(0.5 * (lo + hi))
        d = df.create(qk.DistributionConfig.Uniform(lo, hi))
        m.random_dists.append(d)
        return d

    def cm(kind, comp):
        return getattr(qk.ComponentModel, P + kind)(comp, qk.Mailbox.new())

    def stock(name, low, maxc, init):
        return cm("Stock", qk.VectorStock.new().with_name(name).with_code(name).with_low_capacity(low)
                  .with_max_capacity(maxc).with_initial_resource(V(init)))

    src = cm("Source", qk.VectorSource.new().with_name("Source").with_code("SRC").with_source_vector(V([1.0, 4.0, 5.0]))
             .with_process_quantity_distr(dist(0.5, 1.5)).with_process_time_distr(dist(0.5, 1.5)))
    s1 = stock("S1", 5.0, 101.0, [3.0, 3.0, 3.0])
    proc = qk.VectorProcess.new().with_name("Process").with_code("PRC").with_process_quantity_distr(dist(0.8, 1.6)) \
        .with_process_time_distr(dist(0.5, 1.5))
    if breakdowns:
        proc = proc.with_delay_mode(qk.DelayModeChange.Add(qk.DelayMode("Breakdown", dist(20.0, 40.0), dist(2.0, 6.0))))
        proc = proc.with_delay_mode(qk.DelayModeChange.Add(qk.DelayMode("Service", dist(50.0, 70.0), dist(1.0, 3.0))))
    p = cm("Process", proc)
    s2 = stock("S2", 2.0, 60.0, [0.0, 0.0, 0.0])
    s3 = stock("S3", 2.0, 60.0, [10.0, 0.0, 5.0])
    comb = cm("Combiner2", qk.VectorCombiner.new().with_name("Combiner").with_code("CMB")
              .with_process_quantity_distr(dist(0.3, 0.9)).with_process_time_distr(dist(1.0, 2.0)))
    s4 = stock("S4", 1.0, 80.0, [0.0, 0.0, 0.0])
    split = qk.VectorSplitter.new().with_name("Splitter").with_code("SPL").with_split_ratios([0.3, 0.7]) \
        .with_process_quantity_distr(dist(0.5, 1.5)).with_process_time_distr(dist(0.7, 1.3))
    spl = cm("Splitter2", split)
    s5 = stock("S5", 0.5, 50.0, [0.0, 0.0, 0.0])
    snk_c = qk.VectorSink.new().with_name("Sink").with_code("SNK").with_process_quantity_distr(dist(0.4, 1.2)) \
        .with_process_time_distr(dist(0.5, 1.0))
    if sink_delay:
        snk_c = snk_c.with_delay_mode(qk.DelayModeChange.Add(qk.DelayMode("SinkDelay", dist(10.0, 20.0), dist(1.0, 2.0))))
    snk = cm("Sink", snk_c)
    qk.connect_components(src, s1)
    qk.connect_components(s1, p)
    qk.connect_components(p, s2)
    qk.connect_components(s2, comb, 0)
    qk.connect_components(s3, comb, 1)
    qk.connect_components(comb, s4)
    qk.connect_components(s4, spl)
    qk.connect_components(spl, s3, 0)  # feedback into the combiner's second input
    qk.connect_components(spl, s5, 1)
    qk.connect_components(s5, snk)
    comps = [src, s1, p, s2, s3, comb, s4, spl, s5, snk]
    if env_events:
        env = qk.ComponentModel.BasicEnvironment(qk.BasicEnvironment.new().with_name("Env").with_code("ENV"),
                                                 qk.Mailbox.new())
        for x in (src, p, comb, spl, snk):
            qk.connect_components(env, x)
        comps.append(env)
        m.ext.append((qk.Duration.from_secs(40), env, qk.BasicEnvironmentState.Stopped))
        m.ext.append((qk.Duration.from_secs(55), env, qk.BasicEnvironmentState.Normal))
    for x in comps:
        m.reg(x)
    _log_all(m)
    return m


def transport():
    """quokkasim_examples/src/bin/transport.rs: Vector3 Stock -> Process (two delay modes, Uniform until-delay,
    Constant until-fix) -> Stock; registration order stock_1, stock_2, process; seeds from DistributionFactory::new(1234)."""
    m = Model()
    c = qk.Distribution.Constant
    df = qk.DistributionFactory.new(1234)
    s1 = qk.ComponentModel.Vector3Stock(
        qk.VectorStock.new().with_name("Stock 1").with_low_capacity(10.0).with_max_capacity(100.0)
        .with_initial_resource(qk.Vector3([99.0, 0.0, 0.0])), qk.Mailbox.new())
    d1 = df.create(qk.DistributionConfig.Uniform(5.0, 10.0))
    d2 = df.create(qk.DistributionConfig.Uniform(40.0, 60.0))
    m.random_dists += [d1, d2]
    p = qk.ComponentModel.Vector3Process(
        qk.VectorProcess.new().with_name("Process").with_process_quantity_distr(c(1.0)).with_process_time_distr(c(1.0))
        .with_delay_mode(qk.DelayModeChange.Add(qk.DelayMode("ShortDelay", d1, c(1.0))))
        .with_delay_mode(qk.DelayModeChange.Add(qk.DelayMode("LongDelay", d2, c(5.0)))), qk.Mailbox.new())
    s2 = qk.ComponentModel.Vector3Stock(
        qk.VectorStock.new().with_name("Stock 2").with_low_capacity(10.0).with_max_capacity(100.0)
        .with_initial_resource(qk.Vector3([0.0, 0.0, 0.0])), qk.Mailbox.new())
    qk.connect_components(s1, p)
    qk.connect_components(p, s2)
    for x in (s1, s2, p):
        m.reg(x)
    _log_all(m)
    return m


def delay_testing():
    """quokkasim_examples/src/bin/delay_testing.rs: F64 Stock -> Process (DelayMode Constant 7 / 2) -> Stock, with a
    BasicEnvironment stopped at 4 s and resumed at 24 s; horizon 40 s."""
    m = Model()
    c = qk.Distribution.Constant
    s1 = qk.ComponentModel.F64Stock(
        qk.VectorStock.new().with_name("Stock1").with_code("S1").with_initial_resource(100.0).with_low_capacity(10.0)
        .with_max_capacity(1000.0), qk.Mailbox.new())
    p = qk.ComponentModel.F64Process(
        qk.VectorProcess.new().with_name("Process").with_code("P1").with_process_quantity_distr(c(10.0))
        .with_process_time_distr(c(10.0))
        .with_delay_mode(qk.DelayModeChange.Add(qk.DelayMode("ProcessDelay", c(7.0), c(2.0)))), qk.Mailbox.new())
    s2 = qk.ComponentModel.F64Stock(
        qk.VectorStock.new().with_name("Stock2").with_code("S2").with_low_capacity(10.0).with_max_capacity(1000.0),
        qk.Mailbox.new())
    env = qk.ComponentModel.BasicEnvironment(qk.BasicEnvironment.new().with_name("Environment").with_code("E1"),
                                             qk.Mailbox.new())
    qk.connect_components(s1, p)
    qk.connect_components(p, s2, 0)  # the example passes Some(0); the arm ignores it (core.rs:578-583)
    qk.connect_components(env, p)
    for x in (s1, s2, p, env):
        m.reg(x)
    m.ext.append((qk.Duration.from_secs(4), env, qk.BasicEnvironmentState.Stopped))
    m.ext.append((qk.Duration.from_secs(24), env, qk.BasicEnvironmentState.Normal))
    _log_all(m)
    return m


def the_example():
    """quokkasim_examples/src/bin/the_example.rs ("rocks"): two VectorSources into one stock, trucking process, a
    Splitter2 with a breakdown mode, two sinks (one with Uniform time and TruncNormal quantity); t0 = 2025-06-22."""
    m = Model()
    c = qk.Distribution.Constant
    df = qk.DistributionFactory.new(12345)
    V = qk.Vector3

    def cm(kind, comp):
        return getattr(qk.ComponentModel, "Vector3" + kind)(comp, qk.Mailbox.new())

    rs1 = cm("Source", qk.VectorSource.new().with_name("Rock Source 1").with_code("RS1").with_process_time_distr(c(600.0))
             .with_process_quantity_distr(c(5000.0)).with_source_vector(V([300.0, 100.0, 0.0])))
    rs2 = cm("Source", qk.VectorSource.new().with_name("Rock Source 2").with_code("RS2").with_process_time_distr(c(300.0))
             .with_process_quantity_distr(c(2000.0)).with_source_vector(V([200.0, 200.0, 0.0])))
    st1 = cm("Stock", qk.VectorStock.new().with_name("Rock Stock 1").with_code("RS1").with_low_capacity(100.0)
             .with_max_capacity(2000.0).with_initial_resource(V([1600.0, 400.0, 0.0])))
    trucking = cm("Process", qk.VectorProcess.new().with_name("Trucking").with_code("T1").with_process_time_distr(c(18.0))
                  .with_process_quantity_distr(c(60.0)))
    st2 = cm("Stock", qk.VectorStock.new().with_name("Rock Stock 2").with_code("RS2").with_low_capacity(100.0)
             .with_max_capacity(2000.0).with_initial_resource(V([1600.0, 400.0, 0.0])))
    ud = df.create(qk.DistributionConfig.Uniform(60.0, 180.0))
    uf = df.create(qk.DistributionConfig.Uniform(10.0, 60.0))
    conveyors = cm("Splitter2", qk.VectorSplitter.new().with_name("Conveyor").with_code("C1")
                   .with_process_time_distr(c(18.0)).with_process_quantity_distr(c(60.0))
                   .with_split_ratios([2.0 / 3.0, 1.0 / 3.0])
                   .with_delay_mode(qk.DelayModeChange.Add(qk.DelayMode("Short Delay", ud, uf))))
    pf1 = cm("Stock", qk.VectorStock.new().with_name("Process Feed 1").with_code("PF1").with_low_capacity(100.0)
             .with_max_capacity(500.0))
    p1 = cm("Sink", qk.VectorSink.new().with_name("Processing 1").with_code("P1").with_process_time_distr(c(18.0))
            .with_process_quantity_distr(c(60.0)))
    pf2 = cm("Stock", qk.VectorStock.new().with_name("Process Feed 2").with_code("PF2").with_low_capacity(100.0)
             .with_max_capacity(500.0))
    t2 = df.create(qk.DistributionConfig.Uniform(10.0, 26.0))
    q2 = df.create(qk.DistributionConfig.TruncNormal(30.0, 10.0, min=1.0, max=None))
    p2 = cm("Sink", qk.VectorSink.new().with_name("Processing 2").with_code("P2").with_process_time_distr(t2)
            .with_process_quantity_distr(q2))
    m.random_dists += [ud, uf, t2]  # q2 (TruncNormal) stays on its native stream in tape tests
    qk.connect_components(rs1, st1)
    qk.connect_components(rs2, st1)
    qk.connect_components(st1, trucking)
    qk.connect_components(trucking, st2)
    qk.connect_components(st2, conveyors)
    qk.connect_components(conveyors, pf1, 0)
    qk.connect_components(conveyors, pf2, 1)
    qk.connect_components(pf1, p1)
    qk.connect_components(pf2, p2)
    for x in (rs1, rs2, st1, trucking, st2, conveyors, pf1, p1, pf2, p2):
        m.reg(x)
    m.t0 = qk.MonotonicTime.try_from_date_time(2025, 6, 22, 0, 0, 0, 0)
    _log_all(m)
    return m


# ---------------------------------------------------------------------------------------- container models
def container_haulage(n_trucks=4, loaders=2, random=True, breakdowns=True, unload_ratio=(0.85, 0.98)):
    """Payload-true truck cycle: a fixed fleet of Vector3Containers circulates
    ReadyToLoad -> ContainerLoadingProcess (ore from a Vector3Stock) -> Loaded -> Vector3ContainerParallelProcess
    (haul) -> AtDump -> ContainerUnloadingProcess (ore into a Vector3Stock) -> Emptied -> Vector3ContainerProcess
    (return) -> ReadyToLoad.  With an unload ratio below 1 mass is conserved: ore + dumped + carried is constant."""
    m = Model()
    df = qk.DistributionFactory(base_seed=777, next_seed=0)

    def dist(kind, *a):
        if not random:
            return qk.Distribution.Constant(sum(a[:2]) / 2.0)
        d = df.create(getattr(qk.DistributionConfig, kind)(*a))
        m.random_dists.append(d)
        return d

    ore = m.reg(qk.ComponentModel.Vector3Stock(
        qk.VectorStock.new().with_name("Ore").with_code("ORE").with_low_capacity(0.0).with_max_capacity(1e9)
        .with_initial_resource(qk.Vector3([60000.0, 30000.0, 10000.0])), qk.Mailbox.new()))
    trucks = [qk.Vector3Container("Truck_%02d" % i, None if i % 3 else qk.Vector3([1.0, 0.5, 0.25]), 80.0 + 10.0 * (i % 4))
              for i in range(n_trucks)]
    ready = m.reg(qk.ComponentModel.Vector3ContainerStock(
        qk.DiscreteStock.new().with_name("ReadyToLoad").with_code("RTL").with_low_capacity(0)
        .with_max_capacity(n_trucks + 2).with_initial_resource(qk.ItemDeque(trucks)), qk.Mailbox.new()))
    load_c = qk.ContainerLoadingProcess.new().with_name("Loader").with_code("LD").with_max_process_count(loaders) \
        .with_process_time_distr(dist("Triangular", 20.0, 60.0, 30.0)) \
        .with_process_capacity_ratio_distr(dist("Uniform", 0.8, 1.0))
    if breakdowns:
        load_c = load_c.with_delay_mode(qk.DelayModeChange.Add(
            qk.DelayMode("Breakdown", dist("Uniform", 300.0, 600.0), dist("Uniform", 20.0, 60.0))))
    load = m.reg(qk.ComponentModel.Vector3ContainerLoadProcess(load_c, qk.Mailbox.new()))
    loaded = m.reg(qk.ComponentModel.Vector3ContainerStock(
        qk.DiscreteStock.new().with_name("Loaded").with_code("LDD").with_max_capacity(n_trucks + 2), qk.Mailbox.new()))
    haul = m.reg(qk.ComponentModel.Vector3ContainerParallelProcess(
        qk.DiscreteParallelProcess.new().with_name("Haul").with_code("HL")
        .with_process_time_distr(dist("Triangular", 100.0, 200.0, 140.0)), qk.Mailbox.new()))
    at_dump = m.reg(qk.ComponentModel.Vector3ContainerStock(
        qk.DiscreteStock.new().with_name("AtDump").with_code("ATD").with_max_capacity(n_trucks + 2), qk.Mailbox.new()))
    unload_c = qk.ContainerUnloadingProcess.new().with_name("Dump").with_code("DMP").with_max_process_count(1) \
        .with_process_time_distr(dist("Uniform", 15.0, 25.0))
    if unload_ratio is not None:
        unload_c = unload_c.with_process_quantity_ratio_distr(dist("Uniform", *unload_ratio))
    unload = m.reg(qk.ComponentModel.Vector3ContainerUnloadProcess(unload_c, qk.Mailbox.new()))
    dumped = m.reg(qk.ComponentModel.Vector3Stock(
        qk.VectorStock.new().with_name("Dumped").with_code("DPD").with_low_capacity(0.0).with_max_capacity(1e9),
        qk.Mailbox.new()))
    emptied = m.reg(qk.ComponentModel.Vector3ContainerStock(
        qk.DiscreteStock.new().with_name("Emptied").with_code("EMP").with_max_capacity(n_trucks + 2), qk.Mailbox.new()))
    ret = m.reg(qk.ComponentModel.Vector3ContainerProcess(
        qk.DiscreteProcess.new().with_name("Return").with_code("RET")
        .with_process_time_distr(dist("Triangular", 30.0, 70.0, 45.0)), qk.Mailbox.new()))
    env = m.reg(qk.ComponentModel.BasicEnvironment(qk.BasicEnvironment.new().with_name("Shift").with_code("ENV"),
                                                   qk.Mailbox.new()))
    qk.connect_components(ore, load)
    qk.connect_components(ready, load)
    qk.connect_components(load, loaded)
    qk.connect_components(loaded, haul)
    qk.connect_components(haul, at_dump)
    qk.connect_components(at_dump, unload)
    qk.connect_components(unload, emptied)
    qk.connect_components(unload, dumped)
    qk.connect_components(emptied, ret)
    qk.connect_components(ret, ready)
    for x in (load, unload, ret):
        qk.connect_components(env, x)
    m.ext.append((qk.Duration.from_secs(1500), env, qk.BasicEnvironmentState.Stopped))
    m.ext.append((qk.Duration.from_secs(1700), env, qk.BasicEnvironmentState.Normal))
    m.mass = dict(vector_stocks=[ore, dumped], container_holders=[ready, load, loaded, haul, at_dump, unload, emptied, ret])
    _log_all(m)
    return m


def container_open(lossy_unload=True, log_unloggable=True):
    """Open container line: Vector3ContainerSource (factory draws what each container arrives with) -> stock ->
    ContainerLoadingProcess (tops it up) -> stock -> ContainerUnloadingProcess -> stock -> Vector3ContainerSink.
    lossy_unload keeps the default process_quantity_ratio_distr Constant(1.): the reference then moves
    Vector3::default() to the resource stock and drops what the container carried (vector_container.rs:582-586)."""
    m = Model()
    m.log_unloggable = log_unloggable
    df = qk.DistributionFactory(base_seed=4242, next_seed=0)

    def dist(kind, *a):
        d = df.create(getattr(qk.DistributionConfig, kind)(*a))
        m.random_dists.append(d)
        return d

    fac = qk.Vector3ContainerFactory("Bin", 1, 3, dist("Uniform", 2.0, 6.0), [0.5, 0.3, 0.2], 25.0)
    src = m.reg(qk.ComponentModel.Vector3ContainerSource(
        qk.DiscreteSource.new().with_name("Arrivals").with_code("ARR").with_item_factory(fac)
        .with_process_time_distr(dist("Uniform", 4.0, 8.0)), qk.Mailbox.new()))
    q0 = m.reg(qk.ComponentModel.Vector3ContainerStock(
        qk.DiscreteStock.new().with_name("Q0").with_code("Q0").with_max_capacity(6), qk.Mailbox.new()))
    bulk = m.reg(qk.ComponentModel.Vector3Stock(
        qk.VectorStock.new().with_name("Bulk").with_code("BLK").with_low_capacity(10.0).with_max_capacity(5000.0)
        .with_initial_resource(qk.Vector3([900.0, 500.0, 100.0])), qk.Mailbox.new()))
    load = m.reg(qk.ComponentModel.Vector3ContainerLoadProcess(
        qk.ContainerLoadingProcess.new().with_name("Fill").with_max_process_count(2)
        .with_process_time_distr(dist("Triangular", 3.0, 9.0, 5.0)), qk.Mailbox.new()))
    q1 = m.reg(qk.ComponentModel.Vector3ContainerStock(
        qk.DiscreteStock.new().with_name("Q1").with_code("Q1").with_max_capacity(3), qk.Mailbox.new()))
    un_c = qk.ContainerUnloadingProcess.new().with_name("Empty").with_code("UNL").with_max_process_count(2) \
        .with_process_time_distr(dist("Uniform", 5.0, 12.0))
    if not lossy_unload:
        un_c = un_c.with_process_quantity_ratio_distr(dist("Uniform", 0.2, 0.9))
    unl = m.reg(qk.ComponentModel.Vector3ContainerUnloadProcess(un_c, qk.Mailbox.new()))
    out = m.reg(qk.ComponentModel.Vector3Stock(
        qk.VectorStock.new().with_name("Out").with_code("OUT").with_low_capacity(0.0).with_max_capacity(400.0),
        qk.Mailbox.new()))
    q2 = m.reg(qk.ComponentModel.Vector3ContainerStock(
        qk.DiscreteStock.new().with_name("Q2").with_code("Q2").with_max_capacity(2), qk.Mailbox.new()))
    snk = m.reg(qk.ComponentModel.Vector3ContainerSink(
        qk.DiscreteSink.new().with_name("Leave").with_code("LV")
        .with_process_time_distr(dist("Uniform", 6.0, 10.0)), qk.Mailbox.new()))
    qk.connect_components(src, q0)
    qk.connect_components(q0, load)
    qk.connect_components(bulk, load)
    qk.connect_components(load, q1)
    qk.connect_components(q1, unl)
    qk.connect_components(unl, q2)
    qk.connect_components(unl, out)
    qk.connect_components(q2, snk)
    _log_all(m)
    return m

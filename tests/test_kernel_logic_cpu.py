"""Kernel LOGIC on the CPU: the real kernel source (quokkasim_b200/csrc/qk_kernels.cuh) compiled with g++ under the
test-only shim in tests/emul, driven through the same C ABI and Python wrapper as the GPU library, compared bit
for bit with the oracle.  This keeps the CPU suite meaningful between GPU sessions; the parity tests proper are
tests/test_gpu_parity.py (-m gpu), which run the CUDA build of the same source."""
import numpy as np
import pytest

import helpers as H
from quokkasim_b200 import _capi as capi


def _check(lib, model, n_reps, n_events=None, t_until=None, tape_len=None, t0=0, seed=42, log_capacity=65536,
           flags=0):
    om, dist_ids = H.lower_to_oracle(model)
    tapes = H.make_tapes(model, n_reps, tape_len, seed=1) if tape_len else None
    sim = model.sim(lib, n_reps, t0=t0, base_seed=seed, log_capacity=log_capacity, flags=flags)
    if tapes:
        for d, v in tapes.values():
            sim.set_tape(d, v)
    if n_events is not None:
        sim.step_events(n_events)
    if t_until is not None:
        sim.step_until(t_until)
    for r in range(n_reps):
        osim = H.run_oracle(model, om, dist_ids, r, t0=t0, base_seed=seed, tapes=tapes, n_events=n_events,
                            t_until=t_until)
        H.assert_rep_parity(sim, osim, model, r)
    s = sim.summary()
    sim.close()
    return s


def test_discrete_queue_tape(emul_lib):
    s = _check(emul_lib, H.discrete_queue(), 6, n_events=1000, tape_len=400, log_capacity=4096)
    assert s["n_events"] == 6000


def test_discrete_queue_horizon(emul_lib):
    _check(emul_lib, H.discrete_queue(), 4, t_until=200 * 10**9, tape_len=200, log_capacity=2048)


def test_discrete_queue_native(emul_lib):
    _check(emul_lib, H.discrete_queue(), 4, n_events=1500, log_capacity=8192)


def test_hbm_resident_state_path(emul_lib):
    _check(emul_lib, H.discrete_queue(), 3, n_events=1000, log_capacity=4096, flags=1)
    _check(emul_lib, H.haulage(), 2, n_events=3000, log_capacity=16384, flags=1)


def test_eager_refill_path(emul_eager_lib):
    """deferred variate prefetch: results do not depend on WHEN the refill phase runs"""
    _check(emul_eager_lib, H.discrete_queue(), 3, n_events=1200, log_capacity=8192)
    _check(emul_eager_lib, H.serial_line(), 2, n_events=6000, tape_len=3000, log_capacity=32768)
    _check(emul_eager_lib, H.haulage(), 2, n_events=4000, log_capacity=16384)


def test_car_workshop(emul_lib):
    m = H.car_workshop()
    _check(emul_lib, m, 2, t_until=m.t0 + 9 * 3600 * 10**9, t0=m.t0, log_capacity=2048)


def test_assembly_line_parallel_env(emul_lib):
    m = H.assembly_line()
    _check(emul_lib, m, 2, t_until=m.t0 + 3 * 24 * 3600 * 10**9, t0=m.t0, log_capacity=16384)
    m = H.assembly_line()
    _check(emul_lib, m, 2, t_until=m.t0 + 24 * 3600 * 10**9, t0=m.t0, tape_len=800, log_capacity=8192)
    m = H.assembly_line(parallel=False)
    _check(emul_lib, m, 2, t_until=m.t0 + 24 * 3600 * 10**9, t0=m.t0, log_capacity=8192)


def test_serial_line_breakdowns(emul_lib):
    _check(emul_lib, H.serial_line(), 3, n_events=6000, tape_len=3000, log_capacity=32768)
    _check(emul_lib, H.serial_line(), 3, n_events=15000, log_capacity=65536)
    _check(emul_lib, H.serial_line(exp=False), 2, n_events=8000, log_capacity=32768)


def test_haulage_network(emul_lib):
    _check(emul_lib, H.haulage(), 3, n_events=5000, tape_len=3000, log_capacity=16384)
    _check(emul_lib, H.haulage(n_src=3, per_queue=4), 2, n_events=12000, log_capacity=65536)


def test_split_steps_and_logging_off_agree(emul_lib):
    """trajectories do not depend on how the run is cut into launches, nor on whether logging is compiled in"""
    m = H.discrete_queue()
    a = m.sim(emul_lib, 3, base_seed=5, log_capacity=0)
    a.step_events(400)
    a.step_events(1)
    a.step_events(599)
    b = H.discrete_queue().sim(emul_lib, 3, base_seed=5, log_capacity=4096)
    b.step_events(1000)
    for ci in range(len(m.components)):
        assert a.comp_stats(ci).tobytes() == b.comp_stats(ci).tobytes()
    assert a.status()[1].tobytes() == b.status()[1].tobytes()


def test_log_ring_keeps_the_newest_records(emul_lib):
    m = H.discrete_queue()
    full = m.sim(emul_lib, 2, base_seed=8, log_capacity=8192)
    ring = H.discrete_queue().sim(emul_lib, 2, base_seed=8, log_capacity=64)
    full.step_events(1200)
    ring.step_events(1200)
    lf, tf = full.read_log(1)
    lr, tr = ring.read_log(1)
    assert tf == tr and len(lr) == 64 and lr.tobytes() == lf[-64:].tobytes()


def test_faults(emul_lib):
    m = H.discrete_queue()
    tapes = H.make_tapes(m, 4, 300, seed=9)
    _, v0 = list(tapes.values())[0]
    v0[1, 3] = 0.0
    v0[2, 2] = -1.0
    v0[3, 5] = float("nan")
    om, dist_ids = H.lower_to_oracle(m)
    sim = m.sim(emul_lib, 4, log_capacity=4096)
    for d, v in tapes.values():
        sim.set_tape(d, v)
    sim.step_events(600)
    st, _, _ = sim.status()
    assert list(st) == [0, 1, 2, 2]
    for r in range(4):
        H.assert_rep_parity(sim, H.run_oracle(m, om, dist_ids, r, tapes=tapes, n_events=600), m, r)


def test_stock_with_initial_items_and_low_capacity(emul_lib):
    """with_initial_resource + low_capacity > 0 (Empty means `len <= low`, discrete.rs:164)"""
    import quokkasim_b200 as qk
    from quokkasim_b200.workloads import Model, _log_all

    m = Model()
    q1 = m.reg(qk.ComponentModel.StringStock(
        qk.DiscreteStock.new().with_name("Q1").with_low_capacity(2).with_max_capacity(6)
        .with_initial_resource(qk.ItemDeque(["a", "b", "c", "d", "e"])), qk.Mailbox.new()))
    p = m.reg(qk.ComponentModel.StringProcess(
        qk.DiscreteProcess.new().with_name("P").with_process_time_distr(qk.Distribution.Constant(2.0)),
        qk.Mailbox.new()))
    q2 = m.reg(qk.ComponentModel.StringStock(
        qk.DiscreteStock.new().with_name("Q2").with_max_capacity(2), qk.Mailbox.new()))
    snk = m.reg(qk.ComponentModel.StringSink(
        qk.DiscreteSink.new().with_name("S").with_process_time_distr(qk.Distribution.Constant(5.0)), qk.Mailbox.new()))
    src = m.reg(qk.ComponentModel.StringSource(
        qk.DiscreteSource.new().with_name("Src").with_process_time_distr(qk.Distribution.Constant(7.0)),
        qk.Mailbox.new()))
    qk.connect_components(q1, p)
    qk.connect_components(p, q2)
    qk.connect_components(q2, snk)
    qk.connect_components(src, q1)
    _log_all(m)
    _check(emul_lib, m, 1, t_until=300 * 10**9, log_capacity=4096)


def test_unconnected_ports(emul_lib):
    """"Upstream is not connected" / "Downstream is not connected" arms (discrete.rs:504-515, 868-871, 1121-1124)"""
    import quokkasim_b200 as qk
    from quokkasim_b200.workloads import Model, _log_all

    m = Model()
    m.reg(qk.ComponentModel.StringSource(qk.DiscreteSource.new().with_name("LonelySource"), qk.Mailbox.new()))
    m.reg(qk.ComponentModel.StringProcess(qk.DiscreteProcess.new().with_name("LonelyProcess"), qk.Mailbox.new()))
    m.reg(qk.ComponentModel.StringSink(qk.DiscreteSink.new().with_name("LonelySink"), qk.Mailbox.new()))
    m.reg(qk.ComponentModel.StringParallelProcess(qk.DiscreteParallelProcess.new().with_name("LonelyPar"),
                                                  qk.Mailbox.new()))
    q = m.reg(qk.ComponentModel.StringStock(qk.DiscreteStock.new().with_name("Q").with_max_capacity(3)
                                            .with_initial_resource(qk.ItemDeque(["x", "y"])), qk.Mailbox.new()))
    half = m.reg(qk.ComponentModel.StringProcess(qk.DiscreteProcess.new().with_name("NoDownstream"), qk.Mailbox.new()))
    qk.connect_components(q, half)
    _log_all(m)
    s = _check(emul_lib, m, 1, t_until=50 * 10**9, log_capacity=256)
    assert s["n_log_records"] >= 5


# ------------------------------------------------------------------------------------- continuous components
S = 10**9


def test_vector_source_sink_example(emul_lib):
    """source_sink.rs: Vector3 Source -> Stock -> Process -> Stock -> Sink, all Constant, 120 s"""
    _check(emul_lib, H.source_sink(), 2, t_until=120 * S, log_capacity=8192)


def test_vector_material_blending_example(emul_lib):
    """material_blending.rs: Combiner2 x2 + Combiner1 + Splitter2 on shared stockpiles, feedback loop, 1 h"""
    _check(emul_lib, H.material_blending(), 2, t_until=3600 * S, log_capacity=8192)


def test_vector_scheduled_low_capacity(emul_lib):
    """scheduled_event.rs: SetLowCapacity on an F64Stock at 60 s (core.rs:1382-1386)"""
    _check(emul_lib, H.scheduled_event_f64(), 2, t_until=120 * S, log_capacity=2048)


def test_user_defined_resource_iron_ore_example(emul_lib):
    """iron_ore.rs: a five-field resource whose total counts two fields (feature set 3, qk_model_set_vector_n):
    IronOreStock -> IronOreProcess -> IronOreStock until 300 s, native streams and a tape"""
    _check(emul_lib, H.iron_ore(), 3, t_until=300 * S, log_capacity=2048)
    _check(emul_lib, H.iron_ore(), 2, t_until=300 * S, tape_len=64, log_capacity=2048)
    for flags in (1, 16):
        _check(emul_lib, H.iron_ore(), 2, t_until=300 * S, log_capacity=2048, flags=flags)


def test_user_defined_resource_flow(emul_lib):
    """the C3 topology (source, process with breakdowns, combiner, splitter with feedback, sink, environment) carrying
    the IronOre resource: every vector code path at five fields"""
    _check(emul_lib, H.vector_flow(5), 3, n_events=4000, log_capacity=65536)
    _check(emul_lib, H.vector_flow(5), 2, n_events=2500, tape_len=2500, log_capacity=65536, flags=16)


@pytest.mark.parametrize("dim", [1, 3])
def test_vector_flow_native(emul_lib, dim):
    """config C3 topology: random quantities and times, two breakdown modes, environment stop/resume"""
    _check(emul_lib, H.vector_flow(dim), 3, n_events=4000, log_capacity=32768)


def test_vector_flow_tape(emul_lib):
    _check(emul_lib, H.vector_flow(3), 3, n_events=2500, tape_len=2500, log_capacity=32768)


def test_vector_flow_sink_delay_and_constant(emul_lib):
    """VectorSink's own post rule (vector.rs:1780-1786) under a delay mode; all-Constant variant -> exact ties"""
    _check(emul_lib, H.vector_flow(3, sink_delay=True), 2, n_events=3000, log_capacity=32768)
    _check(emul_lib, H.vector_flow(3, random=False), 1, n_events=3000, log_capacity=32768)
    _check(emul_lib, H.vector_flow(1, random=False, breakdowns=False, env_events=False), 1, n_events=2000,
           log_capacity=32768)


def test_vector_hbm_resident_and_eager(emul_lib, emul_eager_lib):
    _check(emul_lib, H.vector_flow(3), 2, n_events=2000, log_capacity=32768, flags=1)
    _check(emul_eager_lib, H.vector_flow(3), 2, n_events=2000, log_capacity=32768)


def test_chunked_stepping_equals_one_launch(emul_lib):
    """qk_batch_step_events_chunked: K events per state residency, same trajectories (bench --events-per-launch)"""
    m = H.discrete_queue()
    a = m.sim(emul_lib, 3, base_seed=17, log_capacity=8192)
    a.step_events(900)
    for chunk in (1, 7, 900):
        b = H.discrete_queue().sim(emul_lib, 3, base_seed=17, log_capacity=8192)
        b.step_events_chunked(900, chunk)
        for ci in range(len(m.components)):
            assert a.comp_stats(ci).tobytes() == b.comp_stats(ci).tobytes()
        assert a.read_log(2)[0].tobytes() == b.read_log(2)[0].tobytes()
        b.close()
    a.close()


def test_reference_vector_examples_transport_delay_testing_rocks(emul_lib):
    """transport.rs (120 s), delay_testing.rs (40 s), the_example.rs (700 s from 2025-06-22): the remaining
    reference examples that are built from the stock component set"""
    _check(emul_lib, H.transport(), 3, t_until=120 * S, log_capacity=8192)
    _check(emul_lib, H.transport(), 2, t_until=120 * S, tape_len=200, log_capacity=8192)
    _check(emul_lib, H.delay_testing(), 1, t_until=40 * S, log_capacity=1024)
    m = H.the_example()
    _check(emul_lib, m, 3, t_until=m.t0 + 700 * S, t0=m.t0, log_capacity=8192)


# ------------------------------------------------------------------------------------- container family (row f3)
def test_container_haulage_cycle(emul_lib):
    """closed truck cycle: ContainerLoadingProcess / ParallelProcess / ContainerUnloadingProcess / Process on a fleet
    of initial Vector3Containers, breakdowns on the loader, environment stop/resume"""
    _check(emul_lib, H.container_haulage(), 3, n_events=3000, log_capacity=65536)
    _check(emul_lib, H.container_haulage(n_trucks=7, loaders=3), 2, n_events=3000, tape_len=3000, log_capacity=65536)
    _check(emul_lib, H.container_haulage(random=False, breakdowns=False), 1, n_events=1500, log_capacity=32768)


def test_container_open_line(emul_lib, emul_eager_lib):
    """Vector3ContainerSource with a factory quantity distribution -> Load -> Unload -> Sink; the default unload ratio
    Constant(1.) reproduces the reference's lost-resource quirk (vector_container.rs:582-586)"""
    _check(emul_lib, H.container_open(), 3, n_events=3000, log_capacity=65536)
    _check(emul_lib, H.container_open(lossy_unload=False), 2, n_events=3000, log_capacity=65536)
    _check(emul_eager_lib, H.container_open(lossy_unload=False), 2, n_events=2000, tape_len=2500, log_capacity=65536)
    _check(emul_lib, H.container_open(), 2, n_events=2000, log_capacity=65536, flags=1)


def test_container_mass_balance_and_lossy_quirk(emul_lib):
    """with an unload ratio below 1 nothing is lost: ore + dumped + carried by the fleet is constant"""
    import numpy as np

    def total(sim, m, r):
        t = sum(sum(sim.read_vector(c.component._id, r)) for c in m.mass["vector_stocks"])
        for c in m.mass["container_holders"]:
            _, res = sim.read_containers(c.component._id, r)
            for row in res:
                if int(row[0]) != capi.CONTAINER_NONE_BITS:
                    t += float(row.view(np.float64).sum())
        return t

    m = H.container_haulage()
    sim = m.sim(emul_lib, 2, base_seed=3)
    t0 = [total(sim, m, r) for r in range(2)]
    sim.step_events(4000)
    for r in range(2):
        assert abs(total(sim, m, r) - t0[r]) <= 1e-9 * t0[r]
    dumped = m.by_name["Dumped"].component._id
    assert sum(sim.read_vector(dumped, 0)) > 100.0
    # default ratio Constant(1.): the resource stock receives zeros, the load vanishes
    m2 = H.container_haulage(unload_ratio=None)
    sim2 = m2.sim(emul_lib, 1, base_seed=3)
    sim2.step_events(4000)
    assert sum(sim2.read_vector(m2.by_name["Dumped"].component._id, 0)) == 0.0
    assert total(sim2, m2, 0) < t0[0] - 100.0


def test_container_connection_rules():
    """connect_components / connect_logger arms of the container family (core.rs:923-995, 1149-1172, 1330-1334)"""
    import pytest as _pt
    import quokkasim_b200 as qk

    cs = qk.ComponentModel.Vector3ContainerStock(qk.DiscreteStock.new(), qk.Mailbox.new())
    ss = qk.ComponentModel.StringStock(qk.DiscreteStock.new(), qk.Mailbox.new())
    vs = qk.ComponentModel.Vector3Stock(qk.VectorStock.new(), qk.Mailbox.new())
    fs = qk.ComponentModel.F64Stock(qk.VectorStock.new(), qk.Mailbox.new())
    ld = qk.ComponentModel.Vector3ContainerLoadProcess(qk.ContainerLoadingProcess.new(), qk.Mailbox.new())
    un = qk.ComponentModel.Vector3ContainerUnloadProcess(qk.ContainerUnloadingProcess.new(), qk.Mailbox.new())
    pp = qk.ComponentModel.Vector3ContainerParallelProcess(qk.DiscreteParallelProcess.new(), qk.Mailbox.new())
    sp = qk.ComponentModel.StringProcess(qk.DiscreteProcess.new(), qk.Mailbox.new())
    src = qk.ComponentModel.Vector3ContainerSource(qk.DiscreteSource.new(), qk.Mailbox.new())
    env = qk.ComponentModel.BasicEnvironment(qk.BasicEnvironment.new(), qk.Mailbox.new())
    assert isinstance(src.component.item_factory, qk.Vector3ContainerFactory)
    for a, b in ((cs, ld), (vs, ld), (ld, cs), (cs, un), (un, cs), (un, vs), (cs, pp), (pp, cs), (src, cs), (env, ld),
                 (env, un), (env, src)):
        qk.connect_components(a, b)
    for a, b in ((ss, ld), (cs, sp), (fs, ld), (un, fs), (ld, vs), (vs, un), (env, pp), (src, ss)):
        with _pt.raises(qk.ComponentConnectionError, match="No component connection defined"):
            qk.connect_components(a, b)
    cpl = qk.ComponentLogger.Vector3ContainerProcessLogger(qk.ContainerProcessLogger.new("L"))
    qk.connect_logger(cpl, ld)
    qk.connect_logger(cpl, pp)
    with _pt.raises(qk.ComponentConnectionError):
        qk.connect_logger(cpl, src)  # no Source / Sink arm in the reference
    assert ld.component.element_code == "CLP" and un.component.element_code == ""


def _mini_container_model(kind):
    import quokkasim_b200 as qk
    from quokkasim_b200.workloads import Model, _log_all

    m = Model()
    C = qk.Distribution.Constant

    def cstock(name, maxc, n_init=0, low=0):
        items = [qk.Vector3Container("%s_%d" % (name, i), qk.Vector3([1.0 + i, 2.0, 0.5]) if i % 2 else None, 10.0 + i)
                 for i in range(n_init)]
        return m.reg(qk.ComponentModel.Vector3ContainerStock(
            qk.DiscreteStock.new().with_name(name).with_code(name).with_low_capacity(low).with_max_capacity(maxc)
            .with_initial_resource(qk.ItemDeque(items)), qk.Mailbox.new()))

    def vstock(name, init, maxc=1000.0):
        return m.reg(qk.ComponentModel.Vector3Stock(
            qk.VectorStock.new().with_name(name).with_code(name).with_max_capacity(maxc)
            .with_initial_resource(qk.Vector3(init)), qk.Mailbox.new()))

    def load(name, n, t):
        return m.reg(qk.ComponentModel.Vector3ContainerLoadProcess(
            qk.ContainerLoadingProcess.new().with_name(name).with_max_process_count(n).with_process_time_distr(C(t)),
            qk.Mailbox.new()))

    def unload(name, n, t, ratio=0.5):
        return m.reg(qk.ComponentModel.Vector3ContainerUnloadProcess(
            qk.ContainerUnloadingProcess.new().with_name(name).with_code(name).with_max_process_count(n)
            .with_process_time_distr(C(t)).with_process_quantity_ratio_distr(C(ratio)), qk.Mailbox.new()))

    if kind == "upstream_full":  # 3 of max 3 -> Full -> "Upstream is full" (the Empty|Normal test, vector_container.rs:352)
        a, b, l = cstock("A", 3, 3), vstock("B", [5.0, 5.0, 5.0]), load("L", 2, 4.0)
        qk.connect_components(a, l)
        qk.connect_components(b, l)
    elif kind == "no_resource_stock":  # withdraw_us_resource...unwrap() on nothing -> fault
        a, l, q = cstock("A", 5, 2), load("L", 2, 4.0), cstock("Q", 5)
        qk.connect_components(a, l)
        qk.connect_components(l, q)
    elif kind == "downstream_full":  # drain loop loses containers when the downstream stock is Full
        a, b, l, q = cstock("A", 6, 4), vstock("B", [50.0, 5.0, 5.0]), load("L", 3, 5.0), cstock("Q", 1)
        snk = m.reg(qk.ComponentModel.Vector3ContainerSink(
            qk.DiscreteSink.new().with_name("S").with_process_time_distr(C(100.0)), qk.Mailbox.new()))
        for x, y in ((a, l), (b, l), (l, q), (q, snk)):
            qk.connect_components(x, y)
    elif kind == "lonely":
        load("L", 1, 1.0)
        unload("U", 1, 1.0)
        m.reg(qk.ComponentModel.Vector3ContainerParallelProcess(qk.DiscreteParallelProcess.new().with_name("P"),
                                                                qk.Mailbox.new()))
        m.reg(qk.ComponentModel.Vector3ContainerProcess(qk.DiscreteProcess.new().with_name("X"), qk.Mailbox.new()))
    elif kind == "unload_half_connected":  # "Downstream is not connected" / "Downstream resource stock is full"
        a, u1, q = cstock("A", 8, 3, low=0), unload("U1", 2, 3.0), cstock("Q", 8)
        qk.connect_components(a, u1)
        qk.connect_components(u1, q)  # no resource stock: (_, _) arm
        a2, u2, v = cstock("A2", 8, 3), unload("U2", 2, 3.0), vstock("V", [1.0, 1.0, 1.0], maxc=3.0)
        qk.connect_components(a2, u2)
        qk.connect_components(u2, v)  # no container stock; V is Full (3 of 3): (_, Some(Full)) arm
    m.log_unloggable = True
    _log_all(m)
    return m


@pytest.mark.parametrize("kind", ["upstream_full", "no_resource_stock", "downstream_full", "lonely",
                                  "unload_half_connected"])
def test_container_edge_arms(emul_lib, kind):
    """the match arms the two example models do not reach (vector_container.rs:294-301, 358, 372-379, 600-611, 683-686)"""
    m = _mini_container_model(kind)
    s = _check(emul_lib, m, 1, t_until=400 * 10**9, log_capacity=4096)
    if kind == "no_resource_stock":
        assert s["n_faulted"] == 1


def test_large_models_on_hbm_planes_with_ties(emul_lib):
    """> 16 schedulable components on the HBM planes (the placement the product picks for them), including exact
    same-timestamp ties between many components (constant tapes): the scheduler scan must keep registration order."""
    import numpy as np

    _check(emul_lib, H.haulage(n_src=3, per_queue=4), 2, n_events=6000, log_capacity=65536, flags=1)
    _check(emul_lib, H.serial_line(n_proc=8), 2, n_events=8000, log_capacity=65536, flags=1)
    _check(emul_lib, H.container_haulage(n_trucks=6), 2, n_events=3000, log_capacity=65536, flags=1)
    # every process time identical -> many components fire at the same instant, in different groups
    m = H.haulage(n_src=4, per_queue=5)
    om, dist_ids = H.lower_to_oracle(m)
    tapes = {id(d): (d, np.full((2, 3000), 2.0 if i % 2 else 4.0)) for i, d in enumerate(m.random_dists)}
    sim = m.sim(emul_lib, 2, log_capacity=65536, flags=1)
    for d, v in tapes.values():
        sim.set_tape(d, v)
    sim.step_events(5000)
    assert sim.info()["state_in_hbm"] == 1
    for r in range(2):
        H.assert_rep_parity(sim, H.run_oracle(m, om, dist_ids, r, tapes=tapes, n_events=5000), m, r)
    sim.close()


def test_wave_sliced_stepping_equals_one_step(emul_lib):
    """the (block, chunk) work-item schedule of qk_batch_step_events_sliced: same trajectories as one launch, for
    chunk counts that do and do not divide the event budget"""
    for chunks in (1, 3, 7, 16):
        m = H.discrete_queue()
        a = m.sim(emul_lib, 5, base_seed=11, log_capacity=8192)
        a.step_events(1000)
        b = H.discrete_queue().sim(emul_lib, 5, base_seed=11, log_capacity=8192)
        b.step_events_sliced(1000, chunks)
        assert (b.status()[2] == 1000).all()
        for r in range(5):
            assert a.read_log(r)[0].tobytes() == b.read_log(r)[0].tobytes()
        for ci in range(len(m.components)):
            assert a.comp_stats(ci).tobytes() == b.comp_stats(ci).tobytes()


def test_read_logs_equals_read_log(emul_lib):
    m = H.discrete_queue()
    s = m.sim(emul_lib, 6, base_seed=3, log_capacity=64)
    s.step_events(300)
    logs, totals = s.read_logs(1, 4)
    for i in range(4):
        one, tot = s.read_log(1 + i)
        assert tot == totals[i] and logs[i].tobytes() == one.tobytes()


# ---------------------------------------------------------------------------------------------- lane-group kernel
# qk_group.cuh under the host shim: every lane is a host thread, the block's only warp is the team, shuffles / ballots
# go through an exchange array between barriers.  flags = QK_BATCH_STATE_LANE_GROUP | lanes << 8.
def _gflags(lanes, phased=False):
    return 4 | (8 if phased else 0) | (lanes << 8)


@pytest.mark.parametrize("lanes", [2, 4, 8])
def test_group_kernel_discrete_queue(emul_lib, lanes):
    s = _check(emul_lib, H.discrete_queue(), 5, n_events=1200, log_capacity=8192, flags=_gflags(lanes))
    assert s["n_events"] == 6000
    _check(emul_lib, H.discrete_queue(), 3, t_until=150 * 10**9, tape_len=200, log_capacity=2048, flags=_gflags(lanes))


def test_group_kernel_shared_stocks_and_ties(emul_lib):
    m = H.car_workshop()
    _check(emul_lib, m, 3, t_until=m.t0 + 9 * 3600 * 10**9, t0=m.t0, log_capacity=2048, flags=_gflags(4))
    _check(emul_lib, H.haulage(), 3, n_events=5000, tape_len=3000, log_capacity=16384, flags=_gflags(4))
    _check(emul_lib, H.haulage(n_src=3, per_queue=4), 2, n_events=8000, log_capacity=65536, flags=_gflags(8))


def test_group_kernel_general_features(emul_lib):
    """delay modes, environment events, parallel process, vector and container components on the group's lane 0"""
    _check(emul_lib, H.serial_line(), 3, n_events=6000, tape_len=3000, log_capacity=32768, flags=_gflags(4))
    _check(emul_lib, H.serial_line(), 2, n_events=9000, log_capacity=65536, flags=_gflags(2))
    m = H.assembly_line()
    _check(emul_lib, m, 2, t_until=m.t0 + 24 * 3600 * 10**9, t0=m.t0, log_capacity=16384, flags=_gflags(4))
    _check(emul_lib, H.vector_flow(3), 2, n_events=2500, log_capacity=32768, flags=_gflags(4))
    _check(emul_lib, H.container_haulage(), 2, n_events=1500, log_capacity=32768, flags=_gflags(4))


def test_group_kernel_statistics_only_and_split_steps(emul_lib):
    m = H.haulage()
    a = m.sim(emul_lib, 4, base_seed=5, log_capacity=0, flags=_gflags(4))
    a.step_events(700)
    a.step_events(1)
    a.step_events_sliced(1299, 4)
    b = H.haulage().sim(emul_lib, 4, base_seed=5, log_capacity=16384)
    b.step_events(2000)
    for ci in range(len(m.components)):
        assert a.comp_stats(ci).tobytes() == b.comp_stats(ci).tobytes()
    assert a.status()[1].tobytes() == b.status()[1].tobytes()
    for ci, c in enumerate(m.components):
        if c.variant == "StringStock":
            for r in range(4):
                assert list(a.read_stock(ci, r)) == list(b.read_stock(ci, r))


def test_group_kernel_faults(emul_lib):
    m = H.discrete_queue()
    tapes = H.make_tapes(m, 4, 300, seed=9)
    _, v0 = list(tapes.values())[0]
    v0[1, 3] = 0.0
    v0[2, 2] = -1.0
    sim = m.sim(emul_lib, 4, log_capacity=4096, flags=_gflags(4))
    for d, v in tapes.values():
        sim.set_tape(d, v)
    sim.step_events(600)
    st, _, _ = sim.status()
    assert list(st) == [0, 1, 2, 0]


def test_delay_mode_quick_paths_are_taken(emul_lib):
    """the quick handling of components WITH delay modes (qk_dm_quick: busy tick, idle and blocked) is exercised by the
    serial line, in both kernels, and agrees with the oracle"""
    import ctypes

    L = ctypes.CDLL(emul_lib)
    L.qk_emul_counter.restype = ctypes.c_uint64
    for flags in (0, 1, _gflags(4)):
        before = (L.qk_emul_counter(0), L.qk_emul_counter(1))
        _check(emul_lib, H.serial_line(n_proc=6), 2, n_events=6000, log_capacity=65536, flags=flags)
        assert L.qk_emul_counter(0) - before[0] > 500, "busy-tick quick path never taken"
        assert L.qk_emul_counter(1) - before[1] > 20, "idle-and-blocked quick path never taken"


def test_scheduled_process_quantity(emul_lib):
    """create_scheduled_event!(SetProcessQuantity) on an F64Process (core.rs:1387-1391): the instance made at schedule
    time replaces process_quantity_distr at t and the process gets update_state(SCH_000000) -- both kernels, native and
    tape, against the oracle; and by hand on the all-Constant variant"""
    for flags in (0, 1, _gflags(4)):
        _check(emul_lib, H.scheduled_process_quantity(), 3, t_until=70 * 10**9, log_capacity=8192, flags=flags)
    _check(emul_lib, H.scheduled_process_quantity(), 2, t_until=70 * 10**9, tape_len=200, log_capacity=8192)
    # Constant variant: 1.0 per 1 s cycle until 20 s (the cycle started at 19 s still moves 1.0), 3.0 per cycle from the
    # cycle that starts at 20 s, 0.25 from the one that starts at 45 s
    m = H.scheduled_process_quantity(random=False)
    sim = m.sim(emul_lib, 1, log_capacity=8192)
    sim.step_until(70 * 10**9)
    log, _ = sim.read_log(0)
    from quokkasim_b200 import _capi as capi

    pid = m.by_name["Process"].component._id
    starts = [(int(r["time_ns"]) // 10**9, float(np.array([r["payload"]], dtype=np.uint64).view(np.float64)[0]))
              for r in log if r["comp"] == pid and r["kind"] == capi.LOG_VPROC_START]
    assert [q for t, q in starts if t < 20] == [1.0] * 20
    assert all(q == 3.0 for t, q in starts if 20 <= t < 45) and all(q == 0.25 for t, q in starts if t >= 45)
    assert {q for _, q in starts} == {1.0, 3.0, 0.25}
    # the scheduled update_state arrives with source SCH_000000 at exactly 20 s and 45 s
    sch = [int(r["time_ns"]) for r in log if r["comp"] == pid and r["src_comp"] == capi.SRC_SCH]
    assert sorted(set(sch)) == [20 * 10**9, 45 * 10**9]


def test_scheduled_event_arms_the_reference_does_not_have():
    """core.rs:1397-1399: every other (event, component) pair is the reference's Err string"""
    import quokkasim_b200 as qk

    m = H.scheduled_process_quantity()
    si = qk.BatchedSimInit.new()
    for c in m.components:
        si = qk.register_component(si, c)
    _, sched = si.init(0)
    df = qk.DistributionFactory(base_seed=0, next_seed=0)
    cfg = qk.DistributionConfig.Uniform(1.0, 2.0)
    for ev, target in ((qk.ScheduledEventConfig.SetProcessQuantity(cfg), m.by_name["Source"]),
                       (qk.ScheduledEventConfig.SetProcessTime(cfg), m.by_name["Process"]),
                       (qk.ScheduledEventConfig.SetMaxCapacity(5.0), m.by_name["S1"])):
        with pytest.raises(NotImplementedError, match="schedule_event not implemented for types"):
            qk.create_scheduled_event(sched, 10**9, ev, target.get_address(), df)
    # the factory is consulted at schedule time: two events in a row get consecutive seeds
    a = qk.create_scheduled_event(sched, 10**9, qk.ScheduledEventConfig.SetProcessQuantity(cfg), m.by_name["Process"].get_address(), df)
    b = qk.create_scheduled_event(sched, 2 * 10**9, qk.ScheduledEventConfig.SetProcessQuantity(cfg), m.by_name["Process"].get_address(), df)
    assert (a.stream, b.stream) == (0, 1)


# the PHASED variant of the lane-group kernel: pop + quick by lane groups, then one thread per replication runs the
# handlers of the whole block (here: blocks of 2-3 replications, a team of lanes x replications host threads)
@pytest.mark.parametrize("lanes", [2, 4, 8])
def test_phased_kernel_discrete_queue(emul_lib, lanes):
    s = _check(emul_lib, H.discrete_queue(), 5, n_events=1200, log_capacity=8192, flags=_gflags(lanes, True))
    assert s["n_events"] == 6000
    _check(emul_lib, H.discrete_queue(), 3, t_until=150 * 10**9, tape_len=200, log_capacity=2048, flags=_gflags(lanes, True))


def test_phased_kernel_models(emul_lib):
    f = _gflags(4, True)
    m = H.car_workshop()
    _check(emul_lib, m, 3, t_until=m.t0 + 9 * 3600 * 10**9, t0=m.t0, log_capacity=2048, flags=f)
    _check(emul_lib, H.haulage(), 3, n_events=5000, tape_len=3000, log_capacity=16384, flags=f)
    _check(emul_lib, H.haulage(n_src=3, per_queue=4), 2, n_events=8000, log_capacity=65536, flags=_gflags(8, True))
    _check(emul_lib, H.serial_line(), 3, n_events=6000, tape_len=3000, log_capacity=32768, flags=f)
    _check(emul_lib, H.serial_line(), 2, n_events=9000, log_capacity=65536, flags=_gflags(2, True))
    m = H.assembly_line()
    _check(emul_lib, m, 2, t_until=m.t0 + 24 * 3600 * 10**9, t0=m.t0, log_capacity=16384, flags=f)
    _check(emul_lib, H.vector_flow(3), 2, n_events=2500, log_capacity=32768, flags=f)
    _check(emul_lib, H.container_haulage(), 2, n_events=1500, log_capacity=32768, flags=f)
    _check(emul_lib, H.scheduled_process_quantity(), 3, t_until=70 * 10**9, log_capacity=8192, flags=f)


def test_phased_kernel_statistics_only_split_steps_and_faults(emul_lib):
    m = H.haulage()
    a = m.sim(emul_lib, 5, base_seed=5, log_capacity=0, flags=_gflags(4, True), threads_per_block=3)
    a.step_events(700)
    a.step_events(1)
    a.step_events_sliced(1299, 4)
    b = H.haulage().sim(emul_lib, 5, base_seed=5, log_capacity=16384)
    b.step_events(2000)
    for ci in range(len(m.components)):
        assert a.comp_stats(ci).tobytes() == b.comp_stats(ci).tobytes()
    assert a.status()[1].tobytes() == b.status()[1].tobytes()
    m = H.discrete_queue()
    tapes = H.make_tapes(m, 4, 300, seed=9)
    _, v0 = list(tapes.values())[0]
    v0[1, 3] = 0.0
    v0[2, 2] = -1.0
    sim = m.sim(emul_lib, 4, log_capacity=4096, flags=_gflags(4, True))
    for d, v in tapes.values():
        sim.set_tape(d, v)
    sim.step_events(600)
    assert list(sim.status()[0]) == [0, 1, 2, 0]

"""Kernel LOGIC on the CPU: the real kernel source (quokkasim_b200/csrc/qk_kernels.cuh) compiled with g++ under the
test-only shim in tests/emul, driven through the same C ABI and Python wrapper as the GPU library, compared bit
for bit with the oracle.  This keeps the CPU suite meaningful between GPU sessions; the parity tests proper are
tests/test_gpu_parity.py (-m gpu), which run the CUDA build of the same source."""
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

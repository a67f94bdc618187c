"""GPU parity tests: the CUDA path (through the C ABI of libquokka_b200.so) against the CPU oracle.

Bit-exact on: per-replication status, clock, executed-action count, the complete record stream
(time, component, kind, reason, seq, source id, payload), per-component statistics and stock contents.
"""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu


def _check(lib, model, n_reps, n_events=None, t_until=None, tape_len=None, t0=0, seed=42, sample=None,
           log_capacity=65536, rep_offset=0, threads_per_block=0, flags=0):
    om, dist_ids = H.lower_to_oracle(model)
    tapes = H.make_tapes(model, n_reps, tape_len, seed=3) if tape_len else None
    sim = model.sim(lib, n_reps, t0=t0, base_seed=seed, log_capacity=log_capacity, rep_offset=rep_offset,
                    threads_per_block=threads_per_block, flags=flags)
    if tapes:
        for d, v in tapes.values():
            sim.set_tape(d, v)
    if n_events is not None:
        sim.step_events(n_events)
    if t_until is not None:
        sim.step_until(t_until)
    reps = range(n_reps) if sample is None else sample
    for r in reps:
        osim = H.run_oracle(model, om, dist_ids, rep_offset + r if not tapes else r, t0=t0, base_seed=seed,
                            tapes=tapes, n_events=n_events, t_until=t_until)
        H.assert_rep_parity(sim, osim, model, r)
    s = sim.summary()
    sim.close()
    return s


def test_discrete_queue_tape_events(gpu_lib):
    """north_star parity clause: identical pre-drawn variate tape per replication -> identical trajectories."""
    s = _check(gpu_lib, H.discrete_queue(), 256, n_events=2000, tape_len=800, log_capacity=8192)
    assert s["n_events"] == 256 * 2000 and s["n_faulted"] == 0


def test_discrete_queue_reference_horizon(gpu_lib):
    """config C1: discrete_queue.rs exactly, step_until(200 s)."""
    s = _check(gpu_lib, H.discrete_queue(), 64, t_until=200 * 10**9, tape_len=200, log_capacity=2048)
    assert s["min_now_ns"] == s["max_now_ns"] == 200 * 10**9


def test_discrete_queue_native_philox(gpu_lib):
    """native Philox streams are specified operation by operation, so they are bit-exact too."""
    _check(gpu_lib, H.discrete_queue(), 4096, n_events=3000, sample=list(range(0, 4096, 131)), log_capacity=16384)


def test_rep_offset_is_the_philox_counter(gpu_lib):
    """a shard [offset, offset+n) reproduces the same replications as the full batch (multi-GPU contract)."""
    _check(gpu_lib, H.discrete_queue(), 96, n_events=1500, rep_offset=1000003, sample=[0, 1, 50, 95],
           log_capacity=8192)


def test_split_steps_equal_one_step(gpu_lib):
    """state survives the HBM round trip between launches: 3 launches == 1 launch."""
    m = H.discrete_queue()
    a = m.sim(gpu_lib, 128, base_seed=5, log_capacity=8192)
    a.step_events(700)
    a.step_events(1)
    a.step_events(799)
    b = H.discrete_queue().sim(gpu_lib, 128, base_seed=5, log_capacity=8192)
    b.step_events(1500)
    for r in (0, 17, 127):
        la, ta = a.read_log(r)
        lb, tb = b.read_log(r)
        assert ta == tb and la.tobytes() == lb.tobytes()
    for ci in range(len(m.components)):
        assert a.comp_stats(ci).tobytes() == b.comp_stats(ci).tobytes()
    a.close()
    b.close()


def test_car_workshop_ties_and_shared_stocks(gpu_lib):
    """car_workshop.rs: all-Constant times -> exact same-timestamp ties, shared stocks, stale keyed events (Q3)."""
    m = H.car_workshop()
    _check(gpu_lib, m, 32, t_until=m.t0 + 9 * 3600 * 10**9, t0=m.t0, log_capacity=4096)


def test_assembly_line_parallel_process_and_environment(gpu_lib):
    """assembly_line.rs: DiscreteParallelProcess + BasicEnvironment stop@1h / resume@2h, 3 days."""
    m = H.assembly_line()
    _check(gpu_lib, m, 64, t_until=m.t0 + 3 * 24 * 3600 * 10**9, t0=m.t0, sample=[0, 7, 63], log_capacity=16384)


def test_assembly_line_tape(gpu_lib):
    m = H.assembly_line()
    _check(gpu_lib, m, 32, t_until=m.t0 + 2 * 24 * 3600 * 10**9, t0=m.t0, tape_len=1500, sample=[0, 31],
           log_capacity=16384)


def test_serial_line_breakdowns(gpu_lib):
    """config C4 in miniature: DelayModes with Exponential until-delay / until-fix on every process."""
    _check(gpu_lib, H.serial_line(n_proc=4), 256, n_events=20000, sample=[0, 100, 255], log_capacity=131072)


def test_serial_line_32_processes(gpu_lib):
    """config C4 topology (32 DiscreteProcess / 33 DiscreteStock), fewer events."""
    _check(gpu_lib, H.serial_line(n_proc=32), 64, n_events=30000, sample=[0, 63], log_capacity=131072)


def test_haulage_network(gpu_lib):
    """config C5 in miniature: shared load/dump queues, several sources and sinks."""
    _check(gpu_lib, H.haulage(n_src=3, per_queue=4), 128, n_events=20000, sample=[0, 64, 127], log_capacity=131072)


def test_haulage_full_topology(gpu_lib):
    """config C5 topology: 8 Sources, 64 Processes, 16 shared stocks, 8 Sinks."""
    _check(gpu_lib, H.haulage(n_src=8, per_queue=8), 64, n_events=20000, sample=[0, 63], log_capacity=131072)


def test_faults_freeze_one_replication_only(gpu_lib):
    """tape value 0.0 -> Duration 0 -> "Time until next event is zero!" in that replication only."""
    m = H.discrete_queue()
    n = 64
    tapes = H.make_tapes(m, n, 300, seed=9)
    d0, v0 = list(tapes.values())[0]
    v0[5, 3] = 0.0  # zero duration -> QK_REP_ZERO_DURATION
    v0[9, 2] = -1.0  # negative -> QK_REP_BAD_DURATION
    om, dist_ids = H.lower_to_oracle(m)
    sim = m.sim(gpu_lib, n, log_capacity=4096)
    for d, v in tapes.values():
        sim.set_tape(d, v)
    sim.step_events(600)
    st, _, _ = sim.status()
    assert st[5] == 1 and st[9] == 2 and (np.delete(st, [5, 9]) == 0).all()
    for r in (4, 5, 9):
        osim = H.run_oracle(m, om, dist_ids, r, tapes=tapes, n_events=600)
        H.assert_rep_parity(sim, osim, m, r)
    sim.close()


def test_tape_exhaustion_is_reported(gpu_lib):
    m = H.discrete_queue()
    sim = m.sim(gpu_lib, 32, log_capacity=0)
    for d, v in H.make_tapes(m, 32, 10, seed=1).values():
        sim.set_tape(d, v)
    sim.step_events(5000)
    st, _, _ = sim.status()
    assert (st == 3).all()
    sim.close()


def test_block_size_does_not_change_results(gpu_lib):
    ref = None
    for nt in (32, 64, 128, 256):
        m = H.discrete_queue()
        sim = m.sim(gpu_lib, 300, base_seed=11, threads_per_block=nt)
        sim.step_events(4000)
        blob = b"".join(sim.comp_stats(ci).tobytes() for ci in range(len(m.components)))
        sim.close()
        if ref is None:
            ref = blob
        assert blob == ref


def test_reduce_stats_matches_per_replication_sums(gpu_lib):
    """the on-device summary reduction (what multi-GPU runs all-reduce) equals numpy over per-rep stats."""
    m = H.discrete_queue()
    n = 5000
    sim = m.sim(gpu_lib, n, base_seed=3)
    sim.step_events(3000)
    red = sim.reduce_stats()
    for ci in range(len(m.components)):
        st = sim.comp_stats(ci)
        t = [int(x) for x in st["time_ns"]]
        assert int(red[ci]["sum_count"]) == int(st["count"].sum())
        assert (int(red[ci]["sum_time_hi"]) << 64) + int(red[ci]["sum_time_lo"]) == sum(t)
        assert int(red[ci]["min_time"]) == min(t) and int(red[ci]["max_time"]) == max(t)
        assert int(red[ci]["min_count"]) == int(st["count"].min()) and int(red[ci]["max_count"]) == int(st["count"].max())
        assert int(red[ci]["hist_time"].sum()) == n and int(red[ci]["hist_count"].sum()) == n
        lo, hi = int(red[ci]["hist_lo_time"]), int(red[ci]["hist_hi_time"])
        w = (hi - lo) // 32 + 1
        expect = np.bincount([(x - lo) // w for x in t], minlength=32)
        assert (expect == red[ci]["hist_time"]).all()
        ts = np.array(t, dtype=np.float64) * 1e-9
        assert abs(red[ci]["sumsq_time"] - float((ts * ts).sum())) <= 1e-9 * max(1.0, float((ts * ts).sum()))
    s = sim.summary()
    assert s["n_events"] == n * 3000
    sim.close()


def test_invariants_at_scale(gpu_lib):
    """size-independent properties at 262,144 replications x 2,000 events (no oracle): item conservation,
    stock bounds, FIFO order of the items in every stock, monotone clocks."""
    m = H.discrete_queue()
    n = 262144
    sim = m.sim(gpu_lib, n, base_seed=2024)
    sim.step_events(2000)
    st, now, ev = sim.status()
    assert (st == 0).all() and (ev == 2000).all() and (now > 0).all()
    stats = [sim.comp_stats(ci) for ci in range(len(m.components))]
    src, q1, p1, q2, p2, q3, snk = stats
    # items created = items in the system + items sunk
    created = src["count"] + src["level"]
    in_system = q1["level"] + p1["level"] + q2["level"] + p2["level"] + q3["level"] + snk["level"] + src["level"]
    assert (created == in_system + snk["count"]).all()
    # each stock: adds - removes == level, where removes == downstream starts
    assert (q1["count"] - (p1["count"] + p1["level"]) == q1["level"]).all()
    assert (q2["count"] - (p2["count"] + p2["level"]) == q2["level"]).all()
    assert (q3["count"] - (snk["count"] + snk["level"]) == q3["level"]).all()
    # single-producer stocks never exceed max_capacity (10); busy time never exceeds elapsed time
    for q in (q1, q2, q3):
        assert (q["aux"] <= 10).all()
        assert (q["time_ns"] <= 10 * now).all()
    for p in (src, p1, p2, snk):
        assert (p["time_ns"] <= now).all()
    for r in (0, n // 2, n - 1):
        for ci in (1, 3, 5):
            items = sim.read_stock(ci, r)
            idx = [int(x) & 0x3FFFFFF for x in items]
            assert idx == sorted(idx)
    sim.close()


# ------------------------------------------------------------------------------------- continuous components
S = 10**9


def test_vector_examples(gpu_lib):
    """source_sink.rs, material_blending.rs, scheduled_event.rs (SetLowCapacity) -- fp64 stock levels bit-exact"""
    _check(gpu_lib, H.source_sink(), 32, t_until=120 * S, sample=[0, 31], log_capacity=8192)
    _check(gpu_lib, H.material_blending(), 32, t_until=3600 * S, sample=[0, 31], log_capacity=8192)
    _check(gpu_lib, H.scheduled_event_f64(), 32, t_until=120 * S, sample=[0, 31], log_capacity=2048)


def test_user_defined_resource_iron_ore(gpu_lib):
    """iron_ore.rs (five-field resource, total over two fields; qk_model_set_vector_n, feature-set-3 kernels): the example
    itself until 300 s under native streams and a tape, in every state placement; then the C3 topology carrying that
    resource through source, process with breakdowns, combiner, splitter, sink -- records and every field of every stock
    bit for bit, and the masses conserved field by field"""
    _check(gpu_lib, H.iron_ore(), 64, t_until=300 * S, sample=[0, 31, 63], log_capacity=2048)
    _check(gpu_lib, H.iron_ore(), 64, t_until=300 * S, tape_len=64, sample=[0, 63], log_capacity=2048)
    for flags in (1, 16):  # HBM planes ; split placement
        _check(gpu_lib, H.iron_ore(), 64, t_until=300 * S, sample=[0, 63], log_capacity=2048, flags=flags)
    _check(gpu_lib, H.vector_flow(5), 512, n_events=6000, sample=[0, 129, 511], log_capacity=65536)
    _check(gpu_lib, H.vector_flow(5), 64, n_events=2500, tape_len=2500, sample=[0, 63], log_capacity=65536, flags=16)
    m = H.iron_ore()
    sim = m.sim(gpu_lib, 4096, base_seed=9)
    sim.step_until(300 * S)
    init = [63.0, 42.0, 10.5, 5.25, 15.75]
    for r in (0, 1000, 4095):
        tot = [0.0] * 5
        for c in m.components:
            for k, x in enumerate(sim.read_resource(c.component._id, r)):
                tot[k] += x
        assert all(abs(t - i) <= 1e-9 * i for t, i in zip(tot, init)), (r, tot)
    sim.close()


@pytest.mark.parametrize("dim", [1, 3])
def test_vector_flow_native(gpu_lib, dim):
    """config C3 topology with native Philox streams, breakdowns and an environment stop/resume"""
    _check(gpu_lib, H.vector_flow(dim), 512, n_events=6000, sample=[0, 129, 511], log_capacity=65536)


def test_vector_flow_tape(gpu_lib):
    _check(gpu_lib, H.vector_flow(3), 64, n_events=2500, tape_len=2500, sample=[0, 63], log_capacity=32768)


def test_vector_flow_sink_delay(gpu_lib):
    _check(gpu_lib, H.vector_flow(3, sink_delay=True), 64, n_events=4000, sample=[0, 63], log_capacity=32768)


def test_vector_mass_balance_at_scale(gpu_lib):
    """config C3 size (262,144 replications), no oracle: initial + sourced == stocks + in-flight + sunk within
    1e-9 relative (BASELINE north_star tolerance for fp64 stock quantities)."""
    m = H.vector_flow(3)
    n = 262144
    sim = m.sim(gpu_lib, n, base_seed=99)
    sim.step_events(2000)
    st, now, ev = sim.status()
    assert (st == 0).all() and (ev == 2000).all()
    names = [c.component.element_name for c in m.components]
    stats = {nm: sim.comp_stats(i) for i, nm in enumerate(names)}
    initial = 9.0 + 15.0  # S1 [3,3,3] + S3 [10,0,5]
    stocks = sum(stats[k]["fvalue"] for k in ("S1", "S2", "S3", "S4", "S5"))
    inflight = sum(stats[k]["faux"] for k in ("Process", "Combiner", "Splitter", "Sink"))
    lhs = initial + stats["Source"]["fvalue"]
    rhs = stocks + inflight + stats["Sink"]["fvalue"]
    rel = np.abs(lhs - rhs) / np.maximum(1.0, np.abs(lhs))
    assert rel.max() < 1e-9, rel.max()
    assert (stats["Source"]["count"] > 0).all()
    sim.close()


def test_state_placement_does_not_change_results(gpu_lib):
    """shared-memory staging (flag 2) and HBM-resident state (flag 1) are two placements of the same state"""
    blobs = []
    for flags in (1, 2):
        m = H.haulage(n_src=3, per_queue=4)
        sim = m.sim(gpu_lib, 96, base_seed=21, flags=flags, log_capacity=1024)
        sim.step_events(5000)
        assert sim.info()["state_in_hbm"] == (1 if flags == 1 else 0)
        blob = b"".join(sim.comp_stats(ci).tobytes() for ci in range(len(m.components)))
        blob += sim.read_log(95)[0].tobytes()
        blobs.append(blob)
        sim.close()
    assert blobs[0] == blobs[1]
    # the default picks by residency: the 7-component line is staged on chip whole, one thread per replication; of the
    # 96-component network only the stocks' part of the scheduler queue is (split placement, two-tier queue: about 400 B per
    # replication instead of 3,924); the 67-component line with breakdowns keeps its whole queue on chip (split)
    a = H.discrete_queue().sim(gpu_lib, 64)
    b = H.haulage(n_src=8, per_queue=8).sim(gpu_lib, 64)
    c = H.serial_line(n_proc=32).sim(gpu_lib, 64)
    ia, ib, ic = a.info(), b.info(), c.info()
    assert ia["state_in_hbm"] == 0 and ia["lanes_per_replication"] == 1 and ia["state_split"] == 0
    assert ib["state_in_hbm"] == 0 and ib["lanes_per_replication"] == 1 and ib["state_split"] == 1
    assert ib["state_bytes_per_rep"] < 500 < ic["state_bytes_per_rep"] < 1000 and ic["state_split"] == 1
    c.close()
    a.close()
    b.close()


def test_tma_tile_staging_equals_row_copies(gpu_lib, monkeypatch):
    """The step kernel stages a block's hot planes as one 2-D TMA tile per plane (tensor maps made by qk_batch_create);
    QK_NO_TMA=1 makes the same library move them one bulk copy per slot row.  Same state, statistics and logs, bit for
    bit, across many launches (every launch stages in and out), in the on-chip and the split placements, with block
    sizes from 32 to 256 and a replication count that leaves the last block partly empty."""
    cases = [(H.discrete_queue, {}, 0, 0), (H.discrete_queue, {}, 0, 32), (H.discrete_queue, {}, 0, 256),
             (lambda: H.haulage(n_src=3, per_queue=4), {}, 16, 0), (lambda: H.vector_flow(3), {}, 16, 0),
             (lambda: H.vector_flow(5), {}, 0, 0), (lambda: H.serial_line(n_proc=8), {}, 48, 0)]
    for build, kw, flags, nt in cases:
        blobs = []
        for no_tma in (False, True):
            if no_tma:
                monkeypatch.setenv("QK_NO_TMA", "1")
            else:
                monkeypatch.delenv("QK_NO_TMA", raising=False)
            m = build()
            sim = m.sim(gpu_lib, 1000, base_seed=33, flags=flags, log_capacity=256, threads_per_block=nt)
            sim.step_events_chunked(600, 7)  # 86 launches
            sim.step_events(900)
            st, now, ev = sim.status()
            assert (ev == 1500).all() or (st != 0).any()
            blob = st.tobytes() + now.tobytes() + ev.tobytes()
            blob += b"".join(sim.comp_stats(ci).tobytes() for ci in range(len(m.components)))
            for r in (0, 511, 999):
                blob += sim.read_log(r)[0].tobytes()
            blobs.append(blob)
            sim.close()
        monkeypatch.delenv("QK_NO_TMA", raising=False)
        assert blobs[0] == blobs[1], (build, flags, nt)


# ------------------------------------------------------------------------------------- lane-group kernel (qk_group.cuh)
def _gflags(lanes, phased=False):
    # QK_BATCH_STATE_LANE_GROUP [| QK_BATCH_GROUP_PHASED] | lanes_per_replication << 8 (api.py convention)
    return 4 | (8 if phased else 0) | (lanes << 8)


@pytest.mark.parametrize("lanes", [4, 8, 16])
def test_group_kernel_matches_the_oracle(gpu_lib, lanes):
    """a lane group per replication: records, statistics, stocks bit-exact against the oracle; replication counts that
    do not fill the last block"""
    f = _gflags(lanes)
    _check(gpu_lib, H.haulage(n_src=8, per_queue=8), 101, n_events=20000, sample=[0, 50, 100], log_capacity=131072, flags=f)
    _check(gpu_lib, H.serial_line(n_proc=32), 70, n_events=30000, sample=[0, 69], log_capacity=131072, flags=f)
    _check(gpu_lib, H.discrete_queue(), 300, n_events=3000, tape_len=1200, sample=[0, 1, 150, 299], log_capacity=16384,
           flags=f)
    m = H.car_workshop()
    _check(gpu_lib, m, 33, t_until=m.t0 + 9 * 3600 * 10**9, t0=m.t0, sample=[0, 32], log_capacity=4096, flags=f)


def test_group_kernel_general_components(gpu_lib):
    f = _gflags(8)
    m = H.assembly_line()
    _check(gpu_lib, m, 64, t_until=m.t0 + 3 * 24 * 3600 * 10**9, t0=m.t0, sample=[0, 63], log_capacity=16384, flags=f)
    _check(gpu_lib, H.vector_flow(3), 200, n_events=6000, sample=[0, 99, 199], log_capacity=65536, flags=f)
    _check(gpu_lib, H.container_haulage(), 128, n_events=6000, sample=[0, 127], log_capacity=65536, flags=f)


def test_group_kernel_state_round_trip_and_placements_agree(gpu_lib):
    """the bulk-copy staging (one cp.async.bulk per slot row in, one out): a step cut into launches, and sliced into
    parts on side streams, equals one launch; and equals the thread-per-replication placements"""
    blobs = []
    for flags, how in ((_gflags(8), "one"), (_gflags(8), "pieces"), (_gflags(4), "sliced"), (_gflags(16), "chunked"),
                       (1, "one"), (2, "one")):
        m = H.haulage(n_src=3, per_queue=4)
        sim = m.sim(gpu_lib, 1000, base_seed=21, flags=flags, log_capacity=256)
        if how == "one":
            sim.step_events(6000)
        elif how == "pieces":
            sim.step_events(1)
            sim.step_events(2999)
            sim.step_events(3000)
        elif how == "sliced":
            sim.step_events_sliced(6000, 5)
        else:
            sim.step_events_chunked(6000, 700)
        blob = b"".join(sim.comp_stats(ci).tobytes() for ci in range(len(m.components)))
        logs, totals = sim.read_logs(0, 1000)
        blob += b"".join(x.tobytes() for x in logs) + totals.tobytes()
        for ci, c in enumerate(m.components):
            if c.variant == "StringStock":
                blob += sim.read_stock(ci, 999).tobytes()
        blobs.append(blob)
        sim.close()
    assert all(b == blobs[0] for b in blobs[1:])


def test_reference_vector_examples_transport_delay_testing_rocks(gpu_lib):
    """transport.rs, delay_testing.rs, the_example.rs at their own horizons"""
    _check(gpu_lib, H.transport(), 64, t_until=120 * S, sample=[0, 63], log_capacity=8192)
    _check(gpu_lib, H.delay_testing(), 32, t_until=40 * S, sample=[0], log_capacity=1024)
    m = H.the_example()
    _check(gpu_lib, m, 64, t_until=m.t0 + 700 * S, t0=m.t0, sample=[0, 31, 63], log_capacity=8192)


# ------------------------------------------------------------------------------------- container family (row f3)
def test_container_haulage_cycle(gpu_lib):
    """payload-true truck cycle: ContainerLoadingProcess / Vector3ContainerParallelProcess / ContainerUnloadingProcess /
    Vector3ContainerProcess over a fleet of initial Vector3Containers; bit-exact records, container payloads, stocks"""
    _check(gpu_lib, H.container_haulage(), 256, n_events=6000, sample=[0, 1, 131, 255], log_capacity=65536)
    _check(gpu_lib, H.container_haulage(n_trucks=7, loaders=3), 64, n_events=4000, tape_len=4000, sample=[0, 63],
           log_capacity=65536)
    _check(gpu_lib, H.container_haulage(random=False, breakdowns=False), 32, n_events=1500, sample=[0, 31],
           log_capacity=32768)


def test_container_open_line(gpu_lib):
    """Vector3ContainerSource (factory quantity distribution) -> Load -> Unload -> Sink, incl. the lost-resource quirk"""
    _check(gpu_lib, H.container_open(), 128, n_events=5000, sample=[0, 64, 127], log_capacity=65536)
    _check(gpu_lib, H.container_open(lossy_unload=False), 128, n_events=5000, sample=[0, 127], log_capacity=65536)


def test_container_state_in_shared_memory_and_in_hbm(gpu_lib):
    """both state placements of the feature-set-2 kernel give the same trajectories"""
    m = H.container_haulage()
    a = m.sim(gpu_lib, 96, base_seed=9, log_capacity=32768, flags=1)  # QK_BATCH_STATE_IN_HBM
    b = H.container_haulage().sim(gpu_lib, 96, base_seed=9, log_capacity=32768, flags=2)  # QK_BATCH_STATE_ON_CHIP
    assert a.info()["state_in_hbm"] == 1 and b.info()["state_in_hbm"] == 0
    a.step_events(3000)
    b.step_events(3000)
    for r in (0, 50, 95):
        assert a.read_log(r)[0].tobytes() == b.read_log(r)[0].tobytes()
    for ci in range(len(m.components)):
        assert a.comp_stats(ci).tobytes() == b.comp_stats(ci).tobytes()
    a.close()
    b.close()


def test_container_mass_balance_at_scale(gpu_lib):
    """65,536 replications of the truck cycle: ore + dumped + carried by the fleet stays constant (1e-9 relative)"""
    m = H.container_haulage()
    n = 65536
    sim = m.sim(gpu_lib, n, base_seed=77)
    sim.step_events(3000)
    st, _, ev = sim.status()
    assert not st.any() and (ev == 3000).all()
    from quokkasim_b200 import _capi as capi

    for r in (0, 1, 4097, n - 1):
        t = sum(sum(sim.read_vector(c.component._id, r)) for c in m.mass["vector_stocks"])
        fleet = 0
        for c in m.mass["container_holders"]:
            h, res = sim.read_containers(c.component._id, r)
            fleet += len(h)
            for row in res:
                if int(row[0]) != capi.CONTAINER_NONE_BITS:
                    t += float(row.view(np.float64).sum())
        assert fleet == 4
        assert abs(t - (100000.0 + 1.75 * 2)) <= 1e-9 * t, t
    sim.close()


def test_c2_full_size_invariants_and_sampled_oracle_parity(gpu_lib):
    """BASELINE config C2 at its full size -- 1,048,576 replications x 10,000 events, Philox seeds, the bench's own
    shape (newest 64 records kept per replication): conservation and bound invariants on EVERY replication, and
    bit-exact agreement with the CPU oracle (status, clock, statistics, stock contents, the retained log records) on a
    sample of replications spread over the batch."""
    m = H.discrete_queue()
    n, n_events, seed = 1048576, 10000, 1234
    sim = m.sim(gpu_lib, n, base_seed=seed, log_capacity=64)
    sim.step_events(n_events)
    st, now, ev = sim.status()
    assert (st == 0).all() and (ev == n_events).all()
    src, q1, p1, q2, p2, q3, snk = [sim.comp_stats(ci) for ci in range(len(m.components))]
    in_system = q1["level"] + p1["level"] + q2["level"] + p2["level"] + q3["level"] + snk["level"] + src["level"]
    assert (src["count"] + src["level"] == in_system + snk["count"]).all()
    assert (q1["count"] - (p1["count"] + p1["level"]) == q1["level"]).all()
    assert (q2["count"] - (p2["count"] + p2["level"]) == q2["level"]).all()
    assert (q3["count"] - (snk["count"] + snk["level"]) == q3["level"]).all()
    for q in (q1, q2, q3):
        assert (q["aux"] <= 10).all() and (q["time_ns"] <= 10 * now).all()
    for p in (src, p1, p2, snk):
        assert (p["time_ns"] <= now).all()
    summ = sim.summary()
    assert summ["n_events"] == n * n_events and summ["n_faulted"] == 0
    om, dist_ids = H.lower_to_oracle(m)
    for r in (0, 1, 63, 64, 4099, 524288, n - 2, n - 1):
        osim = H.run_oracle(m, om, dist_ids, r, base_seed=seed, n_events=n_events)
        H.assert_rep_parity(sim, osim, m, r, check_log=False)
        glog, total = sim.read_log(r)
        olog = osim.log()
        assert total == len(olog) and len(glog) == 64
        assert glog.tobytes() == olog[-64:].tobytes()
    sim.close()


def test_batch_that_does_not_fit_fails_cleanly(gpu_lib):
    """a batch larger than the GPU's memory: qk_batch_create reports the CUDA error, releases what it had allocated,
    and the device is as usable as before"""
    from quokkasim_b200 import _capi as capi

    m = H.discrete_queue()
    huge = m.sim(gpu_lib, 1 << 31, base_seed=1, log_capacity=64)  # ~ 1 TB of state + 4 TB of log rings
    with pytest.raises(capi.QkError) as e:
        huge.step_events(1)
    assert e.value.status == -5 and "cudaMalloc" in e.value.detail
    for _ in range(3):  # repeated failures do not eat memory either
        with pytest.raises(capi.QkError):
            H.discrete_queue().sim(gpu_lib, 1 << 31, log_capacity=64).step_events(1)
    ok = H.discrete_queue().sim(gpu_lib, 1 << 20, base_seed=1, log_capacity=64)
    ok.step_events(100)
    assert ok.summary()["n_events"] == 100 << 20
    ok.close()


def test_scheduled_process_quantity_on_gpu(gpu_lib):
    """create_scheduled_event!(SetProcessQuantity) on an F64Process (core.rs:1387-1391), every placement"""
    for flags in (0, 1, 2, _gflags(8)):
        _check(gpu_lib, H.scheduled_process_quantity(), 64, t_until=70 * 10**9, sample=[0, 63], log_capacity=8192,
               flags=flags)


def test_delay_mode_quick_path_at_scale(gpu_lib):
    """the quick handling of busy / idle components with delay modes: a longer serial line run, thread-per-replication
    (HBM planes) and lane groups, against the oracle"""
    for flags in (1, _gflags(8)):
        _check(gpu_lib, H.serial_line(n_proc=12), 200, n_events=60000, sample=[0, 199], log_capacity=262144, flags=flags)


@pytest.mark.parametrize("lanes", [4, 8, 16])
def test_phased_kernel_matches_the_oracle(gpu_lib, lanes):
    """the phased variant: pop + quick by lane groups, the handlers of a whole block side by side in one or two warps"""
    f = _gflags(lanes, True)
    _check(gpu_lib, H.haulage(n_src=8, per_queue=8), 101, n_events=20000, sample=[0, 50, 100], log_capacity=131072, flags=f)
    _check(gpu_lib, H.serial_line(n_proc=32), 70, n_events=30000, sample=[0, 69], log_capacity=131072, flags=f)
    _check(gpu_lib, H.discrete_queue(), 300, n_events=3000, tape_len=1200, sample=[0, 1, 150, 299], log_capacity=16384,
           flags=f)
    m = H.assembly_line()
    _check(gpu_lib, m, 64, t_until=m.t0 + 3 * 24 * 3600 * 10**9, t0=m.t0, sample=[0, 63], log_capacity=16384, flags=f)
    _check(gpu_lib, H.vector_flow(3), 200, n_events=6000, sample=[0, 99, 199], log_capacity=65536, flags=f)
    _check(gpu_lib, H.container_haulage(), 128, n_events=6000, sample=[0, 127], log_capacity=65536, flags=f)


@pytest.mark.parametrize("f", [16, 16 | 32])
@pytest.mark.parametrize("tpb", [0, 64, 128])
def test_split_placement_matches_the_oracle(gpu_lib, tpb, f):
    """QK_BATCH_STATE_SPLIT: the scheduler queue and the stocks' state words staged in shared memory, the component
    records in global memory -- every model family; with QK_BATCH_QUEUE_TWO_TIER the keyed events of the process-like
    components leave the chip too (cached minimum + busy mask)"""
    kw = dict(flags=f, threads_per_block=tpb)
    _check(gpu_lib, H.haulage(n_src=8, per_queue=8), 301, n_events=20000, sample=[0, 150, 300], log_capacity=131072, **kw)
    _check(gpu_lib, H.serial_line(n_proc=32), 170, n_events=30000, sample=[0, 169], log_capacity=131072, **kw)
    _check(gpu_lib, H.discrete_queue(), 300, n_events=3000, tape_len=1200, sample=[0, 1, 150, 299], log_capacity=16384, **kw)
    m = H.assembly_line()
    _check(gpu_lib, m, 64, t_until=m.t0 + 3 * 24 * 3600 * 10**9, t0=m.t0, sample=[0, 63], log_capacity=16384, **kw)
    _check(gpu_lib, H.vector_flow(3), 200, n_events=6000, sample=[0, 99, 199], log_capacity=65536, **kw)
    _check(gpu_lib, H.container_haulage(), 128, n_events=6000, sample=[0, 127], log_capacity=65536, **kw)


def test_split_placement_round_trips_and_agrees_with_the_other_placements(gpu_lib):
    blobs = []
    for flags, how in ((16, "one"), (16, "pieces"), (16, "sliced"), (48, "one"), (48, "pieces"), (48, "sliced"),
                       (_gflags(8), "one"), (1, "one")):
        m = H.haulage(n_src=3, per_queue=4)
        sim = m.sim(gpu_lib, 1000, base_seed=21, flags=flags, log_capacity=256)
        if how == "one":
            sim.step_events(6000)
        elif how == "pieces":
            sim.step_events(1)
            sim.step_events(2999)
            sim.step_events(3000)
        else:
            sim.step_events_sliced(6000, 5)
        blob = b"".join(sim.comp_stats(ci).tobytes() for ci in range(len(m.components)))
        logs, totals = sim.read_logs(0, 1000)
        blob += b"".join(x.tobytes() for x in logs) + totals.tobytes()
        blobs.append(blob)
        sim.close()
    assert all(b == blobs[0] for b in blobs[1:])


def test_phased_kernel_round_trips_and_agrees_with_the_other_placements(gpu_lib):
    blobs = []
    for flags, how in ((_gflags(8, True), "one"), (_gflags(8, True), "pieces"), (_gflags(4, True), "sliced"),
                       (_gflags(8), "one"), (1, "one")):
        m = H.haulage(n_src=3, per_queue=4)
        sim = m.sim(gpu_lib, 1000, base_seed=21, flags=flags, log_capacity=256)
        if how == "one":
            sim.step_events(6000)
        elif how == "pieces":
            sim.step_events(1)
            sim.step_events(2999)
            sim.step_events(3000)
        else:
            sim.step_events_sliced(6000, 5)
        blob = b"".join(sim.comp_stats(ci).tobytes() for ci in range(len(m.components)))
        logs, totals = sim.read_logs(0, 1000)
        blob += b"".join(x.tobytes() for x in logs) + totals.tobytes()
        blobs.append(blob)
        sim.close()
    assert all(b == blobs[0] for b in blobs[1:])

"""BASELINE configs C3, C4 and C5 at their stated sizes: size-independent invariants on EVERY
replication plus bit-exact agreement with the CPU oracle on a sample of replications spread over the batch.  (C2 at
full size is tests/test_gpu_parity.py::test_c2_full_size_invariants_and_sampled_oracle_parity.)  Also: the wave-sliced
launch schedule, the batched log read-back and the summary reduction against their plain counterparts on the GPU."""
import os

import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu


def _discrete_flow_invariants(m, sim, n_events):
    """item conservation in a network of discrete stocks and single-server processes, for every replication:
    what the producers of a stock finished is what was added to it; what was added (+ what it started with) minus what
    its consumers started is what it holds; nothing is busy for longer than the clock has run."""
    st, now, ev = sim.status()
    assert (st == 0).all() and (ev == n_events).all() and (now > 0).all()
    comps = [c.component for c in m.components]
    stats = [sim.comp_stats(i) for i in range(len(comps))]
    conns = [x for c in comps for x in c._conns]
    for i, c in enumerate(comps):
        if m.components[i].variant != "StringStock":
            assert (stats[i]["time_ns"] <= now).all(), c.element_name
            assert (stats[i]["aux"] <= now).all(), c.element_name
            continue
        producers = [a._id for _, a, b, _ in conns if b is c]
        consumers = [b._id for _, a, b, _ in conns if a is c]
        added = sum(stats[p]["count"] for p in producers)
        started = sum(stats[q]["count"] + stats[q]["level"] for q in consumers)
        assert (stats[i]["count"] == added).all(), c.element_name
        assert (stats[i]["count"] + len(c.resource) - started == stats[i]["level"]).all(), c.element_name
        assert (stats[i]["aux"] <= c.max_capacity + len(producers)).all(), c.element_name
    return stats, now


def _sampled_oracle(m, sim, seed, n_events, reps, log_tail):
    om, dist_ids = H.lower_to_oracle(m)
    for r in reps:
        osim = H.run_oracle(m, om, dist_ids, r, base_seed=seed, n_events=n_events, log=log_tail > 0)
        H.assert_rep_parity(sim, osim, m, r, check_log=False)
        if log_tail:
            glog, total = sim.read_log(r)
            olog = osim.log()
            assert total == len(olog) and len(glog) == min(log_tail, total)
            assert glog.tobytes() == olog[-len(glog):].tobytes()


def test_c3_vector_flow_full_size(gpu_lib):
    """C3: continuous / vector stock-and-flow model, 262,144 replications x 10,000 events: mass balance on every
    replication (1e-9 relative, the north-star tolerance) and bit-exact fp64 against the oracle on 8 of them."""
    m = H.vector_flow(3)
    n, n_events, seed = 262144, 10000, 1234
    sim = m.sim(gpu_lib, n, base_seed=seed, log_capacity=64)
    sim.step_events(n_events)
    st, now, ev = sim.status()
    assert (st == 0).all() and (ev == n_events).all()
    names = [c.component.element_name for c in m.components]
    stats = {nm: sim.comp_stats(i) for i, nm in enumerate(names)}
    initial = 9.0 + 15.0  # S1 [3,3,3] + S3 [10,0,5]
    stocks = sum(stats[k]["fvalue"] for k in ("S1", "S2", "S3", "S4", "S5"))
    inflight = sum(stats[k]["faux"] for k in ("Process", "Combiner", "Splitter", "Sink"))
    lhs = initial + stats["Source"]["fvalue"]
    rhs = stocks + inflight + stats["Sink"]["fvalue"]
    rel = np.abs(lhs - rhs) / np.maximum(1.0, np.abs(lhs))
    assert rel.max() < 1e-9, rel.max()
    assert (stats["Source"]["count"] > 0).all() and (stats["Sink"]["count"] > 0).all()
    _sampled_oracle(m, sim, seed, n_events, (0, 1, 63, 64, 4099, n // 2, n - 2, n - 1), 64)
    sim.close()


def test_c5_haulage_full_size(gpu_lib):
    """C5 (one point of the sweep): 8 Sources, 64 Processes, 16 shared stocks, 8 Sinks, 262,144 replications x
    10,000 events."""
    m = H.haulage(n_src=8, per_queue=8)
    n, n_events, seed = 262144, 10000, 1234
    sim = m.sim(gpu_lib, n, base_seed=seed, log_capacity=64)
    sim.step_events(n_events)
    _discrete_flow_invariants(m, sim, n_events)
    _sampled_oracle(m, sim, seed, n_events, (0, 1, 63, 64, 4099, n // 2, n - 2, n - 1), 64)
    sim.close()


@pytest.mark.skipif(os.environ.get("QK_SKIP_C4_FULL") == "1", reason="C4 at full size takes about a minute of GPU time")
def test_c4_serial_line_full_size(gpu_lib):
    """C4: 32 Processes / 33 Stocks with random breakdowns, 100,000 replications x 1,000,000 events (1e11 events),
    statistics only (a log ring would only keep the last 64 of ~3e6 records per replication -- checked too)."""
    m = H.serial_line(n_proc=32)
    n, n_events, seed = 100000, 1000000, 1234
    sim = m.sim(gpu_lib, n, base_seed=seed, log_capacity=64)
    sim.step_events(n_events)
    stats, now = _discrete_flow_invariants(m, sim, n_events)
    # every process broke down at some point of ~1e6 simulated seconds, and spent less time broken than running
    for i, c in enumerate(m.components):
        if c.variant == "StringProcess":
            assert (stats[i]["aux"] > 0).mean() > 0.999 and (stats[i]["aux"] < now).all()
    _sampled_oracle(m, sim, seed, n_events, (0, 1, 63, 64, 4099, n // 2, n - 2, n - 1), 64)
    sim.close()


def test_wave_sliced_schedule_equals_plain_launch(gpu_lib):
    """more tiles than fit the GPU at once: (block, chunk) work items as wave-sized launches == one launch"""
    n = 150000
    a = H.discrete_queue().sim(gpu_lib, n, base_seed=5, log_capacity=64)
    info = a.info()
    assert info["grid_blocks"] > info["wave_blocks"] > 0
    a.step_events_chunked(3000, 3000)  # one plain launch over all blocks
    m = H.discrete_queue()
    for chunks in (3, 7):
        b = m.sim(gpu_lib, n, base_seed=5, log_capacity=64)
        b.step_events_sliced(3000, chunks)
        assert b.info()["last_step_chunks"] == chunks
        assert (b.status()[2] == 3000).all()
        for ci in range(len(m.components)):
            assert a.comp_stats(ci).tobytes() == b.comp_stats(ci).tobytes()
        la, ta = a.read_logs(n - 300, 300)
        lb, tb = b.read_logs(n - 300, 300)
        assert (ta == tb).all() and all(x.tobytes() == y.tobytes() for x, y in zip(la, lb))
        b.close()
    a.close()


def test_read_logs_gather_equals_read_log(gpu_lib):
    m = H.discrete_queue()
    s = m.sim(gpu_lib, 3000, base_seed=3, log_capacity=64)
    s.step_events(40)   # some rings not yet full
    logs, totals = s.read_logs(100, 2500)
    for i in (0, 1, 777, 2499):
        one, tot = s.read_log(100 + i)
        assert tot == totals[i] and logs[i].tobytes() == one.tobytes()
    s.step_events(600)  # wrapped rings
    logs, totals = s.read_logs(0, 3000)
    for i in (0, 1500, 2999):
        one, tot = s.read_log(i)
        assert tot == totals[i] and len(one) == 64 and logs[i].tobytes() == one.tobytes()
    s.close()


def test_reduction_moments_are_exact(gpu_lib):
    """the packed partials (what crosses the GPUs) against Python integers over the per-replication statistics"""
    from quokkasim_b200 import parallel

    m = H.haulage(n_src=3, per_queue=4)
    n = 7001
    sim = m.sim(gpu_lib, n, base_seed=8)
    sim.step_events(2500)
    part = sim.reduce_partials()
    comp, n_ev, n_fault, n_reps = parallel.summarize(part, None, sim.n_comp)
    red = sim.reduce_stats()
    assert n_ev == n * 2500 and n_fault == 0 and n_reps == n
    for ci in range(sim.n_comp):
        st = sim.comp_stats(ci)
        t = [int(x) for x in st["time_ns"]]
        c = [int(x) for x in st["count"]]
        lv = [int(x) for x in st["level"]]
        g = comp[ci]
        assert g["sum_time_ns"] == sum(t) and g["sumsq_time_ns"] == sum(x * x for x in t)
        assert g["sum_count"] == sum(c) and g["sumsq_count"] == sum(x * x for x in c)
        assert g["sum_level"] == sum(lv) and g["sumsq_level"] == sum(x * x for x in lv)
        assert (g["min_time_ns"], g["max_time_ns"], g["min_count"], g["max_count"]) == (min(t), max(t), min(c), max(c))
        assert sum(int(red[ci]["sumsq_time_ns2"][k]) << (64 * k) for k in range(3)) == g["sumsq_time_ns"]
        assert (int(red[ci]["sumsq_count_hi"]) << 64) + int(red[ci]["sumsq_count_lo"]) == g["sumsq_count"]
    sim.close()

"""N > 1 host logic on CPU: world_size 2, gloo backend.

Each rank advances its contiguous replication shard (rep_offset = Philox counter range) with the host build of
the kernel source (tests/emul), forms the same partial arrays the CUDA reduction kernel produces, all-reduces them
(sum / min / max) and the result must equal the single-batch answer.  No data-path collective exists on this path:
the only exchange is this final reduction of summary statistics."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_TOTAL, N_EVENTS, SEED = 24, 800, 321


def _partials(sim, n_comp):
    """what qk_reduce_kernel writes (quokka_b200.h QP_*), formed in numpy from the per-replication statistics --
    independently of the library's own qk_batch_reduce_partials, which must agree with it"""
    from quokkasim_b200 import _capi as capi

    w = np.zeros((n_comp, capi.PARTIAL_WORDS), dtype=np.uint64)
    st, _, ev = sim.status()
    m32 = np.uint64(0xFFFFFFFF)
    s32 = np.uint64(32)
    for c in range(n_comp):
        s = sim.comp_stats(c)
        t, cnt, lvl = s["time_ns"], s["count"], s["level"]
        lo, hi = t & m32, t >> s32
        w[c, capi.QP_COUNT] = cnt.sum()
        w[c, capi.QP_TIME_LO] = lo.sum()
        w[c, capi.QP_TIME_HI] = hi.sum()
        w[c, capi.QP_LEVEL] = lvl.sum()
        w[c, capi.QP_MIN_COUNT], w[c, capi.QP_MIN_TIME] = cnt.min(), t.min()
        w[c, capi.QP_NMAX_COUNT], w[c, capi.QP_NMAX_TIME] = ~cnt.max(), ~t.max()
        for k, v in ((capi.QP_AA_LO, hi * hi), (capi.QP_AB_LO, hi * lo), (capi.QP_BB_LO, lo * lo),
                     (capi.QP_CC_LO, cnt * cnt)):
            w[c, k] = (v & m32).sum()
            w[c, k + 1] = (v >> s32).sum()
        w[c, capi.QP_LEVEL_SQ] = (lvl * lvl).sum()
    w[0, capi.QP_NEVENTS] = ev.sum()
    w[0, capi.QP_NFAULTED] = (st != 0).sum()
    w[0, capi.QP_NREPS] = len(st)
    assert sim.reduce_partials().tobytes() == w.tobytes()
    return w


def _worker(rank, world, port, emul_lib, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from quokkasim_b200 import _capi as capi
    from quokkasim_b200 import parallel, workloads

    lo, hi = parallel.shard_range(N_TOTAL, rank, world)
    m = workloads.discrete_queue()
    sim = m.sim(emul_lib, hi - lo, base_seed=SEED, rep_offset=lo)
    sim.step_events(N_EVENTS)
    n_comp = len(m.components)
    mine = torch.from_numpy(_partials(sim, n_comp).view(np.int64).reshape(-1).copy())
    # the one exchange of the multi-GPU path: all-gather of the packed partials, combined locally
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    combined = parallel.combine_partials([g.numpy().view(np.uint64).reshape(n_comp, capi.PARTIAL_WORDS)
                                          for g in gathered])
    # histograms over the GLOBAL ranges, then a sum all-reduce
    hist = np.zeros((n_comp, 2, capi.HIST_BINS), dtype=np.int64)
    for c in range(n_comp):
        s = sim.comp_stats(c)
        for k, (key, lo_w, hi_w) in enumerate((("time_ns", capi.QP_MIN_TIME, capi.QP_NMAX_TIME),
                                               ("count", capi.QP_MIN_COUNT, capi.QP_NMAX_COUNT))):
            lo_v, hi_v = int(combined[c, lo_w]), int(combined[c, hi_w]) ^ ((1 << 64) - 1)
            width = (hi_v - lo_v) // capi.HIST_BINS + 1
            for v in s[key]:
                hist[c, k, (int(v) - lo_v) // width] += 1
    t_hist = torch.from_numpy(hist)
    dist.all_reduce(t_hist, op=dist.ReduceOp.SUM)
    comp, n_ev, n_fault, n_reps = parallel.summarize(combined, t_hist.numpy().view(np.uint64), n_comp)
    if rank == 0:
        torch.save({"comp": comp, "n_ev": n_ev, "n_fault": n_fault, "n_reps": n_reps}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_tile_the_batch():
    from quokkasim_b200 import parallel

    for n, w in ((1048576, 8), (10, 3), (7, 8), (100000, 4)):
        r = [parallel.shard_range(n, k, w) for k in range(w)]
        assert r[0][0] == 0 and r[-1][1] == n
        assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
        assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


def test_two_rank_reduction_equals_single_batch(emul_lib, tmp_path):
    out = str(tmp_path / "r0.pt")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, emul_lib, out), nprocs=2, join=True)
    got = torch.load(out, weights_only=False)
    sys.path.insert(0, ROOT)
    from quokkasim_b200 import _capi as capi
    from quokkasim_b200 import parallel, workloads

    m = workloads.discrete_queue()
    n_comp = len(m.components)
    sim = m.sim(emul_lib, N_TOTAL, base_seed=SEED)
    sim.step_events(N_EVENTS)
    # the single-batch answer, from the library's own summary (exact integers: moments and histograms included)
    red = sim.reduce_stats()
    assert got["n_ev"] == N_TOTAL * N_EVENTS and got["n_fault"] == 0 and got["n_reps"] == N_TOTAL
    for c in range(n_comp):
        g, r = got["comp"][c], red[c]
        assert g["sum_count"] == int(r["sum_count"]) and g["sum_level"] == int(r["sum_level"])
        assert g["sum_time_ns"] == (int(r["sum_time_hi"]) << 64) + int(r["sum_time_lo"])
        assert (g["min_count"], g["max_count"]) == (int(r["min_count"]), int(r["max_count"]))
        assert (g["min_time_ns"], g["max_time_ns"]) == (int(r["min_time"]), int(r["max_time"]))
        assert g["sumsq_time_ns"] == sum(int(r["sumsq_time_ns2"][k]) << (64 * k) for k in range(3))
        assert g["sumsq_count"] == (int(r["sumsq_count_hi"]) << 64) + int(r["sumsq_count_lo"])
        assert g["sumsq_level"] == int(r["sumsq_level"])
        assert g["hist_time"] == r["hist_time"].tolist() and g["hist_count"] == r["hist_count"].tolist()
        # and against the per-replication values themselves
        s = sim.comp_stats(c)
        assert g["sumsq_time_ns"] == sum(int(t) ** 2 for t in s["time_ns"])
        assert g["sumsq_count"] == sum(int(t) ** 2 for t in s["count"])
        mom = parallel.moments(g, N_TOTAL)
        ts = s["time_ns"].astype(np.float64)
        assert abs(mom["time_ns"][0] - ts.mean()) <= 1e-9 * max(1.0, ts.mean())
        assert abs(mom["time_ns"][1] - ts.var()) <= 1e-6 * max(1.0, ts.var())
    # the host-side statement of the combine agrees with what the ranks exchanged
    parts = []
    for r in range(2):
        lo, hi = parallel.shard_range(N_TOTAL, r, 2)
        s = workloads.discrete_queue().sim(emul_lib, hi - lo, base_seed=SEED, rep_offset=lo)
        s.step_events(N_EVENTS)
        parts.append(_partials(s, n_comp))
    whole = _partials(sim, n_comp)
    assert parallel.combine_partials(parts).tobytes() == whole.tobytes()
    assert parallel.summarize(whole, None, n_comp)[0][3]["sum_count"] == got["comp"][3]["sum_count"]

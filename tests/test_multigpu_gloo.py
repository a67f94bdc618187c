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
    """what qk_reduce_kernel writes: sums[c] = {count, time_lo32, time_hi32, level, n_events, n_faulted}"""
    sums = np.zeros((n_comp, 6), dtype=np.uint64)
    mins = np.zeros((n_comp, 2), dtype=np.uint64)
    maxs = np.zeros((n_comp, 2), dtype=np.uint64)
    st, _, ev = sim.status()
    for c in range(n_comp):
        s = sim.comp_stats(c)
        sums[c, 0] = s["count"].sum()
        sums[c, 1] = (s["time_ns"] & np.uint64(0xFFFFFFFF)).sum()
        sums[c, 2] = (s["time_ns"] >> np.uint64(32)).sum()
        sums[c, 3] = s["level"].sum()
        mins[c] = (s["count"].min(), s["time_ns"].min())
        maxs[c] = (s["count"].max(), s["time_ns"].max())
    sums[0, 4] = ev.sum()
    sums[0, 5] = (st != 0).sum()
    return sums, mins, maxs


def _worker(rank, world, port, emul_lib, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from quokkasim_b200 import parallel, workloads

    lo, hi = parallel.shard_range(N_TOTAL, rank, world)
    m = workloads.discrete_queue()
    sim = m.sim(emul_lib, hi - lo, base_seed=SEED, rep_offset=lo)
    sim.step_events(N_EVENTS)
    n_comp = len(m.components)
    sums, mins, maxs = _partials(sim, n_comp)
    t_sums = torch.from_numpy(sums.view(np.int64).copy())
    t_mins = torch.from_numpy(mins.view(np.int64).copy())
    t_maxs = torch.from_numpy(maxs.view(np.int64).copy())
    dist.all_reduce(t_sums, op=dist.ReduceOp.SUM)
    dist.all_reduce(t_mins, op=dist.ReduceOp.MIN)
    dist.all_reduce(t_maxs, op=dist.ReduceOp.MAX)
    comp, n_ev, n_fault = parallel.summarize(t_sums.numpy().view(np.uint64), t_mins.numpy().view(np.uint64),
                                             t_maxs.numpy().view(np.uint64), None, n_comp)
    if rank == 0:
        torch.save({"comp": comp, "n_ev": n_ev, "n_fault": n_fault}, out)
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
    from quokkasim_b200 import parallel, workloads

    m = workloads.discrete_queue()
    sim = m.sim(emul_lib, N_TOTAL, base_seed=SEED)
    sim.step_events(N_EVENTS)
    sums, mins, maxs = _partials(sim, len(m.components))
    want, n_ev, n_fault = parallel.summarize(sums, mins, maxs, None, len(m.components))
    assert got["n_ev"] == n_ev == N_TOTAL * N_EVENTS and got["n_fault"] == n_fault == 0
    assert got["comp"] == want
    # and the host-side statement of the collective agrees with gloo
    lo0, hi0 = parallel.shard_range(N_TOTAL, 0, 2)
    parts = []
    for lo, hi in ((lo0, hi0), parallel.shard_range(N_TOTAL, 1, 2)):
        s = workloads.discrete_queue().sim(emul_lib, hi - lo, base_seed=SEED, rep_offset=lo)
        s.step_events(N_EVENTS)
        parts.append(_partials(s, len(m.components)))
    c_sums, c_mins, c_maxs = parallel.combine_partials([p[0] for p in parts], [p[1] for p in parts],
                                                       [p[2] for p in parts])
    assert parallel.summarize(c_sums, c_mins, c_maxs, None, len(m.components))[0] == want

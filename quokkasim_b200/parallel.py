"""Multi-GPU plumbing: one process per GPU, replications sharded by contiguous seed range, and ONE exchange --
the final reduction of summary statistics (NCCL over NVLink through torch.distributed).

The simulation itself never communicates: replications are independent, so there is no data-path
collective to fuse with the kernel.  PyTorch is used here only for device buffers, streams and the process group.
(A host without Python uses qk_multi_* of the C ABI, which does the same inside the library with NCCL.)

What crosses the GPUs is each rank's array of packed partials ([n_comp][PARTIAL_WORDS] exact u64 words, quokka_b200.h
QP_*): one all-gather, combined locally (unsigned min on the extrema words, wrapping sum elsewhere -- the second
moments of busy time / occupancy integral, counts and levels included); then every rank bins its replications over the
now global ranges and one sum all-reduce adds the histograms.
"""
import numpy as np

from . import _capi as capi

_MIN_WORDS = (capi.QP_MIN_COUNT, capi.QP_MIN_TIME, capi.QP_NMAX_COUNT, capi.QP_NMAX_TIME)
_M64 = (1 << 64) - 1


def shard_range(n_total, rank, world_size):
    """Contiguous replication range [lo, hi) owned by `rank` (= its Philox counter range)."""
    lo = (n_total * rank) // world_size
    hi = (n_total * (rank + 1)) // world_size
    return lo, hi


def combine_partials(parts):
    """Host-side statement of qk_batch_combine_partials_device (used by the gloo CPU tests and for checking the
    device path): `parts` = list of (n_comp, PARTIAL_WORDS) uint64 arrays."""
    stack = np.stack([np.asarray(p, dtype=np.uint64) for p in parts])
    out = np.sum(stack, axis=0, dtype=np.uint64)  # wraps modulo 2^64 like the device sum
    for k in _MIN_WORDS:
        out[:, k] = stack[:, :, k].min(axis=0)
    return out


def summarize(partials, hist, n_comp):
    """Combined partials (+ histograms) -> (list of per-component dicts, total events, faulted replications,
    replications).  Every value is an exact Python int; the moments are sums over replications."""
    w = np.asarray(partials, dtype=np.uint64).reshape(n_comp, capi.PARTIAL_WORDS)
    hist = None if hist is None else np.asarray(hist, dtype=np.uint64).reshape(n_comp, 2, capi.HIST_BINS)
    out = []
    for c in range(n_comp):
        g = lambda k: int(w[c, k])  # noqa: E731
        two = lambda lo, hi: (g(hi) << 32) + g(lo)  # noqa: E731
        out.append({
            "sum_count": g(capi.QP_COUNT),
            "sum_time_ns": two(capi.QP_TIME_LO, capi.QP_TIME_HI),
            "sum_level": g(capi.QP_LEVEL),
            "min_count": g(capi.QP_MIN_COUNT), "min_time_ns": g(capi.QP_MIN_TIME),
            "max_count": g(capi.QP_NMAX_COUNT) ^ _M64, "max_time_ns": g(capi.QP_NMAX_TIME) ^ _M64,
            "sumsq_time_ns": (two(capi.QP_AA_LO, capi.QP_AA_HI) << 64) + (two(capi.QP_AB_LO, capi.QP_AB_HI) << 33)
                             + two(capi.QP_BB_LO, capi.QP_BB_HI),
            "sumsq_count": two(capi.QP_CC_LO, capi.QP_CC_HI),
            "sumsq_level": g(capi.QP_LEVEL_SQ),
            "hist_time": None if hist is None else hist[c, 0].tolist(),
            "hist_count": None if hist is None else hist[c, 1].tolist(),
        })
    return out, int(w[0, capi.QP_NEVENTS]), int(w[0, capi.QP_NFAULTED]), int(w[0, capi.QP_NREPS])


def moments(summary, n_reps, horizon_ns=None):
    """Mean and (population) variance over replications of a component's count / time_ns / level from the exact sums."""
    from fractions import Fraction as F

    out = {}
    for name, s1, s2 in (("count", "sum_count", "sumsq_count"), ("time_ns", "sum_time_ns", "sumsq_time_ns"),
                         ("level", "sum_level", "sumsq_level")):
        mean = F(summary[s1], n_reps)
        var = F(summary[s2], n_reps) - mean * mean
        out[name] = (float(mean), float(var))
    if horizon_ns:
        out["utilisation"] = out["time_ns"][0] / horizon_ns
    return out


def _ordered(sim):
    """The library enqueues on the batch's stream, torch on its current stream: unless they are the same stream the
    two have to be ordered by hand around every hand-over."""
    import torch

    same = sim.stream is not None and int(sim.stream) == int(torch.cuda.current_stream().cuda_stream)
    return same


def allreduce_summary(sim, device, group=None, timings=None):
    """Per-GPU reduction kernel -> all-gather of the packed partials -> local combine -> histograms over the global
    ranges -> sum all-reduce -> ONE device-to-host copy.

    Returns (per-component summaries, total events, total faulted replications) identical on every rank."""
    import torch
    import torch.distributed as dist

    n = sim.n_comp
    W = capi.PARTIAL_WORDS
    distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    world = dist.get_world_size(group) if distributed else 1
    same_stream = _ordered(sim)
    mine = torch.empty(n * W, dtype=torch.int64, device=device)
    both = torch.empty(n * W + n * 2 * capi.HIST_BINS, dtype=torch.int64, device=device)
    combined, hist = both[: n * W], both[n * W:]
    if not same_stream:
        torch.cuda.current_stream().synchronize()
    sim.reduce_partials_device(mine.data_ptr())
    if distributed:
        gathered = torch.empty(world * n * W, dtype=torch.int64, device=device)
        if not same_stream:
            sim.synchronize()
        dist.all_gather_into_tensor(gathered, mine, group=group)
        if not same_stream:
            torch.cuda.current_stream().synchronize()
    else:
        gathered = mine
    sim.combine_partials_device(gathered.data_ptr(), world, combined.data_ptr())
    sim.histograms_device(combined.data_ptr(), hist.data_ptr())
    if not same_stream:
        sim.synchronize()
    if distributed:
        dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=group)
    host = both.cpu().numpy().view(np.uint64)
    comp, n_ev, n_fault, _ = summarize(host[: n * W], host[n * W:], n)
    return comp, n_ev, n_fault

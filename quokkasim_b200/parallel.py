"""Multi-GPU plumbing: one process per GPU, replications sharded by contiguous seed range, and ONE collective --
the final reduction of summary statistics (NCCL over NVLink through torch.distributed).

The simulation itself never communicates: replications are independent, so there is no data-path
collective to fuse with the kernel.  PyTorch is used here only for device buffers, streams and the process group.
"""
import numpy as np

from . import _capi as capi


def shard_range(n_total, rank, world_size):
    """Contiguous replication range [lo, hi) owned by `rank` (= its Philox counter range)."""
    lo = (n_total * rank) // world_size
    hi = (n_total * (rank + 1)) // world_size
    return lo, hi


def combine_partials(sums_list, mins_list, maxs_list):
    """Host-side statement of what the collective computes (used by the gloo CPU tests and for N == 1):
    sums add, minima take min, maxima take max."""
    sums = np.sum(np.stack(sums_list), axis=0, dtype=np.uint64)
    mins = np.min(np.stack(mins_list), axis=0)
    maxs = np.max(np.stack(maxs_list), axis=0)
    return sums, mins, maxs


def summarize(sums, mins, maxs, hist, n_comp):
    """Arrays -> list of per-component dicts (exact Python ints)."""
    out = []
    sums = np.asarray(sums, dtype=np.uint64).reshape(n_comp, 6)
    mins = np.asarray(mins, dtype=np.uint64).reshape(n_comp, 2)
    maxs = np.asarray(maxs, dtype=np.uint64).reshape(n_comp, 2)
    hist = None if hist is None else np.asarray(hist, dtype=np.uint64).reshape(n_comp, 2, capi.HIST_BINS)
    for c in range(n_comp):
        out.append({
            "sum_count": int(sums[c, 0]),
            "sum_time_ns": (int(sums[c, 2]) << 32) + int(sums[c, 1]),
            "sum_level": int(sums[c, 3]),
            "min_count": int(mins[c, 0]), "min_time_ns": int(mins[c, 1]),
            "max_count": int(maxs[c, 0]), "max_time_ns": int(maxs[c, 1]),
            "hist_time": None if hist is None else hist[c, 0].tolist(),
            "hist_count": None if hist is None else hist[c, 1].tolist(),
        })
    return out, int(sums[0, 4]), int(sums[0, 5])


def allreduce_summary(sim, device, group=None):
    """Per-GPU reduction kernels -> all-reduce (sum / min / max) -> global histograms -> all-reduce (sum).

    Returns (per-component summaries, total events, total faulted replications) identical on every rank.
    `sim` must have been created on the current torch CUDA stream of `device`."""
    import torch
    import torch.distributed as dist

    n = sim.n_comp
    distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    sums = torch.zeros(n * 6, dtype=torch.int64, device=device)
    ext = torch.zeros(n * 4, dtype=torch.int64, device=device)
    sim.reduce_stats_device(sums.data_ptr(), ext.data_ptr())
    e = ext.view(n, 4)
    mins = e[:, 0:2].contiguous()
    maxs = torch.bitwise_not(e[:, 2:4]).contiguous()  # kernel stores ~max so one atomicMin serves both
    if distributed:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(mins, op=dist.ReduceOp.MIN, group=group)
        dist.all_reduce(maxs, op=dist.ReduceOp.MAX, group=group)
    ext2 = torch.cat([mins, torch.bitwise_not(maxs)], dim=1).contiguous()
    hist = torch.zeros(n * 2 * capi.HIST_BINS, dtype=torch.int64, device=device)
    sim.histograms_device(ext2.data_ptr(), hist.data_ptr())
    if distributed:
        dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=group)
    host = [t.cpu().numpy().view(np.uint64) for t in (sums, mins, maxs, hist)]
    return summarize(host[0], host[1], host[2], host[3], n)

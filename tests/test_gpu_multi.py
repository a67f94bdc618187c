"""The multi-GPU plane on real GPUs: (1) qk_multi_* of the C ABI (one process, NCCL inside the library), (2) one process
per GPU with torch.distributed (NCCL) carrying the same exchange, (3) the C++ example that uses (1) without Python.
Each must return exactly what a single batch holding all the replications returns: sums, extrema, second moments,
histograms.  With one GPU in the box the D = 1 paths run; the D >= 2 parts skip."""
import os
import subprocess
import sys

import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    import torch

    return torch.cuda.device_count()


def _summary_fields(red):
    return [(int(r["sum_count"]), int(r["sum_time_lo"]), int(r["sum_time_hi"]), int(r["sum_level"]), int(r["min_count"]),
             int(r["max_count"]), int(r["min_time"]), int(r["max_time"]), [int(x) for x in r["sumsq_time_ns2"]],
             int(r["sumsq_count_lo"]), int(r["sumsq_level"]), r["hist_time"].tolist(), r["hist_count"].tolist())
            for r in red]


@pytest.mark.parametrize("n_dev", [1, 2, 8])
def test_qk_multi_equals_single_batch(gpu_lib, n_dev):
    if _n_gpus() < n_dev:
        pytest.skip("needs %d GPUs" % n_dev)
    from quokkasim_b200.multi import MultiGpuSimulation

    n, n_events, seed = 20011, 3000, 77  # an odd count: uneven shards
    m = H.discrete_queue()
    one = m.sim(gpu_lib, n, base_seed=seed, log_capacity=64)
    one.step_events(n_events)
    want = _summary_fields(one.reduce_stats())
    multi = MultiGpuSimulation(H.discrete_queue(), n, devices=list(range(n_dev)), base_seed=seed, log_capacity=64, lib=gpu_lib)
    assert multi.n_shards == n_dev
    multi.step_events(n_events)
    got, total = multi.stats()
    assert total["n_reps"] == n and total["n_events"] == n * n_events and total["n_faulted"] == 0
    assert _summary_fields(got) == want
    assert total["n_log_records"] == one.summary()["n_log_records"]
    # per-replication read-back goes through the shards: replication r of the job is r - rep0 of its shard
    for i in range(n_dev):
        view, rep0, cnt, dev = multi.shard(i)
        assert dev == i and cnt in (n // n_dev, n // n_dev + 1)
        for r in (0, cnt - 1):
            assert view.read_log(r)[0].tobytes() == one.read_log(rep0 + r)[0].tobytes()
    multi.close()
    one.close()


def _nccl_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(rank)
    device = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=device)
    from quokkasim_b200 import parallel, workloads

    n, n_events, seed = 20011, 3000, 77
    lo, hi = parallel.shard_range(n, rank, world)
    stream = torch.cuda.Stream(device)
    torch.cuda.set_stream(stream)
    # rank 0 launches on torch's stream, the others on a stream the library makes itself: both orderings must work
    sim = workloads.discrete_queue().sim(None, hi - lo, base_seed=seed, rep_offset=lo, device=rank,
                                         stream=stream.cuda_stream if rank == 0 else None)
    sim.step_events(n_events)
    comp, n_ev, n_fault = parallel.allreduce_summary(sim, device)
    if rank == 0:
        torch.save({"comp": comp, "n_ev": n_ev, "n_fault": n_fault}, out)
    dist.barrier()
    sim.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 8])
def test_torchrun_nccl_reduction_equals_single_batch(gpu_lib, world, tmp_path):
    if _n_gpus() < world:
        pytest.skip("needs %d GPUs" % world)
    import torch
    import torch.multiprocessing as mp

    from quokkasim_b200 import parallel

    out = str(tmp_path / "r0.pt")
    mp.spawn(_nccl_worker, args=(world, 29700 + os.getpid() % 200, out), nprocs=world, join=True)
    got = torch.load(out, weights_only=False)
    n, n_events, seed = 20011, 3000, 77
    one = H.discrete_queue().sim(gpu_lib, n, base_seed=seed)
    one.step_events(n_events)
    red = one.reduce_stats()
    assert got["n_ev"] == n * n_events and got["n_fault"] == 0
    for c, (g, r) in enumerate(zip(got["comp"], red)):
        assert g["sum_count"] == int(r["sum_count"]) and g["sum_level"] == int(r["sum_level"])
        assert g["sum_time_ns"] == (int(r["sum_time_hi"]) << 64) + int(r["sum_time_lo"])
        assert (g["min_count"], g["max_count"], g["min_time_ns"], g["max_time_ns"]) == (
            int(r["min_count"]), int(r["max_count"]), int(r["min_time"]), int(r["max_time"]))
        assert g["sumsq_time_ns"] == sum(int(r["sumsq_time_ns2"][k]) << (64 * k) for k in range(3))
        assert g["sumsq_count"] == (int(r["sumsq_count_hi"]) << 64) + int(r["sumsq_count_lo"])
        assert g["sumsq_level"] == int(r["sumsq_level"])
        assert g["hist_time"] == r["hist_time"].tolist() and g["hist_count"] == r["hist_count"].tolist()
    one.close()


def test_single_process_reduction_path_of_parallel_py(gpu_lib):
    """parallel.allreduce_summary without a process group (N = 1 of the bench's end-to-end leg), on a batch whose stream
    is NOT torch's current stream"""
    import torch

    from quokkasim_b200 import parallel

    m = H.haulage(n_src=3, per_queue=4)
    sim = m.sim(gpu_lib, 5000, base_seed=4)
    sim.step_events(2000)
    comp, n_ev, n_fault = parallel.allreduce_summary(sim, torch.device("cuda", 0))
    red = sim.reduce_stats()
    assert n_ev == 5000 * 2000 and n_fault == 0
    for g, r in zip(comp, red):
        assert g["sum_count"] == int(r["sum_count"]) and g["hist_time"] == r["hist_time"].tolist()
        assert g["sumsq_time_ns"] == sum(int(r["sumsq_time_ns2"][k]) << (64 * k) for k in range(3))
    sim.close()


def test_cpp_example_runs_the_job_on_every_gpu_without_python(gpu_lib, tmp_path):
    """examples/discrete_queue_multi.cpp: same summary for 1 GPU and for all GPUs of the box"""
    exe = str(tmp_path / "dqm")
    libdir = os.path.dirname(gpu_lib)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "discrete_queue_multi.cpp"), "-o", exe, "-L" + libdir,
                           "-l:" + os.path.basename(gpu_lib), "-Wl,-rpath," + libdir])
    outs = []
    for g in sorted({1, _n_gpus()}):
        o = subprocess.check_output([exe, "--replications", "30001", "--events", "2000", "--gpus", str(g)], text=True)
        assert "gpus %d " % g in o
        # (libnccl prints its version banner to stdout when NCCL_DEBUG asks for it)
        outs.append([ln for ln in o.splitlines() if not ln.startswith("step ") and not ln.startswith("NCCL version")])
    assert all(x == outs[0] for x in outs) and any(ln.startswith("Sink: sum_count = ") for ln in outs[0])
    one = H.discrete_queue().sim(gpu_lib, 30001, base_seed=1234)
    one.step_events(2000)
    assert "Sink: sum_count = %d " % int(one.reduce_stats()[6]["sum_count"]) in "\n".join(outs[0])
    one.close()

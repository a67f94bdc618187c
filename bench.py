#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 batched-replication executor.

Metric (BASELINE.json): simulated events/sec (whole box, 1M replications) of the discrete_queue model: 1,048,576
replications x 10,000 events, Philox seeds, split evenly over the N GPUs ("scaling": "strong"; --scaling weak keeps the
replications PER GPU fixed instead).  One "step" = Model::init + 10,000 executed scheduler actions of every
replication (1.05e10 events over the whole box).  --workload selects the other BASELINE configs (C3 vector_flow,
C4 serial_line, C5 haulage) at their stated sizes.

  python bench.py [--gpus N --steps K --warmup W]            # our arm (N>1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference [--gpus N ...]            # CPU arm: the oracle port of the reference engine on
                                                             # all host cores, bounded sample of the same workload

Prints ONE JSON line on rank 0.  See DESIGN.md "Measurement" for what every key means.  `roofline` describes the step
launch of the bench itself (K = 10,000 events per state residency: the state planes cross HBM once per launch, so the
algorithmic bytes are the log records); `roofline.lockstep_K1` is a short probe of the SAME kernel at one event per
launch, where the state planes cross HBM for every event (profiles/README.md section 2 has the whole K sweep).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# stdout carries exactly ONE line -- the JSON result.  Libraries print to file descriptor 1 behind Python's back (NCCL
# writes "NCCL version ..." there), so keep a private duplicate of the original stdout for the result and point fd 1
# at stderr for everybody else.
_RESULT_FD = os.dup(1)
os.dup2(2, 1)
sys.stdout = sys.stderr


def emit(obj):
    os.write(_RESULT_FD, (json.dumps(obj) + "\n").encode())


WORKLOADS = {
    # name: (builder, kwargs, replications (whole job), events per replication) -- BASELINE.json configs / SURVEY 8d
    "discrete_queue": ("discrete_queue", {}, 1048576, 10000),                      # C2
    "vector_flow": ("vector_flow", {"dim": 3}, 262144, 10000),                     # C3
    "serial_line": ("serial_line", {"n_proc": 32}, 100000, 1000000),               # C4
    "haulage": ("haulage", {"n_src": 8, "per_queue": 8}, 262144, 10000),           # C5 (one point of the sweep)
    "container_haulage": ("container_haulage", {"n_trucks": 8, "loaders": 2}, 262144, 10000),  # row f3
    "iron_ore_flow": ("vector_flow", {"dim": 5}, 262144, 10000),  # row f3: C3's topology carrying iron_ore.rs's resource
}


def build_model(name):
    from quokkasim_b200 import workloads

    fn, kw, _, _ = WORKLOADS[name]
    return getattr(workloads, fn)(**kw)


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------- CPU arm
def oracle_model(model):
    from oracle.lowering import lower_to_oracle  # bench.py's CPU legs are the one place the oracle is timed

    return lower_to_oracle(model)[0]


def time_oracle(om, n_reps, n_events, threads, log, base_seed=1234, rep0=0):
    t = time.perf_counter()
    total, _, status, _, _ = om.run_batch(base_seed, rep0, n_reps, n_events=n_events, log=log, threads=threads,
                                          want_stats=False)
    dt = time.perf_counter() - t
    return total, dt, int((status != 0).sum())


def cpu_baseline(model, n_events, log, target_s=12.0):
    om = oracle_model(model)
    threads = os.cpu_count() or 1
    cal_reps = max(threads, 16)
    total, dt, _ = time_oracle(om, cal_reps, n_events, threads, log)
    rate = total / dt
    reps = int(max(cal_reps, min(65536, target_s * rate / max(n_events, 1))))
    reps = max(threads, reps // threads * threads)
    total, dt, faulted = time_oracle(om, reps, n_events, threads, log)
    return {"value": total / dt, "unit": "events/s", "cores": threads, "kind": "port",
            "sample": "%d replications x %d events of the same model and Philox seeds (%.1f s), oracle/ C++ "
                      "restatement of the reference CPU path, %d std::threads, logging %s"
                      % (reps, n_events, dt, threads, "on" if log else "off"), "faulted": faulted}


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port; the Rust engine cannot be built here) on all host
    threads, each step a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    model = build_model(args.workload)
    _, _, reps_gpu, n_events = WORKLOADS[args.workload]
    n_events = args.events or n_events
    om = oracle_model(model)
    threads = os.cpu_count() or 1
    log = args.log != "off"
    total, dt, _ = time_oracle(om, max(threads, 16), n_events, threads, log)
    rate = total / dt
    reps = int(max(threads, min(65536, args.ref_step_seconds * rate / n_events)))
    reps = max(threads, reps // threads * threads)
    for _ in range(args.warmup):
        time_oracle(om, reps, n_events, threads, log)
    t_sum, ev_sum = 0.0, 0
    for _ in range(args.steps):
        total, dt, _ = time_oracle(om, reps, n_events, threads, log)
        t_sum += dt
        ev_sum += total
    value = ev_sum / t_sum
    sample = ("each step = %d replications x %d events (bounded sample of the %d-replication workload), oracle port "
              "of the reference CPU path on %d host threads" % (reps, n_events, args.reps or reps_gpu, threads))
    cfg = workload_config(args, args.reps or reps_gpu, n_events, 1)
    # same workload definition as our arm; what the reference arm actually runs per step is stated next to it
    cfg["reference_sample"] = ("each step of the reference arm = %d replications x %d events: a bounded sample of the "
                               "workload's %d replications" % (reps, n_events, args.reps or reps_gpu))
    cfg["reference_sample_replications_per_step"] = reps
    emit(({
        "impl": "reference", "metric": "simulated events/sec", "value": value, "unit": "events/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_sum / args.steps,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": value, "unit": "events/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "events/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def workload_config(args, reps_total, n_events, world):
    return {"workload": "%s batched: %d replications x %d events over %d GPU(s), Philox4x32-10 seeds"
                        % (args.workload, reps_total, n_events, world),
            "replications_total": reps_total, "replications_per_gpu": reps_total // world,
            "events_per_replication": n_events,
            "logging": args.log, "log_ring_records": args.log_capacity if args.log != "off" else 0,
            "state": args.state, "threads_per_block": args.threads_per_block,
            "events_per_launch": args.events_per_launch or n_events,
            "l2": "state planes (>= 0.5 GB) and log rings are larger than the 126 MB L2; no flush needed"}


# ----------------------------------------------------------------------------------------------- GPU arm
def lib_sha256(path=None):
    import hashlib

    from quokkasim_b200 import _capi

    h = hashlib.sha256()
    with open(path or _capi.DEFAULT_LIB, "rb") as f:
        for blk in iter(lambda: f.read(1 << 20), b""):
            h.update(blk)
    return h.hexdigest()


def lib_build_id(path=None):
    """What the library was built FROM (sha256 over csrc/, the public header and the compiler flags, compiled into the
    library: quokkasim_b200/build.py).  Two builds of the same sources have the same id although the files differ in a
    few dozen bytes of linker bookkeeping, which is why the ncu traffic record is keyed by it and not by the file hash."""
    from quokkasim_b200 import _capi
    from quokkasim_b200 import build as qbuild

    return qbuild.built_id(path or _capi.DEFAULT_LIB)


def median(xs):
    xs = sorted(xs)
    return xs[len(xs) // 2] if len(xs) % 2 else 0.5 * (xs[len(xs) // 2 - 1] + xs[len(xs) // 2])


def run_ours(args):
    import torch
    import torch.distributed as dist

    from quokkasim_b200 import parallel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; quokkasim_b200 has no CPU path (use --impl reference)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    _, _, reps_job, n_events = WORKLOADS[args.workload]
    reps_job = args.reps or reps_job
    if args.scaling == "weak":
        reps_total = reps_job * world  # fixed work per GPU
    else:
        reps_total = reps_job          # BASELINE: the whole box shares the job's replications
    rep_lo, rep_hi = parallel.shard_range(reps_total, rank, world)
    reps_gpu = rep_hi - rep_lo
    n_events = args.events or n_events
    log_cap = args.log_capacity if args.log != "off" else 0
    model = build_model(args.workload)
    # a real (non-NULL) stream: the library launches on it, torch events and NCCL order against it
    stream = torch.cuda.Stream(device)
    torch.cuda.set_stream(stream)
    flags = {"auto": 0, "hbm": 1, "chip": 2, "group": 4, "phased": 12, "split": 16, "qsplit": 48}[args.state]

    def new_sim(log_capacity, n=reps_gpu, off=rep_lo, seed=1234):
        return model.sim(args.lib, n, t0=0, base_seed=seed, rep_offset=off, log_capacity=log_capacity,
                         threads_per_block=args.threads_per_block, stream=stream.cuda_stream, device=local_rank,
                         flags=flags)

    sim = new_sim(log_cap)
    sim._materialize()  # device state allocated; inputs (model tables, seeds) resident before the timed region
    L, b = sim.L, sim._batch
    info = sim.info()
    state_bytes = info["state_bytes_per_rep"]

    K = args.events_per_launch or n_events

    def one_step():
        L.qk_batch_init(b, 0)
        if K >= n_events:
            L.qk_batch_step_events(b, n_events)
        else:
            L.qk_batch_step_events_chunked(b, n_events, K)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    for _ in range(args.warmup):
        one_step()
    barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms, launches = [], []
    ev0.record(stream)
    for _ in range(args.steps):
        one_step()
        # (reading the library's own CUDA events of the step just issued waits for that step: the next step's launch
        # is issued a few microseconds late, once per step of >= tens of milliseconds)
        ms_all, n_l = sim.last_step_ms()
        kernel_ms.append(ms_all / max(1, n_l))
        launches.append(n_l)
    ev1.record(stream)
    barrier()
    elapsed_ms = ev0.elapsed_time(ev1)
    clk = clocks.stop() if rank == 0 else None
    summ = sim.summary()
    step_chunks = sim.info()["last_step_chunks"]  # > 1 when the library cut the step into wave-sized launches
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=device)
    ev_total = torch.tensor([summ["n_events"], summ["n_log_records"], summ["n_faulted"]], dtype=torch.int64,
                            device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(ev_total, op=dist.ReduceOp.SUM)
    elapsed_ms = float(t.item())
    events_per_step = int(ev_total[0].item())
    records_per_step = int(ev_total[1].item())
    value = events_per_step * args.steps / (elapsed_ms * 1e-3)

    # ---- the same kernel where it IS an HBM kernel: one event per state residency (K = 1, pure lockstep).  Every launch
    # stages the hot planes from HBM and stores them back; short probe on the resident batch, CUDA events on its stream.
    lockstep = None
    if K >= n_events and not args.no_lockstep_probe:
        L.qk_batch_init(b, 0)
        L.qk_batch_step_events_chunked(b, 64, 1)
        s0 = sim.summary()
        L.qk_batch_step_events_chunked(b, 256, 1)
        ms_k1, nl_k1 = sim.last_step_ms()
        s1 = sim.summary()
        recs_k1 = (s1["n_log_records"] - s0["n_log_records"]) / max(1, nl_k1)
        lockstep = (ms_k1 / max(1, nl_k1), recs_k1, (s1["n_events"] - s0["n_events"]) / (ms_k1 * 1e-3))

    # ---- end to end through the public API with host buffers, every step: build the model tables on the host, create
    # the batch (H2D), init + step, reduce the summary statistics over the replications (and over the GPUs: one
    # all-gather + one all-reduce over NCCL), read them and a sample of the logs back (D2H).  Timed per step with the
    # host clock around a barrier + device synchronisation; the reported value uses the MEDIAN step (max over ranks).
    sim.close()
    del sim
    e2e_steps = max(1, args.e2e_steps)
    h2d = d2h = 0
    # untimed warm-up of the end-to-end path on a small batch: the first call of the reduction loads torch / NCCL kernels
    # (0.1 s once per process), which is not part of a step
    w = new_sim(log_cap, n=4096, off=rank * 4096, seed=1)
    w.step_events(16)
    parallel.allreduce_summary(w, device)
    w.summary()
    if log_cap:
        w.read_logs(0, min(4096, args.e2e_log_reps))
    w.close()
    step_s, phases = [], {"create": [], "init_step": [], "reduce": [], "d2h_logs": []}
    for _ in range(e2e_steps):
        barrier()
        t0 = time.perf_counter()
        s2 = new_sim(log_cap)
        s2._materialize()  # lowering + allocation + H2D of the tables + the init launch (asynchronous)
        t1 = time.perf_counter()
        s2.step_events(n_events)
        s2.synchronize()
        t2 = time.perf_counter()
        comp, n_ev, n_fault = parallel.allreduce_summary(s2, device)
        d2h = s2.n_comp * (20 + 64) * 8
        t3 = time.perf_counter()
        if log_cap:
            k = min(reps_gpu, args.e2e_log_reps)
            logs, totals = s2.read_logs(0, k)
            d2h += k * (log_cap * 32 + 12)
        torch.cuda.synchronize(device)
        t4 = time.perf_counter()
        h2d = s2.info()["table_bytes"] + 64  # model tables + batch config: the path's only host inputs
        step_s.append(t4 - t0)
        for kname, dt in (("create", t1 - t0), ("init_step", t2 - t1), ("reduce", t3 - t2), ("d2h_logs", t4 - t3)):
            phases[kname].append(dt)
        # releasing the batch is outside the timed region: its planes and log rings go back to the device's
        # stream-ordered pool, which is what the next step's batch is carved from (a batch that had to get its
        # gigabytes from the driver instead paid 0.1-0.8 s in `create`: the 1.26 spread of the first r02 run)
        s2.close()
    barrier()
    e2e_t = torch.tensor(step_s, dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_list = [float(x) for x in e2e_t.tolist()]
    e2e_med = median(e2e_list)
    e2e_value = events_per_step / e2e_med

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"
        # algorithmic bytes of one step on one GPU: B_alg = L + 2*S/K per event; the launches of a step
        # (one, or the wave-sliced / K-chunked sequence) each move the state planes of the replications they advance
        ev_gpu = events_per_step / world
        launches_per_step = int(median(launches))
        residencies = (n_events + K - 1) // K if K < n_events else max(1, step_chunks)
        log_bytes = 32.0 * records_per_step / world if log_cap else 0.0
        alg_bytes_step = log_bytes + 2.0 * state_bytes * reps_gpu * residencies
        k_ms = median(kernel_ms)
        step_kernel_ms = k_ms * launches_per_step
        achieved = alg_bytes_step / (step_kernel_ms * 1e-3) / 1e9
        # dram__bytes_read.sum + dram__bytes_write.sum of ONE step, from the committed ncu capture of this exact
        # workload shape AND this exact build (the source id compiled into libquokka_b200.so); null otherwise
        traffic = None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
            rec = tr.get(args.workload, {}).get(args.log)
            if (rec and rec["replications"] == reps_gpu and rec["events"] == n_events
                    and rec.get("lib_build_id") == lib_build_id(args.lib)):
                traffic = rec["dram_bytes_per_step"]
        except Exception:
            pass
        out = {
            "metric": "simulated events/sec", "value": value, "unit": "events/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed_ms / args.steps,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "u64",
            "data": "synthetic",
            "config": workload_config(args, reps_total, n_events, world),
            "e2e": {"value": e2e_value, "unit": "events/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "step_s": e2e_list, "median_step_s": e2e_med,
                    "spread": (max(e2e_list) - min(e2e_list)) / e2e_med,
                    "phases_s_median_rank0": {k: median(v) for k, v in phases.items()},
                    "log_sample_replications": min(reps_gpu, args.e2e_log_reps) if log_cap else 0},
            "gpu_launches": (1 + launches_per_step) * args.steps,
            "clocks": clk,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel": "qk_group_kernel" if info["lanes_per_replication"] > 1 else "qk_step_kernel",
                         "kernel_ms": k_ms, "step_kernel_ms": step_kernel_ms,
                         "algorithmic_bytes_per_launch": alg_bytes_step / launches_per_step,
                         "algorithmic_bytes_per_step": alg_bytes_step,
                         "bytes_per_event": alg_bytes_step / ev_gpu,
                         "state_bytes_per_replication": state_bytes, "events_per_residency_K": n_events // residencies,
                         "launches_per_step": launches_per_step, "lib_sha256": lib_sha256(args.lib),
                         "lib_build_id": lib_build_id(args.lib)},
            "faulted_replications": int(ev_total[2].item()),
            "per_gpu_events_per_s": value / world,
            "placement": {"state": "hbm" if info["state_in_hbm"] else (
                              "lane-group" if info["lanes_per_replication"] > 1 else (
                                  "split" if info.get("state_split") else "shared-memory")),
                          "lanes_per_replication": info["lanes_per_replication"],
                          "threads_per_block": info["threads_per_block"], "grid_blocks": info["grid_blocks"],
                          "resident_blocks_per_sm": info["resident_blocks_per_sm"],
                          "wave_blocks": info["wave_blocks"], "step_chunks": step_chunks},
        }
        if lockstep is not None:
            ms1, recs1, evps1 = lockstep
            alg1 = 2.0 * state_bytes * reps_gpu + (32.0 * recs1 if log_cap else 0.0)
            out["roofline"]["lockstep_K1"] = {
                "events_per_residency_K": 1, "kernel_ms": ms1, "algorithmic_bytes_per_launch": alg1,
                "achieved": alg1 / (ms1 * 1e-3) / 1e9, "frac": alg1 / (ms1 * 1e-3) / 1e9 / peak,
                "events_per_s_this_gpu": evps1,
                "note": "same kernel, 256 launches of one event each on the resident batch: the operating point where "
                        "the state planes cross HBM for every event (the bench itself runs K = events per step)"}
        if not args.no_cpu_baseline and world == 1:
            out["cpu_baseline"] = cpu_baseline(model, n_events, log_cap != 0)
        emit(out)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="discrete_queue", choices=sorted(WORKLOADS))
    ap.add_argument("--reps", type=int, default=0, help="replications per GPU (default: the workload's)")
    ap.add_argument("--events", type=int, default=0, help="events per replication (default: the workload's)")
    # (--logging: the same switch under a name torchrun's own parser does not claim: it rejects "--log" as ambiguous
    # with its --log-dir even after the script name)
    ap.add_argument("--log", "--logging", dest="log", default="ring", choices=["ring", "off"])
    ap.add_argument("--log-capacity", type=int, default=64)
    ap.add_argument("--threads-per-block", type=int, default=0)
    ap.add_argument("--state", default="auto", choices=["auto", "hbm", "chip", "group", "phased", "split", "qsplit"],
                    help="auto: the library chooses; hbm: operate on the HBM planes in place; chip: one thread per "
                         "replication, state staged in shared memory; group: a lane group per replication, state in "
                         "shared memory (large models)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong (BASELINE): the job's replications are split over the GPUs; weak: that many PER GPU")
    ap.add_argument("--events-per-launch", type=int, default=0,
                    help="K: events advanced per state residency (default: all events of a step in one launch); small K "
                         "makes the state round-trip HBM every K events -- the regime where the kernel is HBM-bound")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--e2e-log-reps", type=int, default=256)
    ap.add_argument("--ref-step-seconds", type=float, default=6.0)
    ap.add_argument("--lib", default=None, help="path of an alternative build of libquokka_b200.so (A/B experiments)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-lockstep-probe", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        sys.stderr.write("bench.py: note: the timing rules ask for >= 3 warm-up steps\n")
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

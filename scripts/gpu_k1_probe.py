"""What bounds a one-event launch (K = 1) of the discrete_queue step kernel: launches that stage the planes in and out
and advance NO event (step_until(0) on a model whose first events are later) against launches that advance one, for
several block sizes.  Prints microseconds per launch and the fraction of the measured HBM peak the plane traffic is."""
import json
import sys

import torch

sys.path.insert(0, ".")
from quokkasim_b200 import workloads  # noqa: E402

PEAK = 6551.7  # GB/s, MEASURED_PEAKS.json of this pool (driver-written, not on the box)
n = 1 << 20
for log_cap in (0, 64):
    for nt in (0, 32, 64, 128, 256):
        try:
            sim = workloads.discrete_queue().sim(None, n, base_seed=1, log_capacity=log_cap, threads_per_block=nt)
            sim.step_events(64)
            info = sim.info()
            S = info["state_bytes_per_rep"]
            L, b = sim.L, sim._batch
            t_now = int(sim.status(0, 1)[1][0])
            out = {}
            for name, fn in (("zero", lambda: L.qk_batch_step_until(b, t_now)),
                             ("one", lambda: L.qk_batch_step_events(b, 1))):
                for _ in range(8):
                    fn()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(torch.cuda.current_stream())
                tot = 0.0
                for _ in range(64):
                    fn()
                    ms, k = sim.last_step_ms()
                    tot += ms / max(1, k)
                out[name] = tot / 64 * 1e3
            planes = 2.0 * S * n
            print("log_cap %3d nt %3d (%s, %d B/rep): zero-event launch %.1f us = %.3f of peak ; one-event launch %.1f us = %.3f"
                  % (log_cap, nt, info.get("threads_per_block"), S, out["zero"], planes / out["zero"] / 1e3 / PEAK,
                     out["one"], planes / out["one"] / 1e3 / PEAK), flush=True)
            sim.close()
        except Exception as e:  # noqa: BLE001
            print("log_cap %d nt %d: %s" % (log_cap, nt, e), flush=True)

"""Where the end-to-end time of one bench step goes (wall clock per phase, one GPU)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quokkasim_b200 import workloads, parallel
dev = torch.device("cuda", 0)
model = workloads.discrete_queue()
for log_cap in (0, 64):
    for it in range(3):
        t = [time.perf_counter()]
        sim = model.sim(None, 1048576, base_seed=1234, log_capacity=log_cap)
        t.append(time.perf_counter()); sim._materialize(); sim.synchronize()
        t.append(time.perf_counter()); sim.step_events(10000); sim.synchronize()
        t.append(time.perf_counter()); parallel.allreduce_summary(sim, dev)
        t.append(time.perf_counter()); sim.summary()
        t.append(time.perf_counter())
        if log_cap:
            for r in range(256):
                sim.read_log(r)
        t.append(time.perf_counter()); sim.close()
        t.append(time.perf_counter())
        names = ["build", "create+init", "step", "reduce", "summary", "read_log x256", "close"]
        print("log_cap=%d " % log_cap + "  ".join("%s %.1f ms" % (n, 1e3 * (b - a)) for n, a, b in zip(names, t, t[1:])))

"""TEST INFRASTRUCTURE: run under LD_PRELOAD=libasan with the AddressSanitizer build of the host emulation
(tests/test_emul_asan.py).  Every state placement x logging mode over random models of all three families: an index that
leaves its plane is an error here, where the plain host build (and a GPU without compute-sanitizer) would read on."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_fuzz_parity as T  # noqa: E402
from quokkasim_b200 import workloads as W  # noqa: E402

lib = sys.argv[1]
n = 0
PLACEMENTS = (0, 1, 16, 16 | 32, 4 | (4 << 8), 4 | 8 | (4 << 8))
for seed in (800, 801, 802, 805, 700, 703):
    for flags in PLACEMENTS:
        for log in (0, 4096):
            sim = T.random_model(seed).sim(lib, 3, base_seed=seed, log_capacity=log, flags=flags)
            sim.step_until(60 * 10**9)
            sim.status()
            sim.close()
            n += 1
    for flags in (0, 16, 48):
        sim = T.random_vector_model(seed).sim(lib, 3, base_seed=seed, log_capacity=4096, flags=flags)
        sim.step_until(60 * 10**9)
        sim.status()
        sim.close()
        n += 1
for name in ("assembly_line", "car_workshop", "discrete_queue", "container_haulage", "vector_flow", "delay_testing"):
    for flags in (0, 1, 16, 48):
        sim = getattr(W, name)().sim(lib, 3, base_seed=1, log_capacity=4096, flags=flags)
        sim.step_events(3000)
        sim.status()
        sim.close()
        n += 1
print("asan sweep ok", n)

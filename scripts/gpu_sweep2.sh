#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
B="python bench.py --steps 2 --warmup 1 --reps 262144 --events 2000 --no-cpu-baseline --e2e-steps 1"
P='import sys,json; d=json.loads(sys.stdin.readline()); print("value %.3e e2e %.3e kernel_ms %.2f frac %.4f" % (d["value"], d["e2e"]["value"], d["roofline"]["kernel_ms"], d["roofline"]["frac"]))'
for nt in ${NTS:-0 64 96 128 160 192 224 256}; do
  for lg in off ring; do
    echo -n "nt=$nt log=$lg  "; timeout 200 $B --threads-per-block $nt --log $lg 2>/dev/null | python -c "$P"
  done
done
for lg in off ring; do echo -n "HBM log=$lg  "; timeout 300 $B --state hbm --log $lg 2>/dev/null | python -c "$P"; done
if [ -z "$NOFULL" ]; then
echo "== full size, log ring"; timeout 600 python bench.py --steps 2 --warmup 1 > gpurun_out/bench_full_ring.json 2> gpurun_out/bench_full_ring.err; cut -c1-400 gpurun_out/bench_full_ring.json
echo "== full size, log off"; timeout 600 python bench.py --steps 2 --warmup 1 --log off --no-cpu-baseline > gpurun_out/bench_full_off.json 2> gpurun_out/bench_full_off.err; cut -c1-400 gpurun_out/bench_full_off.json
fi

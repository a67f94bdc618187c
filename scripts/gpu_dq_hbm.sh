#!/bin/bash
for nt in 64 128 256; do
  for lg in off ring; do
    timeout 600 python bench.py --steps 2 --warmup 1 --reps 262144 --events 2000 --no-cpu-baseline --e2e-steps 1 --log $lg --state hbm --threads-per-block $nt > gpurun_out/wlh.json 2> gpurun_out/wlh.err
    echo "== discrete_queue hbm nt=$nt log=$lg $(python -c "import sys,json; d=json.loads(open('gpurun_out/wlh.json').readline()); print('value %.3e kernel_ms %.2f' % (d['value'], d['roofline']['kernel_ms']))" 2>&1 | tail -1)"
  done
done

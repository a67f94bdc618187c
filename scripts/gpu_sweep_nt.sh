#!/bin/bash
# block-size sweep of the step kernel (reduced size; numbers are relative, not bench values)
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --reps 262144 --events 2000 --no-cpu-baseline --e2e-steps 1"
for nt in ${NTS:-64 96 192}; do
  for lg in off ring; do
    echo "== nt=$nt log=$lg $(timeout 200 $B --threads-per-block $nt --log $lg 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('value %.3e e2e %.3e kernel_ms %.2f frac %.4f' % (d['value'], d['e2e']['value'], d['roofline']['kernel_ms'], d['roofline']['frac']))")"
  done
done

#!/bin/bash
B="python bench.py --steps 2 --warmup 1 --reps 524288 --events 4000 --no-cpu-baseline --e2e-steps 1"
for nt in 64 128; do
for lg in off; do
  for pad in 0 4000 9000 14000; do
    p=$((pad*nt/64))
    echo "== nt=$nt log=$lg pad=$p $(QK_DEBUG_EXTRA_SMEM=$p timeout 200 $B --log $lg --threads-per-block $nt 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('value %.3e kernel_ms %.2f' % (d['value'], d['roofline']['kernel_ms']))")"
  done
done
done

#!/bin/bash
# does the step kernel speed up with more resident warps?  pad shared memory per block to lower residency.
B="python bench.py --steps 2 --warmup 1 --reps 524288 --events 4000 --no-cpu-baseline --e2e-steps 1 --threads-per-block 64"
for lg in off ring; do
  for pad in 0 4000 10000 16000 30000; do
    echo "== log=$lg pad=$pad $(QK_DEBUG_EXTRA_SMEM=$pad timeout 200 $B --log $lg 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('value %.3e kernel_ms %.2f' % (d['value'], d['roofline']['kernel_ms']))")"
  done
done

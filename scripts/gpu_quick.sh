#!/bin/bash
B="python bench.py --steps 2 --warmup 1 --reps 524288 --events 4000 --no-cpu-baseline --e2e-steps 1"
for lg in off ring; do
  echo "== log=$lg $(timeout 200 $B --log $lg 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('value %.3e kernel_ms %.2f' % (d['value'], d['roofline']['kernel_ms']))")"
done

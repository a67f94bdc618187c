#!/bin/bash
# events per state residency K: where the step kernel becomes HBM-bound (state planes in and out every K events)
for lg in off ring; do
for K in 1 2 4 16 64 256; do
  ev=2048; [ $K -ge 16 ] && ev=4096
  echo "== log=$lg K=$K $(timeout 300 python bench.py --steps 2 --warmup 1 --events $ev --events-per-launch $K --no-cpu-baseline --e2e-steps 1 --log $lg 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); r=d['roofline']; print('value %.4e kernel_us %.1f achieved %.0f GB/s frac %.3f B/event %.1f' % (d['value'], 1e3*r['kernel_ms'], r['achieved'], r['frac'], r['bytes_per_event']))")"
done
done

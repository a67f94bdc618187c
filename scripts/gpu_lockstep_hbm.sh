#!/bin/bash
for nt in 128 256; do
for K in 1 2 4; do
  echo "== hbm nt=$nt K=$K $(timeout 300 python bench.py --steps 2 --warmup 1 --events 2048 --events-per-launch $K --threads-per-block $nt --state hbm --no-cpu-baseline --e2e-steps 1 --log off 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); r=d['roofline']; print('value %.4e kernel_us %.1f achieved %.0f GB/s frac %.3f' % (d['value'], 1e3*r['kernel_ms'], r['achieved'], r['frac']))")"
done
done

#!/bin/bash
mkdir -p gpurun_out
for wl in serial_line haulage vector_flow; do
  for nt in 64 128; do
    timeout 600 python bench.py --workload $wl --steps 1 --warmup 1 --reps 65536 --events 2000 --no-cpu-baseline --e2e-steps 1 --log off --state hbm --threads-per-block $nt > gpurun_out/wlh.json 2> gpurun_out/wlh.err
    echo "== $wl hbm nt=$nt $(python -c "import sys,json; d=json.loads(open('gpurun_out/wlh.json').readline()); print('value %.3e kernel_ms %.2f' % (d['value'], d['roofline']['kernel_ms']))" 2>&1 | tail -1)"
  done
done

#!/bin/bash
# round 2, session C: ncu --set full of the lane-group kernel (haulage + serial line, statistics only, 65,536 x 1,000)
mkdir -p gpurun_out/r2c
O=gpurun_out/r2c
export QK_SLICE=0
for wl in haulage serial_line; do
  C="python bench.py --workload $wl --steps 1 --warmup 0 --reps 65536 --events 1000 --no-cpu-baseline --no-lockstep-probe --e2e-steps 1 --log off --state group"
  timeout 300 $C > $O/plain_$wl.log 2>&1 || { echo "plain run failed"; tail -5 $O/plain_$wl.log; continue; }
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:qk_group_kernel -s 2 -c 1 -o $O/prof_group_$wl -f $C > $O/ncu_$wl.log 2>&1
  tail -2 $O/ncu_$wl.log
done
ls -la $O

#!/bin/bash
# ncu captures of the step launch (3rd qk_step_kernel launch of this command: materialize-init, step-init, STEP)
mkdir -p gpurun_out
for lg in off ring; do
  C="python bench.py --steps 1 --warmup 0 --reps 131072 --events 1000 --no-cpu-baseline --e2e-steps 1 --log $lg"
  $C > gpurun_out/plain_$lg.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:qk_step_kernel -s 2 -c 1 -o gpurun_out/prof_$lg -f $C > gpurun_out/ncu_full_$lg.log 2>&1
done
C="python bench.py --steps 2 --warmup 1 --reps 131072 --events 1000 --no-cpu-baseline --e2e-steps 1 --log ring"
$C > gpurun_out/plain_l.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_ring.csv $C > gpurun_out/ncu_launch_ring.log 2>&1
ls -la gpurun_out | head -30

#!/bin/bash
# full-size (1,048,576 x 10,000) captures: launch list + one --set full capture of the step launch per logging mode.
# Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
for lg in ring off; do
  C="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --e2e-steps 1 --log $lg"
  $C > gpurun_out/full_plain_$lg.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/full_plain_$lg.log; continue; }
  ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/full_launches_$lg.csv $C > gpurun_out/full_ncu_launch_$lg.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:qk_step_kernel -s 2 -c 1 -o gpurun_out/full_prof_$lg -f $C > gpurun_out/full_ncu_$lg.log 2>&1
  tail -2 gpurun_out/full_ncu_$lg.log
done
ls -la gpurun_out | grep full_

#!/bin/bash
# one --set full capture of the step kernel (131072 x 1000) for the logging modes given in LGS
mkdir -p gpurun_out/ev
O=gpurun_out/ev
for lg in ${LGS:-ring}; do
  C="python bench.py --steps 1 --warmup 0 --reps 131072 --events 1000 --no-cpu-baseline --e2e-steps 1 --log $lg"
  $C > $O/plain_small_$lg.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:qk_step_kernel -s 2 -c 1 -o $O/prof_small_$lg -f $C > $O/ncu_small_$lg.log 2>&1
  tail -1 $O/ncu_small_$lg.log
done

#!/bin/bash
mkdir -p gpurun_out
lg=${LG:-off}
C="python bench.py --steps 1 --warmup 0 --reps 131072 --events 1000 --no-cpu-baseline --e2e-steps 1 --log $lg"
$C > gpurun_out/plain_$lg.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:qk_step_kernel -s 2 -c 1 -o gpurun_out/prof_$lg -f $C > gpurun_out/ncu_full_$lg.log 2>&1

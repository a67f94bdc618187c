#!/bin/bash
C="python bench.py --steps 1 --warmup 0 --reps 524288 --events 4000 --no-cpu-baseline --e2e-steps 1 --log off ${EXTRA}"
$C > gpurun_out/plain_occ.log 2>&1 && ncu --metrics sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,launch__occupancy_limit_registers,launch__occupancy_limit_shared_mem,launch__block_size,smsp__inst_executed.sum --clock-control none -k regex:qk_step_kernel -s 2 -c 1 $C 2>&1 | grep -E "warps_active|issue_active|duration|occupancy|block_size|inst_executed"

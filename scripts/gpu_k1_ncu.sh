#!/bin/bash
# ncu counters of ONE step launch at K = 1 (one event per state residency), both logging modes
mkdir -p gpurun_out/ev
for lg in off ring; do
  C="python bench.py --steps 1 --warmup 1 --events 64 --events-per-launch 1 --no-cpu-baseline --e2e-steps 1 --log $lg"
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,launch__registers_per_thread \
    --clock-control none -k regex:qk_step_kernel -s 100 -c 1 --csv --log-file gpurun_out/ev/k1_$lg.csv $C > gpurun_out/ev/k1_$lg.log 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/ev/k1_$lg.csv')) if len(r)>10]
h=rows[0]
print('== K=1 log=$lg', rows[1][h.index('Kernel Name')])
for r in rows[1:]:
    print('  %-70s %s %s'%(r[h.index('Metric Name')], r[h.index('Metric Value')], r[h.index('Metric Unit')]))
PY
done

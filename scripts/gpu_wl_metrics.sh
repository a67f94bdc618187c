#!/bin/bash
# memory-side metrics of the step kernel for the large (HBM-resident) models
mkdir -p gpurun_out/ev
for wl in ${WLS:-haulage serial_line}; do
  C="python bench.py --workload $wl --steps 1 --warmup 0 --reps 65536 --events 2000 --no-cpu-baseline --e2e-steps 1 --log off"
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,launch__registers_per_thread,smsp__average_warp_latency_per_inst_issued.ratio \
    --clock-control none -k regex:qk_step_kernel -s 2 -c 1 --csv --log-file gpurun_out/ev/wlm_$wl.csv $C > gpurun_out/ev/wlm_$wl.log 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/ev/wlm_$wl.csv')) if len(r)>10]
h=rows[0]
print('== $wl')
for r in rows[1:]:
    print('  %-70s %s %s'%(r[h.index('Metric Name')], r[h.index('Metric Value')], r[h.index('Metric Unit')]))
PY
done

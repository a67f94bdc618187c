#!/bin/bash
# Session t: the build with the small-resource arena / cache (host side only; device code unchanged): GPU tests, DRAM
# traffic records re-keyed to this build id, the end-to-end leg on one eighth of C2 (what one of 8 GPUs runs).
O=gpurun_out/ev2; mkdir -p $O
timeout 600 python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest_gpu.log
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,launch__registers_per_thread,launch__block_size,launch__grid_size,launch__occupancy_limit_registers,launch__occupancy_limit_shared_mem,lts__t_sector_hit_rate.pct
for lg in ring off; do
  QK_SLICE=0 timeout 300 ncu --metrics $M --clock-control none -k regex:qk_step_kernel -s 2 -c 1 --csv --log-file $O/traffic_discrete_queue_$lg.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --e2e-steps 1 --no-lockstep-probe --log $lg > $O/ncu_traffic_dq_$lg.log 2>&1
  QK_SLICE=0 timeout 300 ncu --metrics $M --clock-control none -k regex:qk_step_kernel -s 2 -c 1 --csv --log-file $O/traffic_haulage_$lg.csv python bench.py --workload haulage --steps 1 --warmup 0 --no-cpu-baseline --e2e-steps 1 --no-lockstep-probe --log $lg > $O/ncu_traffic_h_$lg.log 2>&1
done
timeout 300 python bench.py --reps 131072 --no-cpu-baseline --no-lockstep-probe --e2e-steps 9 > $O/bench_c2_ring_eighth.json 2> $O/bench_c2_ring_eighth.err
python -c "
import json; d=json.loads(open('$O/bench_c2_ring_eighth.json').readline()); print('eighth value %.4e e2e %.4e ratio %.3f spread %.3f' % (d['value'], d['e2e']['value'], d['e2e']['value']/d['value'], d['e2e']['spread']), d['e2e']['phases_s_median_rank0'])"

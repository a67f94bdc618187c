#!/bin/bash
# Round evidence in one GPU session.  Bench values come only from runs WITHOUT a profiler; ncu runs follow each plain run.
mkdir -p gpurun_out/ev
O=gpurun_out/ev
python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest_gpu.log
python bench.py > $O/bench_full_ring.json 2> $O/bench_full_ring.err; echo "bench ring rc=$?"
python bench.py --log off > $O/bench_full_off.json 2> $O/bench_full_off.err; echo "bench off rc=$?"
python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err; echo "bench reference rc=$?"
# launch list of the default command (per-launch durations under ncu are cold-cache and serialised: shares, not values)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_default.csv python bench.py --no-cpu-baseline > $O/ncu_launches.log 2>&1
# DRAM traffic of the full-size step launch (metrics-only capture, 3rd qk_step_kernel launch = first timed-shape step)
for lg in ring off; do
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,launch__registers_per_thread,launch__block_size,launch__occupancy_limit_registers,launch__occupancy_limit_shared_mem \
    --clock-control none -k regex:qk_step_kernel -s 2 -c 1 --csv --log-file $O/traffic_full_$lg.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --e2e-steps 1 --log $lg > $O/ncu_traffic_$lg.log 2>&1
done
# one --set full capture per logging mode at 131072 x 1000 (the full-size kernel runs 0.4-0.7 s: ~40 replays of it cost 9 GPU-minutes)
for lg in ring off; do
  C="python bench.py --steps 1 --warmup 0 --reps 131072 --events 1000 --no-cpu-baseline --e2e-steps 1 --log $lg"
  $C > $O/plain_small_$lg.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:qk_step_kernel -s 2 -c 1 -o $O/prof_small_$lg -f $C > $O/ncu_small_$lg.log 2>&1
done
# events per state residency K: where the kernel becomes HBM-bound (no profiler)
bash scripts/gpu_lockstep.sh > $O/lockstep_sweep.txt 2>&1
ls -la $O

#!/bin/bash
# Round-2 evidence of the FINAL build in one GPU session (the tree with user-defined vector resources and the TMA-tile
# staging): same layout as scripts/gpu_evidence_r02.sh, trimmed to the GPU minutes left.  Bench values come only from runs
# WITHOUT a profiler; every ncu run follows a plain run of the same command that exited 0.
# Output: gpurun_out/ev2 (scratch) -> scripts/collect_evidence_r02.py -> profiles/.
mkdir -p gpurun_out/ev2
O=gpurun_out/ev2
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $O/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --durations=10 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
run() { name=$1; shift; timeout 900 python bench.py "$@" > $O/$name.json 2> $O/$name.err; echo "$name rc=$? $(python - $O/$name.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).readline()); r=d.get('roofline',{})
    print('value %.4e e2e %.4e frac %.4f K1 %s %s cpu %.3e' % (d['value'], d['e2e']['value'], r.get('frac',0), (r.get('lockstep_K1') or {}).get('frac'), d.get('placement',{}).get('state'), d.get('cpu_baseline',{}).get('value',0)))
except Exception as e: print('ERR', e)
PY
)"; }
run bench_c2_ring
run bench_c2_off --log off --no-cpu-baseline
run bench_reference --impl reference
run bench_c3_vector_flow_ring --workload vector_flow
run bench_c3_vector_flow_off --workload vector_flow --log off --no-cpu-baseline
run bench_c5_haulage_ring --workload haulage
run bench_c5_haulage_off --workload haulage --log off --no-cpu-baseline
run bench_c4_serial_line_off --workload serial_line --log off --steps 2 --warmup 3 --e2e-steps 2 --no-lockstep-probe
run bench_container_haulage_ring --workload container_haulage --no-cpu-baseline --steps 3 --e2e-steps 2 --no-lockstep-probe
run bench_iron_ore_flow_ring --workload iron_ore_flow --no-lockstep-probe --steps 5 --e2e-steps 3
run bench_iron_ore_flow_off --workload iron_ore_flow --log off --no-cpu-baseline --no-lockstep-probe --steps 5 --e2e-steps 3
run bench_c2_ring_eighth --reps 131072 --no-cpu-baseline --no-lockstep-probe
run bench_c2_off_eighth --reps 131072 --log off --no-cpu-baseline --no-lockstep-probe
for n in 1024 16384 262144 4194304; do
  run sweep_c5_off_$n --workload haulage --reps $n --events 2000 --log off --no-cpu-baseline --no-lockstep-probe --e2e-steps 1 --steps 2
done
for n in 16384 262144; do
  run sweep_c5_ring_$n --workload haulage --reps $n --events 2000 --no-cpu-baseline --no-lockstep-probe --e2e-steps 1 --steps 2
done
run sweep_c5_off_8388608 --workload haulage --reps 8388608 --events 1000 --log off --no-cpu-baseline --no-lockstep-probe --e2e-steps 1 --steps 1 --warmup 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_default.csv python bench.py --no-cpu-baseline > $O/ncu_launches.log 2>&1
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,launch__registers_per_thread,launch__block_size,launch__grid_size,launch__occupancy_limit_registers,launch__occupancy_limit_shared_mem,lts__t_sector_hit_rate.pct
for lg in ring off; do
  QK_SLICE=0 timeout 600 ncu --metrics $M --clock-control none -k regex:qk_step_kernel -s 2 -c 1 --csv --log-file $O/traffic_discrete_queue_$lg.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --e2e-steps 1 --no-lockstep-probe --log $lg > $O/ncu_traffic_dq_$lg.log 2>&1
  QK_SLICE=0 timeout 600 ncu --metrics $M --clock-control none -k regex:qk_step_kernel -s 2 -c 1 --csv --log-file $O/traffic_haulage_$lg.csv python bench.py --workload haulage --steps 1 --warmup 0 --no-cpu-baseline --e2e-steps 1 --no-lockstep-probe --log $lg > $O/ncu_traffic_h_$lg.log 2>&1
done
export QK_SLICE=0
full() { tag=$1; shift; C="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-lockstep-probe --e2e-steps 1 $*"
  timeout 300 $C > $O/plain_$tag.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:qk_step_kernel -s 2 -c 1 -o $O/prof_$tag -f $C > $O/ncu_$tag.log 2>&1; tail -1 $O/ncu_$tag.log; }
full dq_ring --reps 131072 --events 1000 --log ring
full haulage_off --workload haulage --reps 65536 --events 1000 --log off
full dq_k1_ring --reps 1048576 --events 64 --events-per-launch 1 --log ring
unset QK_SLICE
for lg in off ring; do
for K in 1 2 4 16 256; do
  ev=2048; [ $K -ge 16 ] && ev=4096
  echo "== log=$lg K=$K $(timeout 300 python bench.py --steps 2 --warmup 1 --events $ev --events-per-launch $K --no-cpu-baseline --e2e-steps 1 --log $lg 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); r=d['roofline']; print('value %.4e kernel_us %.1f achieved %.0f GB/s frac %.3f B/event %.1f' % (d['value'], 1e3*r['kernel_ms'], r['achieved'], r['frac'], r['bytes_per_event']))")"
done
done > $O/lockstep_sweep.txt 2>&1
ls $O | wc -l

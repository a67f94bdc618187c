#!/bin/bash
# GPU session: parity tests, block-size / state-placement sweep, full-size bench, ncu captures.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
B="python bench.py --steps 2 --warmup 1 --reps 262144 --events 2000 --no-cpu-baseline --e2e-steps 1"
for nt in 0 64 128 192 256; do
  for lg in off ring; do
    echo "== nt=$nt log=$lg" ; timeout 200 $B --threads-per-block $nt --log $lg 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('value %.3e e2e %.3e kernel_ms %.2f frac %.4f' % (d['value'], d['e2e']['value'], d['roofline']['kernel_ms'], d['roofline']['frac']))"
  done
done
for lg in off ring; do
  echo "== HBM log=$lg"; timeout 300 $B --state hbm --log $lg 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('value %.3e e2e %.3e kernel_ms %.2f frac %.4f' % (d['value'], d['e2e']['value'], d['roofline']['kernel_ms'], d['roofline']['frac']))"
done
echo "== full size, log ring"; timeout 600 python bench.py --steps 2 --warmup 1 > gpurun_out/bench_full_ring.json 2> gpurun_out/bench_full_ring.err; cat gpurun_out/bench_full_ring.json
echo "== full size, log off"; timeout 600 python bench.py --steps 2 --warmup 1 --log off --no-cpu-baseline > gpurun_out/bench_full_off.json 2> gpurun_out/bench_full_off.err; cat gpurun_out/bench_full_off.json
# ncu: per-launch durations, then one full capture of the step kernel (plain run first, same args)
C="python bench.py --steps 1 --warmup 0 --reps 131072 --events 1000 --no-cpu-baseline --e2e-steps 1 --log off"
$C > gpurun_out/plain_off.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_off.csv $C > gpurun_out/ncu_launch_off.log 2>&1
$C > gpurun_out/plain_off2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:qk_step_kernel -s 1 -c 1 -o gpurun_out/prof_off $C > gpurun_out/ncu_full_off.log 2>&1
C2="python bench.py --steps 1 --warmup 0 --reps 131072 --events 1000 --no-cpu-baseline --e2e-steps 1 --log ring"
$C2 > gpurun_out/plain_ring.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:qk_step_kernel -s 1 -c 1 -o gpurun_out/prof_ring $C2 > gpurun_out/ncu_full_ring.log 2>&1
ls -la gpurun_out/

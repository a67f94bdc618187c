#!/bin/bash
# round 2, session E (1 GPU): all GPU tests incl. C4 at full size; A/B of the lane-group kernel variants; the bench
# lines of C3 / C4 / C5 at BASELINE sizes with the default placement; ncu of the lane-group kernel.
mkdir -p gpurun_out/r2e
O=gpurun_out/r2e
timeout 2400 python -m pytest tests -m gpu -x -q --durations=8 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -14 $O/pytest_gpu.log
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).readline())
    print(sys.argv[1].split('/')[-1], 'value %.4e ms %.1f e2e %.4e spread %.3f chunks %s frac %.4f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['spread'], d['placement']['step_chunks'], d['roofline']['frac']), d['placement']['state'], d['placement']['lanes_per_replication'], d['placement']['threads_per_block'], 'cpu %.3e' % d.get('cpu_baseline',{}).get('value',0))
except Exception as e:
    print(sys.argv[1], 'ERR', e)
PY
}
B="python bench.py --no-cpu-baseline --no-lockstep-probe --e2e-steps 1 --steps 2 --warmup 1"
for wl in haulage serial_line; do
  for lib in default build/ab/lib_shfl.so build/ab/lib_eager.so; do
    for lg in off ring; do
      tag=$(basename $lib .so)
      L=""; [ $lib != default ] && L="--lib $lib"
      timeout 600 $B --workload $wl --reps 65536 --events 2000 --state group --log $lg $L > $O/ab_${wl}_${tag}_$lg.json 2> $O/ab_${wl}_${tag}_$lg.err
      show $O/ab_${wl}_${tag}_$lg.json
    done
  done
done
# thread-per-replication kernel: slot rows staged by bulk asynchronous copies (this tree) against the per-thread LDGSTS
# staging of the commit before (build/ab/lib_notma.so): full-size C2 line, with the K = 1 lockstep probe
for lib in default build/ab/lib_notma.so; do
  for lg in ring off; do
    tag=$(basename $lib .so)
    L=""; [ $lib != default ] && L="--lib $lib"
    timeout 600 python bench.py --no-cpu-baseline --log $lg $L > $O/tma_c2_${tag}_$lg.json 2> $O/tma_c2_${tag}_$lg.err
    show $O/tma_c2_${tag}_$lg.json
    python -c "
import json,sys
d=json.loads(open('$O/tma_c2_${tag}_$lg.json').readline()); k=d['roofline']['lockstep_K1']
print('   lockstep K=1: %.4f ms  achieved %.0f GB/s  frac %.3f  events/s %.3e' % (k['kernel_ms'], k['achieved'], k['frac'], k['events_per_s_this_gpu']))"
  done
done
# BASELINE sizes, default placement, statistics only and with the 64-record log ring
for wl in vector_flow haulage; do
  for lg in off ring; do
    timeout 900 python bench.py --workload $wl --log $lg $([ $lg = ring ] && echo --no-cpu-baseline) > $O/bench_${wl}_$lg.json 2> $O/bench_${wl}_$lg.err; show $O/bench_${wl}_$lg.json
  done
done
timeout 1500 python bench.py --workload serial_line --log off --steps 1 --warmup 3 --e2e-steps 1 --no-lockstep-probe > $O/bench_serial_line_off.json 2> $O/bench_serial_line_off.err; show $O/bench_serial_line_off.json
timeout 1500 python bench.py --workload serial_line --log ring --steps 1 --warmup 1 --e2e-steps 1 --no-lockstep-probe --no-cpu-baseline > $O/bench_serial_line_ring.json 2> $O/bench_serial_line_ring.err; show $O/bench_serial_line_ring.json
# ncu --set full of the lane-group kernel as it is now (haulage, statistics only, 65,536 x 1,000, one plain launch)
export QK_SLICE=0
C="python bench.py --workload haulage --steps 1 --warmup 0 --reps 65536 --events 1000 --no-cpu-baseline --no-lockstep-probe --e2e-steps 1 --log off"
timeout 300 $C > $O/plain_haulage.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:qk_group_kernel -s 2 -c 1 -o $O/prof_group_haulage -f $C > $O/ncu_haulage.log 2>&1
tail -2 $O/ncu_haulage.log
C="python bench.py --workload serial_line --steps 1 --warmup 0 --reps 65536 --events 1000 --no-cpu-baseline --no-lockstep-probe --e2e-steps 1 --log off"
timeout 300 $C > $O/plain_serial.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:qk_group_kernel -s 2 -c 1 -o $O/prof_group_serial -f $C > $O/ncu_serial.log 2>&1
tail -2 $O/ncu_serial.log

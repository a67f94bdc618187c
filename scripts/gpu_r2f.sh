#!/bin/bash
# round 2, session F (1 GPU): all GPU tests; TMA staging A/B of the thread-per-replication kernel (C2 + K = 1 probe);
# first run of the PHASED lane-group kernel: group width x blocks per SM on the two large models.
mkdir -p gpurun_out/r2f
O=gpurun_out/r2f
timeout 2400 python -m pytest tests -m gpu -x -q --durations=6 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 $O/pytest_gpu.log
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).readline())
    k=d['roofline'].get('lockstep_K1')
    print(sys.argv[1].split('/')[-1], 'value %.4e ms %.1f e2e %.4e spread %.3f chunks %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['spread'], d['placement']['step_chunks']), d['placement']['state'], d['placement']['lanes_per_replication'], d['placement']['threads_per_block'], d['placement']['resident_blocks_per_sm'], ('K1 frac %.3f %.4f ms' % (k['frac'], k['kernel_ms'])) if k else '')
except Exception as e:
    print(sys.argv[1], 'ERR', e)
PY
}
for lib in default build/ab/lib_notma.so; do
  for lg in ring off; do
    tag=$(basename $lib .so)
    L=""; [ $lib != default ] && L="--lib $lib"
    timeout 600 python bench.py --no-cpu-baseline --log $lg $L > $O/tma_c2_${tag}_$lg.json 2> $O/tma_c2_${tag}_$lg.err
    show $O/tma_c2_${tag}_$lg.json
  done
done
B="python bench.py --no-cpu-baseline --no-lockstep-probe --e2e-steps 1 --steps 2 --warmup 1"
for wl in haulage serial_line; do
  timeout 600 $B --workload $wl --reps 65536 --events 2000 --state group --log off > $O/ph_${wl}_group8_off.json 2> $O/ph_${wl}_group8_off.err; show $O/ph_${wl}_group8_off.json
  for g in 8 4 16; do
    for nb in 1 2 3 4; do
      QK_GROUP_LANES=$g QK_PHASED_BLOCKS=$nb timeout 600 $B --workload $wl --reps 65536 --events 2000 --state phased --log off > $O/ph_${wl}_g${g}_b${nb}_off.json 2> $O/ph_${wl}_g${g}_b${nb}_off.err
      show $O/ph_${wl}_g${g}_b${nb}_off.json
    done
  done
  QK_GROUP_LANES=8 QK_PHASED_BLOCKS=2 timeout 600 $B --workload $wl --reps 65536 --events 2000 --state phased --log ring > $O/ph_${wl}_g8_b2_ring.json 2> $O/ph_${wl}_g8_b2_ring.err; show $O/ph_${wl}_g8_b2_ring.json
done

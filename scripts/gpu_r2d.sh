#!/bin/bash
# round 2, session D (1 GPU): tests; placement sweep after the quick path for delay-mode components, REDUX reduction and
# eager refill; sliced stepping (per-part chunk launches on side streams) on C2 full size and on the N = 8 shard shape.
mkdir -p gpurun_out/r2d
O=gpurun_out/r2d
QK_SKIP_C4_FULL=1 timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_gpu.log
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).readline())
    print(sys.argv[1].split('/')[-1], 'value %.4e ms %.1f e2e %.4e spread %.3f chunks %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['spread'], d['placement']['step_chunks']), d['placement']['state'], d['placement']['lanes_per_replication'], d['placement']['threads_per_block'])
except Exception as e:
    print(sys.argv[1], 'ERR', e)
PY
}
B="python bench.py --no-cpu-baseline --no-lockstep-probe --e2e-steps 1 --steps 2 --warmup 1"
for wl in haulage serial_line vector_flow container_haulage; do
  for cfg in "hbm 0" "group 8" "group 4" "group 16"; do
    set -- $cfg
    for lg in off ring; do
      QK_GROUP_LANES=$2 timeout 600 $B --workload $wl --reps 65536 --events 2000 --state $1 --log $lg > $O/sweep_${wl}_$1$2_$lg.json 2> $O/sweep_${wl}_$1$2_$lg.err
      show $O/sweep_${wl}_$1$2_$lg.json
    done
  done
done
for sl in 0 1; do
  QK_SLICE=$sl timeout 300 python bench.py --reps 131072 --no-cpu-baseline --no-lockstep-probe --e2e-steps 5 > $O/bench_c2_131072_slice$sl.json 2> $O/bench_c2_131072_slice$sl.err; show $O/bench_c2_131072_slice$sl.json
  QK_SLICE=$sl timeout 300 python bench.py --no-cpu-baseline --no-lockstep-probe --e2e-steps 5 > $O/bench_c2_full_slice$sl.json 2> $O/bench_c2_full_slice$sl.err; show $O/bench_c2_full_slice$sl.json
  QK_SLICE=$sl timeout 300 python bench.py --log off --no-cpu-baseline --no-lockstep-probe --e2e-steps 5 > $O/bench_c2_full_off_slice$sl.json 2> $O/bench_c2_full_off_slice$sl.err; show $O/bench_c2_full_off_slice$sl.json
done

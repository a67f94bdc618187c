#!/bin/bash
# round 2, session B: first run of the lane-group kernel.  Tests (under a timeout: a wrong mbarrier count would hang),
# then placement sweeps of the large models at reduced event counts, then full-size lines.
mkdir -p gpurun_out/r2b
O=gpurun_out/r2b
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "group or placement" > $O/pytest_group.log 2>&1; echo "pytest group rc=$?"; tail -15 $O/pytest_group.log
if ! grep -q " passed" $O/pytest_group.log || grep -q "failed" $O/pytest_group.log; then echo "group tests not green: stopping"; exit 1; fi
B="python bench.py --no-cpu-baseline --no-lockstep-probe --e2e-steps 1 --steps 2 --warmup 1"
for wl in haulage serial_line vector_flow; do
  for cfg in "hbm 0" "group 4" "group 8" "group 16"; do
    set -- $cfg
    for lg in off ring; do
      QK_GROUP_LANES=$2 timeout 600 $B --workload $wl --reps 65536 --events 2000 --state $1 --log $lg > $O/sweep_${wl}_$1$2_$lg.json 2> $O/sweep_${wl}_$1$2_$lg.err
      python - "$O/sweep_${wl}_$1$2_$lg.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).readline())
    print(sys.argv[1].split('/')[-1], 'value %.4e ms %.1f' % (d['value'], d['ms_per_step']), d['placement'])
except Exception as e:
    print(sys.argv[1], 'ERR', e)
PY
    done
  done
done
QK_SKIP_C4_FULL=1 timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest all rc=$?"; tail -5 $O/pytest_gpu.log

#!/bin/bash
# round 2, session A: the tree as of "strong scaling + packed reduction + wave-sliced launches", thread-per-replication
# kernels unchanged.  Tests, then the bench lines of C2 / C3 / C5 at BASELINE sizes (no profiler).
mkdir -p gpurun_out/r2a
O=gpurun_out/r2a
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $O/gpu.txt 2>&1
QK_SKIP_C4_FULL=${SKIP_C4:-1} timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_gpu.log
timeout 600 python bench.py > $O/bench_c2_ring.json 2> $O/bench_c2_ring.err; echo "bench c2 ring rc=$?"
timeout 600 python bench.py --log off --no-cpu-baseline > $O/bench_c2_off.json 2> $O/bench_c2_off.err; echo "bench c2 off rc=$?"
for wl in vector_flow haulage; do
  for lg in off ring; do
    timeout 900 python bench.py --workload $wl --log $lg $([ $lg = ring ] && echo --no-cpu-baseline) > $O/bench_${wl}_$lg.json 2> $O/bench_${wl}_$lg.err; echo "bench $wl $lg rc=$?"
  done
done
# strong-scaling shard shape of N = 8 on one GPU: 131,072 replications, with and without the wave-sliced schedule
for sl in 0 1; do
  QK_SLICE=$sl timeout 300 python bench.py --reps 131072 --no-cpu-baseline --no-lockstep-probe --e2e-steps 3 > $O/bench_c2_131072_slice$sl.json 2> $O/bench_c2_131072_slice$sl.err; echo "131072 slice=$sl rc=$?"
done
QK_SLICE=1 timeout 300 python bench.py --no-cpu-baseline --no-lockstep-probe --e2e-steps 3 > $O/bench_c2_full_slice1.json 2> $O/bench_c2_full_slice1.err; echo "full slice=1 rc=$?"
for f in $O/bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).readline())
    print(sys.argv[1].split('/')[-1], 'value %.4e ms %.1f e2e %.4e spread %.3f frac %.3f chunks %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['spread'], d['roofline']['frac'], d['placement']['step_chunks']), d['e2e']['phases_s_median_rank0'])
except Exception as e:
    print(sys.argv[1], 'ERR', e)
PY
done

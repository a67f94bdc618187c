#!/bin/bash
# round 2, session N (1 GPU): pending-emit mask in the two-tier queue; both tiers on chip (qhot) against split / qsplit
mkdir -p gpurun_out/r2n
O=gpurun_out/r2n
timeout 1200 python -m pytest tests -m gpu -x -q -k "split" > $O/pytest_split.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_split.log
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).readline())
    print(sys.argv[1].split('/')[-1], 'value %.4e ms %.1f e2e %.4e chunks %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['placement']['step_chunks']), d['placement']['state'], d['placement']['lanes_per_replication'], d['placement']['threads_per_block'], d['placement']['resident_blocks_per_sm'], d['roofline']['state_bytes_per_replication'])
except Exception as e:
    print(sys.argv[1], 'ERR', e)
PY
}
B="python bench.py --no-cpu-baseline --no-lockstep-probe --e2e-steps 1 --steps 2 --warmup 1"
for wl in haulage serial_line vector_flow container_haulage; do
  for lg in off ring; do
    for st in split qsplit qhot; do
      timeout 600 $B --workload $wl --reps 65536 --events 2000 --state $st --log $lg > $O/n_${wl}_${lg}_$st.json 2> $O/n_${wl}_${lg}_$st.err; show $O/n_${wl}_${lg}_$st.json
    done
  done
done
for tpb in 64 128 192; do
  timeout 600 $B --workload haulage --reps 65536 --events 2000 --state qsplit --threads-per-block $tpb --log off > $O/n_haulage_off_q$tpb.json 2> $O/n_haulage_off_q$tpb.err; show $O/n_haulage_off_q$tpb.json
done
timeout 600 $B --workload haulage --reps 65536 --events 2000 --state qsplit --log off --lib build/ab/lib_aos.so > $O/n_haulage_off_qsplit_aos.json 2> $O/n_haulage_off_qsplit_aos.err; show $O/n_haulage_off_qsplit_aos.json

#!/bin/bash
# round 2, session L (1 GPU): cold planes [replication][slot] under the split placements: parity, split vs two-tier, ncu
mkdir -p gpurun_out/r2l
O=gpurun_out/r2l
timeout 1200 python -m pytest tests -m gpu -x -q -k "split" > $O/pytest_split.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_split.log
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
for wl in haulage serial_line; do
  for lg in off ring; do
    for st in split qsplit; do
      timeout 600 $B --workload $wl --reps 65536 --events 2000 --state $st --log $lg > $O/a_${wl}_${lg}_$st.json 2> $O/a_${wl}_${lg}_$st.err; show $O/a_${wl}_${lg}_$st.json
    done
    for tpb in 128 256; do
      timeout 600 $B --workload $wl --reps 65536 --events 2000 --state qsplit --threads-per-block $tpb --log $lg > $O/a_${wl}_${lg}_q$tpb.json 2> $O/a_${wl}_${lg}_q$tpb.err; show $O/a_${wl}_${lg}_q$tpb.json
    done
  done
done
for wl in vector_flow container_haulage; do
  for st in auto split qsplit; do
    timeout 600 $B --workload $wl --reps 65536 --events 2000 --state $st --log off > $O/a_${wl}_$st.json 2> $O/a_${wl}_$st.err; show $O/a_${wl}_$st.json
  done
done
export QK_SLICE=0
for st in split qsplit; do
C="python bench.py --workload haulage --steps 1 --warmup 0 --reps 65536 --events 1000 --no-cpu-baseline --no-lockstep-probe --e2e-steps 1 --log off --state $st"
timeout 300 $C > $O/plain_haulage_$st.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:'qk_step_kernel' -s 2 -c 1 -o $O/prof_${st}_haulage -f $C > $O/ncu_haulage_$st.log 2>&1
tail -1 $O/ncu_haulage_$st.log
done

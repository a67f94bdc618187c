#!/bin/bash
# round 2, session O (1 GPU): group minima of the cold fire array + eager record loads in the delay-mode quick path ("new")
# against the build of the evidence run (build/ab/lib_ev2.so), one box
mkdir -p gpurun_out/r2o
O=gpurun_out/r2o
timeout 1200 python -m pytest tests -m gpu -x -q -k "split or placement" > $O/pytest_split.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_split.log
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
      for lib in new ev2; do
        L=""; [ $lib = ev2 ] && L="--lib build/ab/lib_ev2.so"
        timeout 600 $B --workload $wl --reps 131072 --events 2000 --state $st --log $lg $L > $O/o_${wl}_${lg}_${st}_$lib.json 2> $O/o_${wl}_${lg}_${st}_$lib.err; show $O/o_${wl}_${lg}_${st}_$lib.json
      done
    done
  done
done
for thr in 256 512 640; do
  for tpb in 128 256; do
    QK_QSPLIT_THREADS=$thr timeout 600 $B --workload haulage --reps 131072 --events 2000 --state qsplit --threads-per-block $tpb --log off > $O/o_haulage_off_q${thr}_$tpb.json 2> $O/o_haulage_off_q${thr}_$tpb.err; show $O/o_haulage_off_q${thr}_$tpb.json
  done
done
for thr in 256 320; do
  QK_QSPLIT_THREADS=$thr timeout 600 $B --workload serial_line --reps 131072 --events 2000 --state qsplit --log off > $O/o_serial_off_q${thr}.json 2> $O/o_serial_off_q${thr}.err; show $O/o_serial_off_q${thr}.json
done
for wl in vector_flow container_haulage; do
  for lib in new ev2; do
    L=""; [ $lib = ev2 ] && L="--lib build/ab/lib_ev2.so"
    timeout 600 $B --workload $wl --reps 131072 --events 2000 --log off $L > $O/o_${wl}_off_$lib.json 2> $O/o_${wl}_off_$lib.err; show $O/o_${wl}_off_$lib.json
  done
done
export QK_SLICE=0
C="python bench.py --workload haulage --steps 1 --warmup 0 --reps 65536 --events 1000 --no-cpu-baseline --no-lockstep-probe --e2e-steps 1 --log off"
timeout 300 $C > $O/plain_haulage.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:'qk_step_kernel' -s 2 -c 1 -o $O/prof_haulage_off -f $C > $O/ncu_haulage.log 2>&1
tail -1 $O/ncu_haulage.log

#!/bin/bash
# round 2, session M (1 GPU): A/B on one box: record prefetch ahead of REFILL (this tree) vs the build before it
# (build/ab/lib_aos.so); default placement of every workload; C2 lines of both builds
mkdir -p gpurun_out/r2m
O=gpurun_out/r2m
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
    for st in split qsplit hbm; do
      for lib in new aos; do
        L=""; [ $lib = aos ] && L="--lib build/ab/lib_aos.so"
        timeout 600 $B --workload $wl --reps 65536 --events 2000 --state $st --log $lg $L > $O/m_${wl}_${lg}_${st}_$lib.json 2> $O/m_${wl}_${lg}_${st}_$lib.err; show $O/m_${wl}_${lg}_${st}_$lib.json
      done
    done
    timeout 600 $B --workload $wl --reps 65536 --events 2000 --log $lg > $O/m_${wl}_${lg}_auto.json 2> $O/m_${wl}_${lg}_auto.err; show $O/m_${wl}_${lg}_auto.json
  done
done
for wl in vector_flow container_haulage; do
  for lg in off ring; do
    timeout 600 $B --workload $wl --reps 65536 --events 2000 --log $lg > $O/m_${wl}_${lg}_auto.json 2> $O/m_${wl}_${lg}_auto.err; show $O/m_${wl}_${lg}_auto.json
  done
done
for lib in new aos; do
  L=""; [ $lib = aos ] && L="--lib build/ab/lib_aos.so"
  for lg in ring off; do
    timeout 600 python bench.py --no-cpu-baseline --e2e-steps 2 --log $lg $L > $O/c2_${lib}_$lg.json 2> $O/c2_${lib}_$lg.err; show $O/c2_${lib}_$lg.json
  done
done

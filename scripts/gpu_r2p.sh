#!/bin/bash
# round 2, session P (1 GPU): draw counters left for the REFILL phase ("new") against the build before (build/ab/lib_gmin.so)
mkdir -p gpurun_out/r2p
O=gpurun_out/r2p
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
for wl in haulage serial_line vector_flow container_haulage; do
  for lg in off ring; do
    for lib in new gmin; do
      L=""; [ $lib = gmin ] && L="--lib build/ab/lib_gmin.so"
      timeout 600 $B --workload $wl --reps 131072 --events 2000 --log $lg $L > $O/p_${wl}_${lg}_$lib.json 2> $O/p_${wl}_${lg}_$lib.err; show $O/p_${wl}_${lg}_$lib.json
    done
  done
done
for lib in new gmin; do
  L=""; [ $lib = gmin ] && L="--lib build/ab/lib_gmin.so"
  timeout 600 $B --workload serial_line --reps 131072 --events 2000 --log off --state qsplit $L > $O/p_serial_off_qsplit_$lib.json 2> $O/p_serial_off_qsplit_$lib.err; show $O/p_serial_off_qsplit_$lib.json
  timeout 600 $B --workload haulage --reps 131072 --events 2000 --log off --state hbm $L > $O/p_haulage_off_hbm_$lib.json 2> $O/p_haulage_off_hbm_$lib.err; show $O/p_haulage_off_hbm_$lib.json
done

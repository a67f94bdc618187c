#!/bin/bash
# round 2, session I (1 GPU): memcheck of the split-placement fuzz failure; lane groups on top of the split placement.
mkdir -p gpurun_out/r2i
O=gpurun_out/r2i
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_fuzz_parity.py -x -q -m gpu -k "split_placement_on_gpu and 800" > $O/memcheck.log 2>&1; echo "memcheck rc=$?"; grep -m1 -B2 -A24 "Invalid\|Error" $O/memcheck.log | head -60
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
    for g in 2 4; do
      QK_GROUP_LANES=$g timeout 600 $B --workload $wl --reps 65536 --events 2000 --state gsplit --log $lg > $O/gs_${wl}_${lg}_g$g.json 2> $O/gs_${wl}_${lg}_g$g.err; show $O/gs_${wl}_${lg}_g$g.json
    done
  done
done
export QK_SLICE=0
C="python bench.py --workload haulage --steps 1 --warmup 0 --reps 65536 --events 1000 --no-cpu-baseline --no-lockstep-probe --e2e-steps 1 --log off --state gsplit"
QK_GROUP_LANES=2 timeout 300 $C > $O/plain_haulage_gsplit.log 2>&1 && QK_GROUP_LANES=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'qk_group_kernel' -s 2 -c 1 -o $O/prof_gsplit2_haulage -f $C > $O/ncu_haulage_gsplit.log 2>&1
tail -2 $O/ncu_haulage_gsplit.log

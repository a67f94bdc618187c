#!/bin/bash
# round 2, session J (1 GPU): which models fault under the split placement; record loads hoisted (split); C2 slicing A/B
mkdir -p gpurun_out/r2j
O=gpurun_out/r2j
timeout 900 python scripts/gpu_split_debug.py > $O/split_debug.log 2>&1; cat $O/split_debug.log | cut -c1-260
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
    timeout 600 $B --workload $wl --reps 65536 --events 2000 --state split --log $lg > $O/sp_${wl}_${lg}.json 2> $O/sp_${wl}_${lg}.err; show $O/sp_${wl}_${lg}.json
    timeout 600 $B --workload $wl --reps 65536 --events 2000 --state hbm --log $lg > $O/hbm_${wl}_${lg}.json 2> $O/hbm_${wl}_${lg}.err; show $O/hbm_${wl}_${lg}.json
  done
done
for sl in 0 d; do
  for lg in ring off; do
    QK_SLICE=$sl timeout 600 python bench.py --no-cpu-baseline --no-lockstep-probe --e2e-steps 2 --log $lg > $O/c2_slice${sl}_$lg.json 2> $O/c2_slice${sl}_$lg.err; show $O/c2_slice${sl}_$lg.json
  done
done
export QK_SLICE=0
C="python bench.py --workload haulage --steps 1 --warmup 0 --reps 65536 --events 1000 --no-cpu-baseline --no-lockstep-probe --e2e-steps 1 --log off --state split"
timeout 300 $C > $O/plain_haulage_split.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:'qk_step_kernel' -s 2 -c 1 -o $O/prof_split_haulage -f $C > $O/ncu_haulage_split.log 2>&1
tail -2 $O/ncu_haulage_split.log

#!/bin/bash
# round 2, session G (1 GPU): all GPU tests; C2 lines; placement probe of the large models (hbm / group / phased);
# ncu --set full of the group and the phased kernel on the haulage network.
mkdir -p gpurun_out/r2g
O=gpurun_out/r2g
timeout 2400 python -m pytest tests -m gpu -x -q --durations=8 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -14 $O/pytest_gpu.log
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).readline())
    k=d['roofline'].get('lockstep_K1')
    print(sys.argv[1].split('/')[-1], 'value %.4e ms %.1f e2e %.4e spread %.3f chunks %s frac %.4f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['spread'], d['placement']['step_chunks'], d['roofline']['frac']), d['placement']['state'], d['placement']['lanes_per_replication'], d['placement']['threads_per_block'], d['placement']['resident_blocks_per_sm'], ('K1 frac %.3f %.4f ms' % (k['frac'], k['kernel_ms'])) if k else '', 'cpu %.3e' % d.get('cpu_baseline',{}).get('value',0))
except Exception as e:
    print(sys.argv[1], 'ERR', e)
PY
}
for lg in ring off; do
  timeout 900 python bench.py --log $lg $([ $lg = off ] && echo --no-cpu-baseline) > $O/bench_c2_$lg.json 2> $O/bench_c2_$lg.err; show $O/bench_c2_$lg.json
done
B="python bench.py --no-cpu-baseline --no-lockstep-probe --e2e-steps 1 --steps 2 --warmup 1"
for wl in haulage serial_line; do
  for st in hbm group; do
    timeout 600 $B --workload $wl --reps 65536 --events 2000 --state $st --log off > $O/pr_${wl}_${st}_off.json 2> $O/pr_${wl}_${st}_off.err; show $O/pr_${wl}_${st}_off.json
  done
  for g in 8 4 16; do
    for nb in 1 2 4; do
      QK_GROUP_LANES=$g QK_PHASED_BLOCKS=$nb timeout 600 $B --workload $wl --reps 65536 --events 2000 --state phased --log off > $O/pr_${wl}_ph_g${g}_b${nb}_off.json 2> $O/pr_${wl}_ph_g${g}_b${nb}_off.err
      show $O/pr_${wl}_ph_g${g}_b${nb}_off.json
    done
  done
  timeout 600 $B --workload $wl --reps 65536 --events 2000 --state group --log ring > $O/pr_${wl}_group_ring.json 2> $O/pr_${wl}_group_ring.err; show $O/pr_${wl}_group_ring.json
  QK_GROUP_LANES=8 QK_PHASED_BLOCKS=2 timeout 600 $B --workload $wl --reps 65536 --events 2000 --state phased --log ring > $O/pr_${wl}_ph_g8_b2_ring.json 2> $O/pr_${wl}_ph_g8_b2_ring.err; show $O/pr_${wl}_ph_g8_b2_ring.json
done
export QK_SLICE=0
for st in group phased; do
  C="python bench.py --workload haulage --steps 1 --warmup 0 --reps 65536 --events 1000 --no-cpu-baseline --no-lockstep-probe --e2e-steps 1 --log off --state $st"
  timeout 300 $C > $O/plain_haulage_$st.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:'qk_(group|phased)_kernel' -s 2 -c 1 -o $O/prof_${st}_haulage -f $C > $O/ncu_haulage_$st.log 2>&1
  tail -2 $O/ncu_haulage_$st.log
done
ls -la $O | grep ncu-rep

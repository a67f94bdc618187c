#!/bin/bash
# Session r of round 2: staging by TMA tensor copies (UTMALDG / UTMASTG, one tile per plane and block) against the
# per-row bulk copies (QK_NO_TMA=1) of the same library: smoke, GPU tests, one-event launches, the bench's K.
O=gpurun_out/r2r; mkdir -p $O
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $O/smoke.log
[ -s $O/smoke.log ] && grep -q "smoke ok" $O/smoke.log || { echo "smoke failed: stopping"; exit 1; }
timeout 800 python -m pytest tests -m gpu -q -x --durations=5 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_gpu.log
echo "--- TMA tiles" | tee $O/k1_probe.txt; timeout 300 python scripts/gpu_k1_probe.py 2>&1 | tee -a $O/k1_probe.txt
echo "--- per-row bulk copies (QK_NO_TMA=1)" | tee -a $O/k1_probe.txt; QK_NO_TMA=1 timeout 300 python scripts/gpu_k1_probe.py 2>&1 | tee -a $O/k1_probe.txt
k1() { lg=$1; K=$2; ev=2048
  timeout 300 python bench.py --steps 2 --warmup 1 --events $ev --events-per-launch $K --no-cpu-baseline --e2e-steps 1 --log $lg 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); r=d['roofline']; print('value %.4e kernel_us %.1f achieved %.0f GB/s frac %.3f' % (d['value'], 1e3*r['kernel_ms'], r['achieved'], r['frac']))"; }
for rep in 1 2; do
  for lg in ring off; do for K in 1 2 4; do
    echo "== tma log=$lg K=$K $(k1 $lg $K)"
    echo "== rows log=$lg K=$K $(QK_NO_TMA=1 k1 $lg $K)"
  done; done
done 2>&1 | tee $O/ab_tma_k.txt
for mode in tma rows; do
  for lg in ring off; do
    [ $mode = rows ] && export QK_NO_TMA=1 || unset QK_NO_TMA
    echo "== $mode log=$lg $(timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --e2e-steps 1 --no-lockstep-probe --log $lg 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('value %.4e kernel_ms %.2f %s' % (d['value'], d['roofline']['kernel_ms'], d.get('placement')))")"
    echo "== $mode haulage log=$lg $(timeout 300 python bench.py --workload haulage --steps 3 --warmup 2 --no-cpu-baseline --e2e-steps 1 --no-lockstep-probe --log $lg 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('value %.4e kernel_ms %.2f %s' % (d['value'], d['roofline']['kernel_ms'], d.get('placement')))")"
  done
done 2>&1 | tee $O/ab_tma_full.txt
unset QK_NO_TMA

#!/bin/bash
# Session q of round 2: GPU tests of the tree with user-defined vector resources (feature set 3); A/B of the bulk-copy
# prologue (tables + state on one mbarrier) against the LDG table copy at K = 1, 2, 4 and at the bench's K; bench lines
# of the workloads whose code the change touched.
O=gpurun_out/r2q; mkdir -p $O
timeout 800 python -m pytest tests -m gpu -q -x --durations=8 > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_gpu.log
k1() { lib=$1; lg=$2; K=$3; ev=2048
  timeout 300 python bench.py --steps 2 --warmup 1 --events $ev --events-per-launch $K --no-cpu-baseline --e2e-steps 1 --log $lg --lib $lib 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); r=d['roofline']; print('value %.4e kernel_us %.1f achieved %.0f GB/s frac %.3f' % (d['value'], 1e3*r['kernel_ms'], r['achieved'], r['frac']))"; }
for rep in 1 2; do
for lib in quokkasim_b200/libquokka_b200.so build/libqk_ldg.so; do
  for lg in ring off; do for K in 1 2 4; do echo "== $lib log=$lg K=$K $(k1 $lib $lg $K)"; done; done
done; done 2>&1 | tee $O/ab_prologue_k.txt
for lib in quokkasim_b200/libquokka_b200.so build/libqk_ldg.so; do
  for lg in ring off; do
    echo "== $lib log=$lg $(timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --e2e-steps 1 --no-lockstep-probe --log $lg --lib $lib 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('value %.4e kernel_ms %.2f' % (d['value'], d['roofline']['kernel_ms']))")"
  done
done 2>&1 | tee $O/ab_prologue_full.txt
run() { name=$1; shift; timeout 900 python bench.py "$@" > $O/$name.json 2> $O/$name.err; echo "$name rc=$? $(head -c 400 $O/$name.json)"; }
run bench_iron_ore_flow_ring --workload iron_ore_flow --no-cpu-baseline --no-lockstep-probe --e2e-steps 2 --steps 3
run bench_iron_ore_flow_off --workload iron_ore_flow --log off --no-lockstep-probe --e2e-steps 2 --steps 3
run bench_c3_ring --workload vector_flow --no-cpu-baseline --no-lockstep-probe --e2e-steps 2 --steps 3

#!/bin/bash
# session C of round 1: full GPU test suite (incl. the container family), default bench lines, container workload
mkdir -p gpurun_out/c
O=gpurun_out/c
python -m pytest tests -m gpu -q -x > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
python bench.py --no-cpu-baseline > $O/bench_ring.json 2> $O/bench_ring.err; echo "bench ring rc=$?"; cut -c1-300 $O/bench_ring.json
python bench.py --no-cpu-baseline --log off > $O/bench_off.json 2> $O/bench_off.err; echo "bench off rc=$?"; cut -c1-300 $O/bench_off.json
WLS="container_haulage haulage serial_line vector_flow" bash scripts/gpu_workloads.sh

#!/bin/bash
# GPU session: parity tests, then the bench in both logging modes (numbers only from runs WITHOUT a profiler)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_ring.json 2> gpurun_out/bench_ring.err; echo "bench ring rc=$?"
python bench.py --log off --no-cpu-baseline > gpurun_out/bench_off.json 2> gpurun_out/bench_off.err; echo "bench off rc=$?"
cat gpurun_out/bench_ring.json gpurun_out/bench_off.json

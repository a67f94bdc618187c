#!/bin/bash
# Session s: the TMA-vs-row-copies equality test on its own (compute-sanitizer is closed on this pool).
O=gpurun_out/r2s; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "tma_tile" > $O/pytest_tma.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_tma.log

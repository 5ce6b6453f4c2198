#!/bin/bash
# Round 2, GPU call (1 GPU): overlapped sweeps on / off, symmetric suite with the C4 oracle result.
set -x
mkdir -p gpurun_out
QB_MAXIT=12 timeout 200 python tools/quick_bench.py 1e7 > gpurun_out/r2g_quick_overlap_n1e7.log 2>&1
SB200_OVERLAP=0 QB_MAXIT=12 timeout 200 python tools/quick_bench.py 1e7 > gpurun_out/r2g_quick_nooverlap_n1e7.log 2>&1
tail -n 2 gpurun_out/r2g_quick_overlap_n1e7.log gpurun_out/r2g_quick_nooverlap_n1e7.log | cut -c1-600
timeout 900 python -m pytest tests/test_gpu_sym.py -m gpu -q 2>&1 | tail -12 > gpurun_out/r2g_tests_sym.log
tail -3 gpurun_out/r2g_tests_sym.log

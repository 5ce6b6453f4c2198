#!/bin/bash
# Round 2, GPU call 1: run the opt-in (previously emulator-only) GPU tests without -x, then the SpMV layout study.
set -x
mkdir -p gpurun_out
nproc > gpurun_out/nproc.txt
SB200_EXPERIMENTAL=1 timeout 1200 python -m pytest tests/test_gpu_experimental.py -m gpu -q 2>&1 | tail -60 > gpurun_out/r2_experimental_tests.log
timeout 200 python tools/gather_roof.py > gpurun_out/r2_gather_roof.log 2>&1
timeout 400 python tools/spmv_roofline.py 1e7 > gpurun_out/r2_spmv_variants_n1e7.log 2>&1
timeout 200 python tools/spmv_roofline.py 1e6 > gpurun_out/r2_spmv_variants_n1e6.log 2>&1
QB_MAXIT=30 SB200_SPMV_FORMAT=sell timeout 400 python tools/quick_bench.py 1e7 > gpurun_out/r2_quick_sell_n1e7.log 2>&1
tail -5 gpurun_out/r2_experimental_tests.log

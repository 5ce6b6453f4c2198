#!/bin/bash
# Round 2, GPU call 2: parity of the fused schedule (operator kernel + first panel pass) on the device, then A/B timings.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_sym.py tests/test_gpu_gen.py -m gpu -q -x 2>&1 | tail -25 > gpurun_out/r2b_tests_sym_gen.log
tail -3 gpurun_out/r2b_tests_sym_gen.log
timeout 200 python tools/gather_roof.py > gpurun_out/r2b_gather_roof.log 2>&1
QB_MAXIT=30 timeout 300 python tools/quick_bench.py 1e7 > gpurun_out/r2b_quick_fused_n1e7.log 2>&1
SB200_FUSE_DOT=0 QB_MAXIT=30 timeout 300 python tools/quick_bench.py 1e7 > gpurun_out/r2b_quick_unfused_n1e7.log 2>&1
timeout 200 python tools/quick_bench.py 1e6 > gpurun_out/r2b_quick_fused_n1e6.log 2>&1
SB200_FUSE_DOT=0 timeout 200 python tools/quick_bench.py 1e6 > gpurun_out/r2b_quick_unfused_n1e6.log 2>&1
tail -2 gpurun_out/r2b_quick_fused_n1e7.log gpurun_out/r2b_quick_unfused_n1e7.log gpurun_out/r2b_quick_fused_n1e6.log

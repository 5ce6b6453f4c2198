#!/bin/bash
# Round 2, GPU call 5 (1 GPU): fused-kernel configurations (threads x loads in flight), restart kernel after the serial-path rewrite.
set -x
mkdir -p gpurun_out
for cfg in 512x4 512x8 256x8 256x16 256x16u8; do
  SB200_FUSED_CFG=$cfg QB_MAXIT=12 timeout 200 python tools/quick_bench.py 1e7 > gpurun_out/r2e_quick_${cfg}_n1e7.log 2>&1
done
for cfg in 512x4 256x16; do
  SB200_FUSED_CFG=$cfg timeout 200 python tools/quick_bench.py 1e6 > gpurun_out/r2e_quick_${cfg}_n1e6.log 2>&1
done
tail -n 1 gpurun_out/r2e_quick_*.log | cut -c1-900
timeout 600 python -m pytest tests/test_gpu_sym.py -m gpu -q -x -k "tridiag or restart or reference_cases or full_size" 2>&1 | tail -5 > gpurun_out/r2e_tests_subset.log
tail -3 gpurun_out/r2e_tests_subset.log

#!/bin/bash
# Round 2, GPU call (1 GPU): fused kernel with three resident CTAs per SM at per-rank row counts; sweep-mode agreement test.
set -x
mkdir -p gpurun_out
for n in 1.25e6 1e6 1e7; do
  for cfg in 512x8 512x4r3; do
    SB200_FUSED_CFG=$cfg QB_MAXIT=15 timeout 200 python tools/quick_bench.py $n > gpurun_out/r2i_quick_${cfg}_n${n}.log 2>&1
  done
done
tail -n 1 gpurun_out/r2i_quick_*.log | cut -c1-500
timeout 600 python -m pytest tests/test_gpu_sym.py tests/test_gpu_shift.py -m gpu -q -x -k "sweep_modes or shift_eigs_reference" 2>&1 | tail -5 > gpurun_out/r2i_tests.log
tail -3 gpurun_out/r2i_tests.log

#!/bin/bash
# Round 2, GPU call 4 (1 GPU): A/B of the fused-kernel variants and of sweep mode, parity of the touched tiers, source-level ncu of the restart kernel.
set -x
mkdir -p gpurun_out
for impl in cpasync reg tma; do
  SB200_FUSED_IMPL=$impl QB_MAXIT=12 QB_NOPROF=0 timeout 200 python tools/quick_bench.py 1e7 > gpurun_out/r2d_quick_${impl}_n1e7.log 2>&1
done
SB200_SWEEP=0 QB_MAXIT=12 timeout 200 python tools/quick_bench.py 1e7 > gpurun_out/r2d_quick_cpasync_nosweep_n1e7.log 2>&1
timeout 200 python tools/quick_bench.py 1e6 > gpurun_out/r2d_quick_n1e6.log 2>&1
SB200_SWEEP=0 timeout 200 python tools/quick_bench.py 1e6 > gpurun_out/r2d_quick_nosweep_n1e6.log 2>&1
SB200_FUSED_IMPL=reg timeout 200 python tools/quick_bench.py 1e6 > gpurun_out/r2d_quick_reg_n1e6.log 2>&1
tail -n 1 gpurun_out/r2d_quick_*.log | cut -c1-700
timeout 1500 python -m pytest tests/test_gpu_sym.py tests/test_gpu_gen.py -m gpu -q 2>&1 | tail -15 > gpurun_out/r2d_tests_sym_gen.log
tail -3 gpurun_out/r2d_tests_sym_gen.log
QB_NOPROF=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:sym_restart_kernel -s 2 -c 1 -o gpurun_out/r2d_restart python tools/quick_bench.py 1e6 > gpurun_out/r2d_ncu_restart.log 2>&1
ls -la gpurun_out/*.ncu-rep

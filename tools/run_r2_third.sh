#!/bin/bash
# Round 2, GPU call 3 (1 GPU): TMA-fed fused kernel: parity of the Gen / shift / layout tiers, then timings.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_gen.py tests/test_gpu_shift.py tests/test_cpp_shim.py -m gpu -q 2>&1 | tail -25 > gpurun_out/r2c_tests_gen_shift.log
tail -3 gpurun_out/r2c_tests_gen_shift.log
timeout 600 python -m pytest tests/test_gpu_sym.py -m gpu -q -x -k "factorization or reference_cases or full_size or golden or givens" 2>&1 | tail -8 > gpurun_out/r2c_tests_sym_subset.log
tail -3 gpurun_out/r2c_tests_sym_subset.log
QB_MAXIT=30 timeout 300 python tools/quick_bench.py 1e7 > gpurun_out/r2c_quick_fused_n1e7.log 2>&1
timeout 200 python tools/quick_bench.py 1e6 > gpurun_out/r2c_quick_fused_n1e6.log 2>&1
tail -n 2 gpurun_out/r2c_quick_fused_n1e7.log gpurun_out/r2c_quick_fused_n1e6.log

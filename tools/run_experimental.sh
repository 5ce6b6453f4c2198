#!/bin/bash
# First GPU call of the next round: verifies the code written after round 1's GPU budget was spent, then measures it.
#   gpurun --timeout 1500 -- 'bash tools/run_experimental.sh'
set -x
mkdir -p gpurun_out
SB200_EXPERIMENTAL=1 timeout 900 python -m pytest tests/test_gpu_experimental.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/experimental_tests.log
timeout 300 python tools/gather_roof.py 2>&1 | tee gpurun_out/gather_roof.log
timeout 600 python tools/spmv_roofline.py 1e7 2>&1 | tee gpurun_out/spmv_variants_n1e7.log
timeout 300 python tools/spmv_roofline.py 1e6 2>&1 | tee gpurun_out/spmv_variants_n1e6.log
SB200_SPMV_FORMAT=sell timeout 600 python tools/quick_bench.py 1e7 2>&1 | tee gpurun_out/quick_sell_n1e7.log
# complex Hermitian path: one mid-size solve for a first timing (SparseHermMatProd + HermEigsSolver)
timeout 300 python tools/herm_probe.py 2>&1 | tee gpurun_out/herm_probe.log
# everything above in one JSON line (what bench.py attaches under "experimental")
timeout 600 python tools/experimental_probe.py 1e7 2>&1 | tail -1 | tee gpurun_out/experimental_probe.json

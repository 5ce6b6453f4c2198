set -x
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 300 python -m pytest tests/test_gpu_sym.py -m gpu -q --timeout 120 -k "restart_gemm" > gpurun_out/pytest_gemm.log 2>&1
tail -c 1500 gpurun_out/pytest_gemm.log
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 -x > gpurun_out/pytest_all.log 2>&1
tail -c 1500 gpurun_out/pytest_all.log
( timeout 200 python tools/quick_bench.py 1e6 ) > gpurun_out/qb10_1e6.log 2>&1; cat gpurun_out/qb10_1e6.log | cut -c1-800
( QB_MAXIT=8 timeout 300 python tools/quick_bench.py 1e7 ) > gpurun_out/qb10_1e7.log 2>&1; cat gpurun_out/qb10_1e7.log | cut -c1-800
( SB200_COMPRESS_FMA=1 QB_MAXIT=8 timeout 300 python tools/quick_bench.py 1e7 ) > gpurun_out/qb10_1e7_fma.log 2>&1; tail -n 1 gpurun_out/qb10_1e7_fma.log | cut -c1-800
( timeout 300 python tools/spmv_roofline.py 1e7 ) > gpurun_out/spmv_roofline.log 2>&1; cat gpurun_out/spmv_roofline.log | cut -c1-600

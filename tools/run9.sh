set -x
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_all.log 2>&1
tail -c 1800 gpurun_out/pytest_all.log
( timeout 200 python tools/quick_bench.py 1e6 ) > gpurun_out/qb9_1e6.log 2>&1; cat gpurun_out/qb9_1e6.log | cut -c1-800
( QB_MAXIT=8 timeout 300 python tools/quick_bench.py 1e7 ) > gpurun_out/qb9_1e7.log 2>&1; cat gpurun_out/qb9_1e7.log | cut -c1-800

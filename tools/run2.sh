set -x
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -80 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 600 python tools/quick_bench.py 1e6 > gpurun_out/quick_1e6.log 2>&1; cat gpurun_out/quick_1e6.log

set -x
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 900 python -m pytest tests/test_gpu_sym.py tests/test_cpp_shim.py -m gpu -q -s --timeout 200 > gpurun_out/pytest_sym.log 2>&1
tail -c 6000 gpurun_out/pytest_sym.log
timeout 300 python tools/quick_bench.py 1e6 > gpurun_out/quick_1e6.log 2>&1; cat gpurun_out/quick_1e6.log

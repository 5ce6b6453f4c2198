set -x
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 900 python -m pytest tests/test_gpu_shift.py tests/test_cpp_shim.py -m gpu -q --timeout 200 > gpurun_out/pytest_shift.log 2>&1
tail -c 1500 gpurun_out/pytest_shift.log
( timeout 300 python tools/shift_bench.py 2e5 ) > gpurun_out/shift_bench.log 2>&1; cat gpurun_out/shift_bench.log | cut -c1-700

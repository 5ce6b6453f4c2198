set -x
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvsmi.txt 2>&1; nproc > gpurun_out/nproc.txt; lscpu | head -20 >> gpurun_out/nproc.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 600 python tools/quick_bench.py 1e6 > gpurun_out/quick_1e6.log 2>&1; cat gpurun_out/quick_1e6.log

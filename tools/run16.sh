set -x
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 600 python -m pytest tests/test_gpu_sym.py -m gpu -q --timeout 200 -k "chunked or column_blocked or operator" > gpurun_out/pytest_chunk.log 2>&1
tail -c 2500 gpurun_out/pytest_chunk.log
N=2
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_check.py > gpurun_out/mgpu${N}_check.log 2>&1
grep -E "PASS|FAIL|Error|error" gpurun_out/mgpu${N}_check.log | cut -c1-330 | head -12
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 2 --warmup 1 --skip-cpu-baseline > gpurun_out/bench_n${N}.json 2> gpurun_out/bench_n${N}.err
tail -c 3000 gpurun_out/bench_n${N}.json; tail -3 gpurun_out/bench_n${N}.err

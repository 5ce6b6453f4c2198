set -x
mkdir -p gpurun_out
export PYTHONPATH=$PWD
nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_check.py > gpurun_out/mgpu2_check.log 2>&1
grep -E "PASS|FAIL|Error|error" gpurun_out/mgpu2_check.log | cut -c1-500; tail -5 gpurun_out/mgpu2_check.log | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 1 --warmup 1 --skip-cpu-baseline > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
tail -c 2500 gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err | cut -c1-300

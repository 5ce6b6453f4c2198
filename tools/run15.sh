set -x
mkdir -p gpurun_out
export PYTHONPATH=$PWD
N=${NG:-8}
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/mgpu_check.py > gpurun_out/mgpu${N}_check.log 2>&1
grep -E "PASS|FAIL|Error|error" gpurun_out/mgpu${N}_check.log | cut -c1-400 | head -30
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 2 --warmup 1 --skip-cpu-baseline > gpurun_out/bench_n${N}.json 2> gpurun_out/bench_n${N}.err
tail -c 3000 gpurun_out/bench_n${N}.json; tail -5 gpurun_out/bench_n${N}.err

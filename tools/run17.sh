set -x
mkdir -p gpurun_out
export PYTHONPATH=$PWD
N=${NG:-4}
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/mgpu_chunks.py 1e7 12 > gpurun_out/mgpu${N}_chunks.log 2>&1
grep -E "^\{|Error|error" gpurun_out/mgpu${N}_chunks.log | cut -c1-500

#!/bin/bash
# Round 2, second 8-GPU call (final code): truncated-solve timing in peer mode and the contract bench line at N = 8.
NP=${NP:-8}
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1"
PEER_LIST=1 CHUNK_LIST=2 timeout 300 $TR --master-port 29513 tools/mgpu_chunks.py 1e7 20 > gpurun_out/r2b_mgpu${NP}_chunks.log 2>&1
grep "^{" gpurun_out/r2b_mgpu${NP}_chunks.log | cut -c1-500
timeout 500 $TR --master-port 29515 bench.py --gpus $NP --steps 2 --warmup 1 > gpurun_out/r2b_bench_n${NP}.json 2> gpurun_out/r2b_bench_n${NP}.err
tail -c 300 gpurun_out/r2b_bench_n${NP}.json

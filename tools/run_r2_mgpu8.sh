#!/bin/bash
# Round 2, 8-GPU call: parity in peer-memory mode, truncated-solve timing peer vs NCCL, the contract bench line at N = 8.
NP=${NP:-8}
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29511 tools/mgpu_check.py > gpurun_out/r2_mgpu${NP}_check_peer.log 2>&1
grep -c PASS gpurun_out/r2_mgpu${NP}_check_peer.log; grep FAIL gpurun_out/r2_mgpu${NP}_check_peer.log | head -3
PEER_LIST=1,0 CHUNK_LIST=2 timeout 400 $TR --master-port 29513 tools/mgpu_chunks.py 1e7 20 > gpurun_out/r2_mgpu${NP}_chunks.log 2>&1
grep "^{" gpurun_out/r2_mgpu${NP}_chunks.log | cut -c1-500
timeout 600 $TR --master-port 29515 bench.py --gpus $NP --steps 2 --warmup 1 > gpurun_out/r2_bench_n${NP}.json 2> gpurun_out/r2_bench_n${NP}.err
tail -c 400 gpurun_out/r2_bench_n${NP}.json

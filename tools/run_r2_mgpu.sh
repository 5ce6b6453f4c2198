#!/bin/bash
# Round 2, multi-GPU call: NP = number of GPUs (gpurun --gpus NP).  Parity of the row-sharded solvers in peer-memory mode and with the
# NCCL collectives, then the truncated-solve timing of both.
NP=${NP:-2}
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tools/mgpu_check.py > gpurun_out/r2_mgpu${NP}_check_peer.log 2>&1
SB200_PEER=0 timeout 600 $TR --master-port 29512 tools/mgpu_check.py > gpurun_out/r2_mgpu${NP}_check_nccl.log 2>&1
grep -c PASS gpurun_out/r2_mgpu${NP}_check_peer.log gpurun_out/r2_mgpu${NP}_check_nccl.log
grep FAIL gpurun_out/r2_mgpu${NP}_check_peer.log gpurun_out/r2_mgpu${NP}_check_nccl.log | head -5
PEER_LIST=1,0 CHUNK_LIST=2 timeout 600 $TR --master-port 29513 tools/mgpu_chunks.py 1e7 ${MAXIT:-15} > gpurun_out/r2_mgpu${NP}_chunks.log 2>&1
tail -n 4 gpurun_out/r2_mgpu${NP}_chunks.log | cut -c1-600

#!/bin/bash
# Round 2, GPU call (1 GPU): where to cut the operand between the head kernel and the fused kernel (equal halves / uneven / one block).
set -x
mkdir -p gpurun_out
QB_MAXIT=12 QB_NOPROF=1 timeout 200 python tools/quick_bench.py 1e7 > gpurun_out/r2h_split_equal.log 2>&1
for mb in 10 20 30; do
  SB200_XSPLIT0_MB=$mb QB_MAXIT=12 QB_NOPROF=1 timeout 200 python tools/quick_bench.py 1e7 > gpurun_out/r2h_split_${mb}mb.log 2>&1
done
SB200_XSLICE_MB=100 QB_MAXIT=12 QB_NOPROF=1 timeout 200 python tools/quick_bench.py 1e7 > gpurun_out/r2h_split_oneblock.log 2>&1
SB200_XSLICE_MB=28 QB_MAXIT=12 QB_NOPROF=1 timeout 200 python tools/quick_bench.py 1e7 > gpurun_out/r2h_split_threeblocks.log 2>&1
tail -n 1 gpurun_out/r2h_split_*.log | cut -c1-400

#!/bin/bash
# ncu captures for the sliced-layout SpMV (next round, after tools/run_experimental.sh shows it correct):
#   gpurun --timeout 1500 -- 'bash tools/run_ncu_sell.sh'
# 1) launch list of a truncated solve with per-kernel durations, 2) one --set full capture of the sliced kernels (plain + fused step).
set -x
mkdir -p gpurun_out
export SB200_SPMV_FORMAT=sell
export QB_MAXIT=3 QB_NOPROF=1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/sell_launches.csv python tools/quick_bench.py 1e7 > gpurun_out/sell_quick.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sell_ -c 6 -o gpurun_out/sell_full python tools/quick_bench.py 1e7 > gpurun_out/sell_full.log 2>&1
python tools/ncu_summary.py launches gpurun_out/sell_launches.csv gpurun_out/sell_launches.md || true
python tools/ncu_summary.py full gpurun_out/sell_full.ncu-rep gpurun_out/sell_full.md || true

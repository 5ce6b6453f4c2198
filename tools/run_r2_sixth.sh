#!/bin/bash
# Round 2, GPU call (1 GPU): the whole GPU suite, the contract bench line, the ncu launch list of the bench command and full captures of
# the three hot kernels (fused operator kernel, plain operator kernel, correction pass).
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r2f_tests_all.log
tail -3 gpurun_out/r2f_tests_all.log
timeout 1500 python bench.py --gpus 1 --steps 2 --warmup 3 > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err
tail -c 600 gpurun_out/r2f_bench_n1.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 3000 --csv --log-file gpurun_out/r2f_launches.csv \
    python bench.py --gpus 1 --steps 1 --warmup 0 --skip-e2e --skip-cpu-baseline --skip-configs > gpurun_out/r2f_bench_under_ncu.log 2>&1
QB_MAXIT=2 QB_NOPROF=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sell_step_dot_kernel|sell_plain_kernel|panel_kernel" -s 70 -c 8 \
    -o gpurun_out/r2f_hot python tools/quick_bench.py 1e7 > gpurun_out/r2f_ncu_hot.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/r2f_launches.csv

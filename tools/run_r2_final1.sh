#!/bin/bash
# Round 2, final 1-GPU call: whole GPU suite, contract bench line (default flags), ncu launch list of the bench command.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r2z_tests_all.log
tail -3 gpurun_out/r2z_tests_all.log
timeout 900 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.log 2>&1; tail -1 gpurun_out/r2z_smoke.log
timeout 1500 python bench.py > gpurun_out/r2z_bench_n1.json 2> gpurun_out/r2z_bench_n1.err
tail -c 300 gpurun_out/r2z_bench_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 3000 --csv --log-file gpurun_out/r2z_launches.csv \
    python bench.py --gpus 1 --steps 1 --warmup 0 --skip-e2e --skip-cpu-baseline --skip-configs > gpurun_out/r2z_bench_under_ncu.log 2>&1
ls -la gpurun_out/r2z_*

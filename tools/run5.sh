set -x
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 900 python bench.py --gpus 1 --steps 3 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 3000 gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
# ncu launch list of a short profiling run of the same command line (1 step)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 2500 --csv --log-file gpurun_out/launches_r1.csv python bench.py --gpus 1 --steps 1 --warmup 0 --skip-e2e --skip-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
tail -3 gpurun_out/bench_ncu.log
# full-set capture of the two hot kernels
timeout 600 ncu --set full --clock-control none --import-source on -k regex:panel_kernel -s 120 -c 2 -o gpurun_out/prof_panel_r1 python tools/quick_bench.py 1e7 > gpurun_out/ncu_panel.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmv_step_kernel -s 60 -c 2 -o gpurun_out/prof_spmv_r1 python tools/quick_bench.py 1e7 > gpurun_out/ncu_spmv.log 2>&1
ls -la gpurun_out

set -x
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 300 python -m pytest tests/test_gpu_sym.py -m gpu -q --timeout 120 -k "restart_gemm" > gpurun_out/pytest_gemm.log 2>&1
tail -c 600 gpurun_out/pytest_gemm.log
( QB_MAXIT=8 timeout 300 python tools/quick_bench.py 1e7 ) > gpurun_out/qb11_1e7.log 2>&1; tail -n 1 gpurun_out/qb11_1e7.log | cut -c1-800
# official bench, N=1
timeout 900 python bench.py --gpus 1 --steps 3 --warmup 3 > gpurun_out/bench_n1_b.json 2> gpurun_out/bench_n1_b.err; tail -c 3500 gpurun_out/bench_n1_b.json; tail -3 gpurun_out/bench_n1_b.err
timeout 600 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > gpurun_out/bench_ref_b.json 2> gpurun_out/bench_ref_b.err; tail -c 1500 gpurun_out/bench_ref_b.json
# ncu launch list of the same command (one solve)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 3000 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --gpus 1 --steps 1 --warmup 0 --skip-e2e --skip-cpu-baseline > gpurun_out/bench_ncu_b.log 2>&1
tail -2 gpurun_out/bench_ncu_b.log
# full-set captures: both column-block kernels of two operator applications, two panel passes, one restart GEMM
( export QB_MAXIT=2 QB_NOPROF=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmv_ -s 40 -c 4 -o gpurun_out/prof_spmv_r1b python tools/quick_bench.py 1e7 > gpurun_out/ncu_spmv_b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:panel_kernel -s 120 -c 2 -o gpurun_out/prof_panel_r1b python tools/quick_bench.py 1e7 > gpurun_out/ncu_panel_b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:compress_dmma -c 1 -o gpurun_out/prof_gemm_r1b python tools/quick_bench.py 1e7 > gpurun_out/ncu_gemm_b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sym_restart_kernel -c 1 -o gpurun_out/prof_restart_r1b python tools/quick_bench.py 1e7 > gpurun_out/ncu_restart_b.log 2>&1
)
ls -la gpurun_out | tail -12

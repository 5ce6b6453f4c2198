set -x
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 600 python -m pytest tests -m gpu -q --timeout 300 -x > gpurun_out/pytest_final.log 2>&1
tail -c 600 gpurun_out/pytest_final.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 300 python bench.py --gpus 1 --steps 1 --warmup 1 --skip-cpu-baseline > gpurun_out/bench_n1_c.json 2> gpurun_out/bench_n1_c.err; tail -c 2500 gpurun_out/bench_n1_c.json

set -x
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 1500 python -m pytest tests -m gpu -q --timeout 400 > gpurun_out/pytest_all.log 2>&1
tail -c 1500 gpurun_out/pytest_all.log
# ncu evidence for the shift-solve kernels (small problem, cheap)
( export SB_NO_CPU=1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_shift.csv python tools/shift_bench.py 2e5 > gpurun_out/ncu_shift_list.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bcr_ -s 60 -c 24 -o gpurun_out/prof_bcr_r1b python tools/shift_bench.py 2e5 > gpurun_out/ncu_bcr.log 2>&1
)
ls -la gpurun_out | tail -5

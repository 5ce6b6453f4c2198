set -x
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_all.log 2>&1
tail -c 2500 gpurun_out/pytest_all.log
( timeout 200 python tools/quick_bench.py 1e6 ) > gpurun_out/qb_1e6.log 2>&1; cat gpurun_out/qb_1e6.log
for xs in 24 40 16; do
  ( QB_MAXIT=8 SB200_XSLICE_MB=$xs timeout 300 python tools/quick_bench.py 1e7 ) > gpurun_out/qb_1e7_xs$xs.log 2>&1; cat gpurun_out/qb_1e7_xs$xs.log
done
( QB_MAXIT=8 SB200_XSLICE_MB=24 SB200_SPMV_LANES=4 timeout 300 python tools/quick_bench.py 1e7 ) > gpurun_out/qb_1e7_xs24_l4.log 2>&1; cat gpurun_out/qb_1e7_xs24_l4.log

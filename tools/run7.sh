set -x
mkdir -p gpurun_out
export PYTHONPATH=$PWD
for suf in "" _g1 _g2; do
  ( SB200_LIB_SUFFIX=$suf QB_MAXIT=8 timeout 300 python tools/quick_bench.py 1e7 ) > gpurun_out/qb7_1e7$suf.log 2>&1; cat gpurun_out/qb7_1e7$suf.log | cut -c1-700
  ( SB200_LIB_SUFFIX=$suf QB_MAXIT=20 QB_NOPROF=1 timeout 300 python tools/quick_bench.py 1e6 ) > gpurun_out/qb7_1e6$suf.log 2>&1; cat gpurun_out/qb7_1e6$suf.log | cut -c1-400
done
( SB200_XSLICE_MB=60 QB_MAXIT=8 timeout 300 python tools/quick_bench.py 1e7 ) > gpurun_out/qb7_1e7_xs60.log 2>&1; cat gpurun_out/qb7_1e7_xs60.log | cut -c1-700

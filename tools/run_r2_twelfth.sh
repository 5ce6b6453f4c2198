#!/bin/bash
# Round 2, last GPU call: the wide-band shift-solve after the CPU-side rework (column-split block products replayed from a CUDA graph,
# blocked Gauss-Jordan, automatic refinement) -- tests in the default configuration and with the fallbacks forced, then the timings.
set -x
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_shift.py -m gpu -q -x -k "(mesh or errors or banded or reference_cases) and not full_size" --durations=5 2>&1 | tail -12 > gpurun_out/r2ac_tests_shift_default.log
cat gpurun_out/r2ac_tests_shift_default.log
SB200_SHIFT_GJ=rank1 SB200_SHIFT_GRAPH=0 timeout 120 python -m pytest tests/test_gpu_shift.py -m gpu -q -x -k "mesh and not full_size and not dims6" 2>&1 | tail -4 > gpurun_out/r2ac_tests_shift_fallbacks.log
cat gpurun_out/r2ac_tests_shift_fallbacks.log
timeout 150 python tools/mesh_shift_bench.py 58x58x58 400x500 30x30x30 > gpurun_out/r2ac_mesh_shift_bench.log 2>&1
tail -8 gpurun_out/r2ac_mesh_shift_bench.log
SB200_SHIFT_GRAPH=0 timeout 60 python tools/mesh_shift_bench.py 400x500 > gpurun_out/r2ac_mesh_shift_bench_nograph.log 2>&1
tail -3 gpurun_out/r2ac_mesh_shift_bench_nograph.log

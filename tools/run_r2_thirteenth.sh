#!/bin/bash
# Round 2, remaining GPU seconds: the wide-band shift-solve with band-widths merged into ~1000-row blocks.
set -x
mkdir -p gpurun_out
timeout 60 python tools/mesh_shift_bench.py 400x500 > gpurun_out/r2ad_mesh_shift_bench_merged.log 2>&1
tail -3 gpurun_out/r2ad_mesh_shift_bench_merged.log
timeout 70 python -m pytest tests/test_gpu_shift.py -m gpu -q -x -k "mesh and not full_size and not dims5" 2>&1 | tail -4 > gpurun_out/r2ad_tests_mesh_merged.log
cat gpurun_out/r2ad_tests_mesh_merged.log

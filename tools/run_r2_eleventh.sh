#!/bin/bash
# Round 2: device check of the wide-band shift-solve route (mesh-like patterns) + its timings.
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_shift.py -m gpu -q -x -k "mesh or errors" --durations=8 2>&1 | tail -16 > gpurun_out/r2ab_tests_mesh.log
cat gpurun_out/r2ab_tests_mesh.log
timeout 200 python tools/mesh_shift_bench.py 58x58x58 400x500 30x30x30 > gpurun_out/r2ab_mesh_shift_bench.log 2>&1
cat gpurun_out/r2ab_mesh_shift_bench.log | tail -8

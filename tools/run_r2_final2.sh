#!/bin/bash
# Round 2: device check of the Arnoldi sweep mode (nonsymmetric tiers + the C3 throughput line) and of the tests added after the full-suite run.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gen.py tests/test_gpu_shift.py tests/test_gpu_sym.py -m gpu -q -x -k "gen or rejects or sweep_modes or restart_gemm or compress" 2>&1 | tail -6 > gpurun_out/r2y_tests_gen.log
tail -3 gpurun_out/r2y_tests_gen.log
timeout 300 python - > gpurun_out/r2y_c3.log 2>&1 <<'PY'
import json, sys
sys.path.insert(0, '.')
import bench, spectra_b200 as sb
from spectra_b200 import synth
print(json.dumps(bench.run_other_configs(sb, synth)))
PY
tail -c 900 gpurun_out/r2y_c3.log

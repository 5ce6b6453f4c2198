#!/bin/bash
# Round 2, last GPU call (12 box-minutes left): device check of the tests written in the CPU-only session -- GPU vs the reference's own code
# (oracle/_ref travels with the snapshot), the Eigen-interop program against the CUDA library -- then smoke() and the reference arm of bench.py.
set -x
mkdir -p gpurun_out
ls -la oracle/_ref oracle/_build spectra_b200/lib 2>&1 | tail -8
timeout 420 python -m pytest tests/test_gpu_vs_reference.py tests/test_cpp_shim.py -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r2aa_tests_vs_reference.log
cat gpurun_out/r2aa_tests_vs_reference.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2aa_smoke.log 2>&1
tail -2 gpurun_out/r2aa_smoke.log
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r2aa_bench_reference.json 2> gpurun_out/r2aa_bench_reference.err
tail -c 1500 gpurun_out/r2aa_bench_reference.json
timeout 240 python -m pytest tests/test_gpu_sym.py -m gpu -q -x -k "full_size or c4" 2>&1 | tail -4 > gpurun_out/r2aa_tests_fullsize.log
cat gpurun_out/r2aa_tests_fullsize.log

set -x
mkdir -p gpurun_out
export PYTHONPATH=$PWD
timeout 400 python -m pytest tests/test_gpu_sym.py -m gpu -q -s --timeout 60 -k "tridiag_qr or restart_step or lanczos_factorization" > gpurun_out/pytest_dbg.log 2>&1
tail -c 30000 gpurun_out/pytest_dbg.log

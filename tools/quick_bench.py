"""Developer probe (not the contract bench): SpMV + one full solve with per-kernel-class timing."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np

import spectra_b200 as sb
from spectra_b200 import synth

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
nev, ncv = 20, 60
t = time.time()
rp, ci, v = synth.csr(n, 20, 0, True)
print("gen", round(time.time() - t, 2), "s nnz", len(ci), flush=True)
t = time.time()
op = sb.SparseGenMatProd.from_csr_slab(n, 0, rp, ci, v)
print("upload", round(time.time() - t, 2), "s", flush=True)
ms = op.spmv_device_time(20)
bytes_spmv = 12 * len(ci) + 4 * (n + 1) + 16 * n
print(json.dumps(dict(kernel="spmv_plain", n=n, ms=ms, gbs=bytes_spmv / ms / 1e6, frac=bytes_spmv / ms / 1e6 / 6571.2)), flush=True)
for prof in (0, 1):
    sb.set_profiling(prof)
    eigs = sb.SymEigsSolver(op, nev, ncv)
    t = time.time()
    eigs.init()
    nconv = eigs.compute(sb.SortRule.LargestAlge)
    wall = time.time() - t
    st = eigs.stats()
    st.update(profiling=prof, wall_s=wall, nconv=nconv, nops=eigs.num_operations(), niter=eigs.num_iterations(), info=int(eigs.info()))
    if prof:
        pb = 8 * n * (st["panel_cols"] + 2 * st["panel_launches"])
        st["panel_gbs"] = pb / st["ms_panel"] / 1e6
        st["spmv_gbs"] = (bytes_spmv + 16 * n) * st["spmv_launches"] / st["ms_spmv"] / 1e6
        st["avg_panel_ms"] = st["ms_panel"] / st["panel_launches"]
        st["avg_spmv_ms"] = st["ms_spmv"] / st["spmv_launches"]
    print(json.dumps(st), flush=True)

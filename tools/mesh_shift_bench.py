"""Developer probe of the wide-band shift-solve route (band_solve.cu, factor_thomas): stencil matrices in natural ordering.
usage: python tools/mesh_shift_bench.py 58x58x58 [400x500 ...]   (27-point stencil for 3-D sizes, 5-point for 2-D)"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import scipy.sparse as sp

import spectra_b200 as sb
from helpers import stencil_matrix

for spec in sys.argv[1:] or ["58x58x58"]:
    dims = tuple(int(t) for t in spec.split("x"))
    A = stencil_matrix(dims, full=len(dims) == 3, seed=5)
    n = A.shape[0]
    t = time.time()
    op = sb.SparseSymShiftSolve(sp.tril(A).tocsc())
    t_up = time.time() - t
    t = time.time()
    op.set_shift(0.5)
    t_fac = time.time() - t
    ms = op.solve_device_time(5)
    op.set_refine(0)
    ms0 = op.solve_device_time(5)
    op.set_refine(1)
    lay = op.layout()
    gb = 3 * lay["block_rows"] * lay["block"] ** 2 * 8 / 1e9
    print(json.dumps(dict(dims=dims, n=n, nnz=int(A.nnz), layout=lay, factors_gb=round(gb, 2), upload_s=round(t_up, 3), set_shift_s=round(t_fac, 3), solve_ms=round(ms, 3),
                          solve_ms_no_refine=round(ms0, 3), stream_gbps_no_refine=round(gb / (ms0 * 1e-3), 1))), flush=True)
    eigs = sb.SymEigsShiftSolver(op, 10, 30, 0.5)
    t = time.time()
    eigs.init()
    nconv = eigs.compute(sb.SortRule.LargestMagn)
    wall = time.time() - t
    ev, X = eigs.eigenvalues(), eigs.eigenvectors()
    res = float((np.linalg.norm(A @ X - X * ev, axis=0) / np.abs(ev)).max())
    print(json.dumps(dict(dims=dims, solve_wall_s=round(wall, 3), nconv=nconv, nops=eigs.num_operations(), niter=eigs.num_iterations(), info=int(eigs.info()), max_rel_res=res)),
          flush=True)
    del eigs, op

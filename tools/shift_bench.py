"""Developer probe for BASELINE config 5 (shift-and-invert): SymEigsShiftSolver + SparseSymShiftSolve on G_band(n, b = 15), sigma = 0.5,
k = 10, ncv = 30, next to the CPU oracle (band LU + the same driver, 1 thread).  usage: python tools/shift_bench.py [n]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import scipy.sparse as sp

import spectra_b200 as sb
from spectra_b200 import synth

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 200_000
b, sigma, nev, ncv = 15, 0.5, 10, 30
rp, ci, v = synth.band_csr(n, b, 0, 0.0)
A = sp.csr_matrix((v, ci, rp), shape=(n, n))
t = time.time()
op = sb.SparseSymShiftSolve(sp.tril(A).tocsc())
t_up = time.time() - t
t = time.time()
op.set_shift(sigma)
t_fac = time.time() - t
t = time.time()
op.set_shift(sigma)
t_fac2 = time.time() - t
ms_solve = op.solve_device_time(20)
op.set_refine(0)
ms_solve_norefine = op.solve_device_time(20)
op.set_refine(1)
out = dict(n=n, layout=op.layout(), upload_s=round(t_up, 3), set_shift_first_s=round(t_fac, 4), set_shift_s=round(t_fac2, 4), solve_ms=round(ms_solve, 4),
           solve_ms_no_refine=round(ms_solve_norefine, 4))
print(json.dumps(out), flush=True)
for rep in range(2):
    eigs = sb.SymEigsShiftSolver(op, nev, ncv, sigma)
    t = time.time()
    eigs.init()
    nconv = eigs.compute(sb.SortRule.LargestMagn)
    wall = time.time() - t
    st = eigs.stats()
    evals = eigs.eigenvalues()
    X = eigs.eigenvectors()
    res = float((np.linalg.norm(A @ X - X * evals, axis=0) / np.abs(evals)).max())
    print(json.dumps(dict(rep=rep, wall_s=round(wall, 4), ms_total=round(st["ms_total"], 2), nconv=nconv, nops=eigs.num_operations(), niter=eigs.num_iterations(),
                          info=int(eigs.info()), max_rel_res=res, ops_per_s=round(eigs.num_operations() / max(st["ms_total"], 1e-9) * 1e3, 1))), flush=True)
if os.environ.get("SB_NO_CPU") != "1":
    import oracle as O

    csr = O.Csr.adopt(n, rp, ci, v)
    t = time.time()
    lu = O.BandLu(csr, sigma)
    t_cfac = time.time() - t
    x = np.ones(n)
    t = time.time()
    for _ in range(5):
        lu.perform_op(x)
    t_csolve = (time.time() - t) / 5
    r = O.sym_shift_eigs(lu, nev, ncv, O.LargestMagn, want_vectors=False)
    print(json.dumps(dict(cpu="oracle band LU + driver, 1 thread", factor_s=round(t_cfac, 4), solve_ms=round(t_csolve * 1e3, 3), eigs_s=round(r.seconds, 4), nops=r.nops,
                          ops_per_s=round(r.nops / r.seconds, 1), ev_rel_diff=float(np.abs(np.sort(r.eigenvalues) - np.sort(evals)).max() / np.abs(evals).max()))), flush=True)

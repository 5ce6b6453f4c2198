"""Developer probe of the complex Hermitian path: builds a random sparse Hermitian matrix (n = 1e6, ~20 nnz/row), times the operator
and one HermEigsSolver solve (k = 10, ncv = 30), checks the residuals.  usage: python tools/herm_probe.py [n]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import scipy.sparse as sp

import spectra_b200 as sb

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
rng = np.random.default_rng(0)
nnz_half = 10 * n
r = rng.integers(0, n, nnz_half)
c = rng.integers(0, n, nnz_half)
v = (rng.random(nnz_half) - 0.5) + 1j * (rng.random(nnz_half) - 0.5)
lo = r > c
L = sp.csc_matrix((v[lo], (r[lo], c[lo])), shape=(n, n))
L.sum_duplicates()
D = sp.diags(rng.random(n) - 0.5).astype(np.complex128)
A = (L + D).tocsc()                      # lower triangle + real diagonal: what SparseHermMatProd reads
Af = (L + L.conj().T + D).tocsr()
t = time.time()
op = sb.SparseHermMatProd(A)
up = time.time() - t
x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
y = op.perform_op(x)
err = float(np.abs(y - Af @ x).max() / np.abs(y).max())
sb.set_profiling(1)
eigs = sb.HermEigsSolver(op, 10, 30)
t = time.time()
eigs.init()
nconv = eigs.compute(sb.SortRule.LargestAlge)
wall = time.time() - t
ev, U = eigs.eigenvalues(), eigs.eigenvectors()
res = float((np.linalg.norm(Af @ U - U * ev, axis=0) / np.abs(ev)).max())
st = eigs.stats()
print(json.dumps(dict(n=n, nnz=int(Af.nnz), upload_s=round(up, 3), spmv_rel_err=err, nconv=int(nconv), nops=eigs.num_operations(), wall_s=round(wall, 3),
                      ms_total=round(st["ms_total"], 1), ms_spmv=round(st["ms_spmv"], 1), ms_panel=round(st["ms_panel"], 1), ms_compress=round(st["ms_compress"], 1),
                      ms_small=round(st["ms_small"], 1), avg_spmv_ms=round(st["ms_spmv"] / max(st["spmv_launches"], 1), 4),
                      avg_panel_ms=round(st["ms_panel"] / max(st["panel_launches"], 1), 4), max_rel_residual=res)))

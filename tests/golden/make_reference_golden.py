"""Generates tests/golden/reference_<config>.json: the answer of THE REFERENCE ITSELF -- yixuan/spectra's own headers compiled from
/root/reference/include into oracle/_ref/libspectra_ref.so (`make -C oracle ref`; Eigen 3.4 is replaced by oracle/eigen_standin, see its
header) -- on BASELINE.json's full-size configurations.  Needs /root/reference, i.e. runs in the development container only; the JSON files
are the fixtures that travel.

    python tests/golden/make_reference_golden.py C2     n = 1e6 G_sym(seed 0), k = 20, ncv = 60, LargestAlge              (~15 min, 1 thread)
    python tests/golden/make_reference_golden.py C3     n = 1e6 G_gen(seed 1), k = 10, ncv = 30, LargestMagn, maxit = 40  (~2 min)
    python tests/golden/make_reference_golden.py C2magn C2 with the reference's default selection rule

The reference is single-threaded as shipped.  tests/test_oracle_vs_reference.py compares these files with baseline_<config>.json (the
restatement's answers); the GPU tests compare the device results with both."""
import json
import os
import sys
import time

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import oracle as O  # noqa: E402
from oracle import ref as R  # noqa: E402
from spectra_b200 import synth  # noqa: E402
from make_baseline_golden import CONFIGS  # noqa: E402


def main():
    name = sys.argv[1]
    cfg = CONFIGS[name]
    assert R.build(force=True), "needs /root/reference"
    rp, ci, v = synth.csr(cfg["n"], cfg["d"], cfg["seed"], cfg["sym"])
    n = cfg["n"]
    sel = getattr(O, cfg["selection"])
    t = time.time()
    if cfg["sym"]:
        # the full symmetric matrix in CSR = the same matrix in Eigen's default ColMajor storage; SparseSymMatProd<double> reads its lower triangle
        A = R.Compressed(n, rp, ci, v, order="col")
        r = R.sym_eigs(A, cfg["nev"], cfg["ncv"], sel, cfg["maxit"], 1e-10, O.LargestAlge, want_vectors=False)
        ev = dict(eigenvalues=r.eigenvalues.tolist())
        op = "SymEigsSolver<SparseSymMatProd<double, Eigen::Lower, Eigen::ColMajor>>"
    else:
        A = R.Compressed(n, rp, ci, v, order="row")
        r = R.gen_eigs(A, cfg["nev"], cfg["ncv"], sel, cfg["maxit"], 1e-10, O.LargestMagn, want_vectors=False)
        ev = dict(eigenvalues_re=r.eigenvalues.real.tolist(), eigenvalues_im=r.eigenvalues.imag.tolist())
        op = "GenEigsSolver<SparseGenMatProd<double, Eigen::RowMajor>>"
    out = dict(config=name, **cfg, tol=1e-10, nnz=int(len(ci)), threads=1, nconv=r.nconv, niter=r.niter, nops=r.nops, info=r.info,
               seconds=round(time.time() - t, 1), solve_seconds=round(r.seconds, 1), **ev, solver=op, library=R.version(),
               generator="tests/golden/make_reference_golden.py (the reference's own headers over oracle/eigen_standin)")
    path = os.path.join(HERE, f"reference_{name}.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path, {k: out[k] for k in ("nconv", "niter", "nops", "info", "seconds")})


if __name__ == "__main__":
    main()

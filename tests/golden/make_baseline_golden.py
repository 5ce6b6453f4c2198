"""Generates tests/golden/baseline_<config>.json: the CPU oracle's answer on BASELINE.json's full-size configurations, so that the GPU tests
can assert eigenvalue parity (north_star: within 1e-10 relative) and the iteration history at sizes where the oracle takes minutes to hours.

    python tests/golden/make_baseline_golden.py C2 [threads]      n = 1e6 G_sym(seed 0), k = 20, ncv = 60, LargestAlge      (~4 min, 8 threads)
    python tests/golden/make_baseline_golden.py C4 [threads]      n = 1e7 G_sym(seed 0), k = 20, ncv = 60, LargestAlge      (~1-2 h, 8 threads)
    python tests/golden/make_baseline_golden.py C3 [threads]      n = 1e6 G_gen(seed 1), k = 10, ncv = 30, LargestMagn, maxit = 40 (unplanted:
                                                                   the circular-law spectrum does not converge; the bounded history is the fixture)
    python tests/golden/make_baseline_golden.py C2magn [threads]  C2 with the reference's default selection LargestMagn

The oracle (oracle/solver.hpp) is the Eigen-free restatement of the reference; its OpenMP team only parallelises the SpMV and the
panel operations (reductions in a thread-count dependent order), so the stored iteration counts belong to the stated thread count and the
eigenvalues are reproducible to rounding.  Input: spectra_b200.synth (counter-based generator, identical on every host)."""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import oracle as O  # noqa: E402
from spectra_b200 import synth  # noqa: E402

CONFIGS = {
    "C2": dict(n=1_000_000, d=20, seed=0, sym=True, nev=20, ncv=60, selection="LargestAlge", maxit=1000),
    "C2magn": dict(n=1_000_000, d=20, seed=0, sym=True, nev=20, ncv=60, selection="LargestMagn", maxit=1000),
    "C4": dict(n=10_000_000, d=20, seed=0, sym=True, nev=20, ncv=60, selection="LargestAlge", maxit=1000),
    "C3": dict(n=1_000_000, d=20, seed=1, sym=False, nev=10, ncv=30, selection="LargestMagn", maxit=40),
}


def main():
    name = sys.argv[1]
    threads = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    cfg = CONFIGS[name]
    O.build()
    rp, ci, v = synth.csr(cfg["n"], cfg["d"], cfg["seed"], cfg["sym"])
    A = O.Csr.adopt(cfg["n"], rp, ci, v)
    sel = getattr(O, cfg["selection"])
    t = time.time()
    if cfg["sym"]:
        r = O.sym_eigs(A, cfg["nev"], cfg["ncv"], sel, cfg["maxit"], 1e-10, O.LargestAlge, threads=threads, want_vectors=False)
        ev = dict(eigenvalues=r.eigenvalues.tolist())
    else:
        r = O.gen_eigs(A, cfg["nev"], cfg["ncv"], sel, cfg["maxit"], 1e-10, O.LargestMagn, threads=threads, want_vectors=False)
        ev = dict(eigenvalues_re=r.eigenvalues.real.tolist(), eigenvalues_im=r.eigenvalues.imag.tolist())
    out = dict(config=name, **cfg, tol=1e-10, nnz=int(len(ci)), threads=threads, nconv=r.nconv, niter=r.niter, nops=r.nops, info=r.info, restarts=r.restarts,
               reorth_passes=r.reorth_passes, seconds=round(time.time() - t, 1), **ev,
               generator="tests/golden/make_baseline_golden.py (CPU oracle = Eigen-free restatement of the reference; see oracle/)")
    path = os.path.join(HERE, f"baseline_{name}.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path, {k: out[k] for k in ("nconv", "niter", "nops", "info", "seconds")})


if __name__ == "__main__":
    main()

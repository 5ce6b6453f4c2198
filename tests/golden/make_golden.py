"""Generates tests/golden/kat_spectra.json: known-answer spectra of the reference's own test fixtures, computed by an INDEPENDENT dense
solver (LAPACK through numpy) -- not by the oracle and not by the CUDA path.

The reference (C++ headers over Eigen) cannot be compiled in this environment (no Eigen, no network), so there are no golden outputs
of the reference itself; what can be pinned is the mathematical answer on the exact inputs its tests use:

  sym   test/SymEigs.cpp:25-42,133-167   gen_sparse_data(n, p), selfadjointView<Lower>, n in {10, 100, 1000}
  gen   test/GenEigs.cpp:21-36,143-174   gen_sparse_data(n, p) as a general matrix
  herm  test/HermEigs.cpp:27-50,140-174  complex gen_sparse_data(n, p), selfadjointView<Lower>
  diag  SymEigsSolver.h:99-126           diag(1..10)
  cyc   test/Example1.cpp:18-32          cycle-graph Laplacian (analytic spectrum 1 - cos(2 pi j / n))

The input generators restate the reference's (std::default_random_engine seeded 0; oracle/capi.cpp), so the file is reproducible:
    python tests/golden/make_golden.py
Stored per fixture: the full sorted spectrum for n <= 100, the 50 largest / 50 smallest (by the relevant key) for n = 1000.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import oracle as O  # noqa: E402  (input generators only)
from oracle import herm as OH  # noqa: E402

SIZES = {10: 0.5, 100: 0.1, 1000: 0.01}


def trim(vals, keep=50):
    vals = np.asarray(vals)
    return vals if len(vals) <= 2 * keep else np.concatenate([vals[:keep], vals[-keep:]])


def main():
    out = {"_doc": "dense-LAPACK spectra of the reference's test fixtures; see make_golden.py", "sym": {}, "herm": {}, "gen": {}}
    O.build()
    for n, p in SIZES.items():
        A = O.gen_sparse_data(n, p)
        Af = O.Csr.from_scipy(A, "lower").to_scipy().toarray()
        w = np.linalg.eigvalsh(Af)  # ascending
        out["sym"][str(n)] = {"prob": p, "ascending": trim(w).tolist(), "by_magnitude_desc": w[np.argsort(-np.abs(w))][:50].tolist(),
                              "by_magnitude_asc": w[np.argsort(np.abs(w))][:50].tolist()}
        H = OH.herm_full(OH.gen_sparse_data_herm(n, p)).toarray()
        wh = np.linalg.eigvalsh(H)
        out["herm"][str(n)] = {"prob": p, "ascending": trim(wh).tolist(), "by_magnitude_desc": wh[np.argsort(-np.abs(wh))][:50].tolist(),
                               "by_magnitude_asc": wh[np.argsort(np.abs(wh))][:50].tolist()}
        wg = np.linalg.eigvals(A.toarray())
        order = np.argsort(-np.abs(wg), kind="stable")
        top = wg[order][:50]
        out["gen"][str(n)] = {"prob": p, "by_magnitude_desc_re": top.real.tolist(), "by_magnitude_desc_im": top.imag.tolist(),
                              "largest_real": np.sort(wg.real)[::-1][:50].tolist()}
    out["diag10"] = {"largest": [10.0, 9.0, 8.0]}
    n = 10
    out["cycle10"] = {"ascending": sorted((1.0 - np.cos(2.0 * np.pi * np.arange(n) / n)).tolist())}
    with open(os.path.join(HERE, "kat_spectra.json"), "w") as f:
        json.dump(out, f)
    print("wrote", os.path.join(HERE, "kat_spectra.json"))


if __name__ == "__main__":
    main()

"""Checks against tests/golden/kat_spectra.json (dense-LAPACK spectra of the reference's test fixtures, make_golden.py): used for the
CPU oracle (tests/test_oracle.py), the CUDA path (tests/test_gpu_sym.py / test_gpu_gen.py) and the emulator (tests/test_emu_kernels.py).
Tolerance: 1e-10 relative to the spectral radius (north star) for symmetric / Hermitian, 1e-9 for the nonsymmetric fixture."""
import json
import os

import numpy as np

import oracle as O

_G = None


def golden():
    global _G
    if _G is None:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kat_spectra.json")) as f:
            _G = json.load(f)
    return _G


KM = {10: (3, 6), 100: (10, 20), 1000: (20, 50)}


def expected_sym(kind, n, rule, k):
    g = golden()[kind][str(n)]
    asc = np.array(g["ascending"])
    if rule == O.LargestAlge:
        return asc[::-1][:k]
    if rule == O.SmallestAlge:
        return asc[:k]
    if rule == O.LargestMagn:
        return np.array(g["by_magnitude_desc"])[:k]
    if rule == O.SmallestMagn:
        return np.array(g["by_magnitude_asc"])[:k]
    raise ValueError(rule)


def check_sym_values(kind, n, rule, evals):
    k = KM[n][0]
    exp = expected_sym(kind, n, rule, k)
    scale = np.abs(np.array(golden()[kind][str(n)]["by_magnitude_desc"])).max()
    assert len(evals) == k
    assert np.abs(np.sort(evals) - np.sort(exp)).max() <= 1e-10 * scale, (kind, n, rule)


def check_gen_values(n, evals, k):
    g = golden()["gen"][str(n)]
    exp = np.array(g["by_magnitude_desc_re"]) + 1j * np.array(g["by_magnitude_desc_im"])
    scale = np.abs(exp).max()
    # every returned value is one of the golden largest-magnitude values (conjugate pairs may straddle the cut at k)
    for e in evals:
        assert np.abs(exp - e).min() <= 1e-9 * scale, (n, e)
    assert len(evals) == k

"""Synthetic sparse workloads (host side, C + OpenMP, ctypes).  See synth.c for the definition of
G_sym / G_gen; used by bench.py and the tests to build identical inputs for the GPU path and the CPU
baseline."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libspectra_synth.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "synth.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.run(["/usr/bin/gcc", "-O2", "-fopenmp", "-fPIC", "-shared", src, "-o", _LIB], check=True, capture_output=True)
    return _LIB


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB)
        _lib.synth_csr_count.restype = C.c_int64
        _lib.synth_band_count.restype = C.c_int64
    return _lib


def csr(n: int, d: int = 20, seed: int = 0, sym: bool = True, row0: int = 0, nrows: int | None = None, alloc=None):
    """Rows [row0, row0+nrows) of G_sym / G_gen as (rowptr int64 (local offsets), col int32, val float64).
    `alloc(count, dtype)` may supply the output arrays (e.g. pinned host memory)."""
    nrows = n - row0 if nrows is None else nrows
    alloc = alloc or (lambda count, dtype: np.empty(count, dtype))
    rowptr = alloc(nrows + 1, np.int64)
    nnz = lib().synth_csr_count(C.c_int64(n), int(d), int(bool(sym)), C.c_uint64(seed), C.c_int64(row0), C.c_int64(nrows), rowptr.ctypes.data_as(C.c_void_p))
    if nnz < 0:
        raise ValueError("bad generator arguments")
    col = alloc(nnz, np.int32)
    val = alloc(nnz, np.float64)
    rc = lib().synth_csr_fill(C.c_int64(n), int(d), int(bool(sym)), C.c_uint64(seed), C.c_int64(row0), C.c_int64(nrows), rowptr.ctypes.data_as(C.c_void_p),
                              col.ctypes.data_as(C.c_void_p), val.ctypes.data_as(C.c_void_p))
    assert rc == 0
    return rowptr, col, val


def band_csr(n: int, b: int = 15, seed: int = 0, diag_add: float = 0.0, row0: int = 0, nrows: int | None = None):
    """Rows [row0, row0+nrows) of G_band: symmetric, half-bandwidth b, entries U(-0.5, 0.5), `diag_add` added on the diagonal."""
    nrows = n - row0 if nrows is None else nrows
    rowptr = np.empty(nrows + 1, np.int64)
    nnz = lib().synth_band_count(C.c_int64(n), int(b), C.c_int64(row0), C.c_int64(nrows), rowptr.ctypes.data_as(C.c_void_p))
    if nnz < 0:
        raise ValueError("bad generator arguments")
    col = np.empty(nnz, np.int32)
    val = np.empty(nnz, np.float64)
    rc = lib().synth_band_fill(C.c_int64(n), int(b), C.c_uint64(seed), C.c_double(diag_add), C.c_int64(row0), C.c_int64(nrows),
                               rowptr.ctypes.data_as(C.c_void_p), col.ctypes.data_as(C.c_void_p), val.ctypes.data_as(C.c_void_p))
    assert rc == 0
    return rowptr, col, val


def scipy_csr(n: int, d: int = 20, seed: int = 0, sym: bool = True):
    import scipy.sparse as sp

    rp, ci, v = csr(n, d, seed, sym)
    return sp.csr_matrix((v, ci, rp), shape=(n, n))

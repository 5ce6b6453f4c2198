"""The C-ABI library loads and exports every symbol include/spectra_b200.h declares; without a GPU
every compute entry fails loudly (there is no CPU fallback).  CPU only."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "spectra_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sb200_[a-z0-9_]+)\s*\(", src)))


def _lib_path():
    import spectra_b200 as sb

    if not os.path.exists(sb.lib_path()):
        from spectra_b200 import _build

        _build.build()
    return sb.lib_path()


def test_every_declared_symbol_is_exported():
    names = _declared()
    assert len(names) >= 45
    out = subprocess.run(["nm", "-D", "--defined-only", _lib_path()], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (sb200_[a-z0-9_]+)", out))
    missing = [n for n in names if n not in exported]
    assert not missing, missing
    L = C.CDLL(_lib_path())
    for n in names:
        assert getattr(L, n) is not None


def test_library_contains_only_sm100a_code():
    out = subprocess.run(["cuobjdump", "-lelf", _lib_path()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_silent_cpu_fallback():
    import spectra_b200 as sb

    try:
        sb.device_info()
    except sb.CudaError:
        pass
    else:
        pytest.skip("a GPU is present; the loud-failure path is exercised on CPU-only hosts")
    import scipy.sparse as sp

    A = sp.identity(8, format="csc")
    with pytest.raises(sb.CudaError):
        sb.SparseSymMatProd(A)
    with pytest.raises(sb.CudaError):
        sb.dense.tridiag_eigen(np.eye(4))


def test_product_does_not_reference_oracle():
    # the shipped package must not import / link the test oracle
    pkg = os.path.join(ROOT, "spectra_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".c")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert "import oracle" not in txt and "liboracle" not in txt and "oracle/" not in txt, os.path.join(dp, f)
    out = subprocess.run(["ldd", _lib_path()], capture_output=True, text=True).stdout
    assert "oracle" not in out


def test_result_buffer_never_overruns_a_caller_array():
    # perform_op(x, y_out): a caller-supplied y_out is handed to the C ABI only when it is a writeable C-contiguous array of exactly the
    # element type and length the library writes; anything else gets the result through a temporary (round-1 advisor finding)
    import spectra_b200 as sb

    y = np.zeros(5)
    buf, user = sb._result_buffer(y, 5, np.float64)
    assert buf is y and user is None
    for bad in (np.zeros(5, dtype=np.float32), np.zeros(10)[::2]):
        buf, user = sb._result_buffer(bad, 5, np.float64)
        assert user is bad and buf is not bad and buf.dtype == np.float64 and buf.flags.c_contiguous and buf.size == 5
    with pytest.raises(sb.InvalidArgument):
        sb._result_buffer(np.zeros(4), 5, np.float64)
    with pytest.raises(sb.InvalidArgument):
        sb._result_buffer([0.0] * 5, 5, np.float64)
    ro = np.zeros(5)
    ro.setflags(write=False)
    with pytest.raises(sb.InvalidArgument):
        sb._result_buffer(ro, 5, np.float64)
    buf, user = sb._result_buffer(None, 3, np.complex128)
    assert user is None and buf.dtype == np.complex128 and buf.size == 3

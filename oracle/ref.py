"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes binding of ``oracle/_ref/libspectra_ref.so``: the REFERENCE'S OWN HEADERS
(/root/reference/include/Spectra, compiled where they lie by ``make -C oracle ref``) behind a C
shim (``ref_capi.cpp``), with ``oracle/eigen_standin`` in place of Eigen 3.4, which this image
does not have.  Used to pin the restatement in ``oracle/*.hpp`` on outputs of the reference
itself and, in ``bench.py --impl reference``, as the reference's CPU arm.

``/root/reference`` exists only in the development container; the built ``.so`` travels to the
GPU box (git-ignored, not gpurun-ignored).  Nothing here reads ``/root/reference`` at run time.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import EigsResult, OracleError

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libspectra_ref.so")
REFERENCE_INCLUDE = "/root/reference/include"
_lib = None


class _RefResult(C.Structure):
    _fields_ = [("nconv", C.c_int64), ("niter", C.c_int64), ("nops", C.c_int64), ("info", C.c_int32), ("seconds", C.c_double)]


class _Compressed(C.Structure):
    _fields_ = [("n", C.c_int64), ("nnz", C.c_int64), ("outer", C.c_void_p), ("inner", C.c_void_p), ("val", C.c_void_p)]


def available() -> bool:
    """True when the reference library has been built (or can be built here)."""
    return os.path.exists(_LIB_PATH) or os.path.isdir(os.path.join(REFERENCE_INCLUDE, "Spectra"))


def build(force: bool = False) -> str | None:
    """Compile the reference's headers with the committed recipe; a no-op without /root/reference."""
    srcs = [os.path.join(_HERE, f) for f in ("ref_capi.cpp", "Makefile", os.path.join("eigen_standin", "Eigen", "src", "standin.h"))]
    have_ref = os.path.isdir(os.path.join(REFERENCE_INCLUDE, "Spectra"))
    stale = (not os.path.exists(_LIB_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs)
    if have_ref and (force or stale):
        subprocess.run(["make", "-C", _HERE, "--no-print-directory", "ref"], check=True, capture_output=True)
    return _LIB_PATH if os.path.exists(_LIB_PATH) else None


def lib():
    global _lib
    if _lib is None:
        if build() is None:
            raise OracleError(-1, "oracle/_ref/libspectra_ref.so is not built and /root/reference is not present")
        _lib = C.CDLL(_LIB_PATH)
        _lib.ref_last_error.restype = C.c_char_p
        _lib.ref_version.restype = C.c_char_p
        _lib.ref_givens.argtypes = [C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
    return _lib


def version() -> str:
    return lib().ref_version().decode()


def _check(rc):
    if rc != 0:
        raise OracleError(rc, lib().ref_last_error().decode())


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _colmajor(M):
    return np.asfortranarray(np.asarray(M, dtype=np.float64))


class Compressed:
    """A square compressed matrix in Eigen's layout with the default 32-bit StorageIndex.

    order 'col': outer = column pointers (Eigen::ColMajor); 'row': row pointers (Eigen::RowMajor).
    """

    def __init__(self, n, outer, inner, val, order="col"):
        self.n = int(n)
        self.order = 0 if order == "col" else 1
        self.outer = np.ascontiguousarray(outer, dtype=np.int32)
        self.inner = np.ascontiguousarray(inner, dtype=np.int32)
        self.val = _f64(val)
        assert self.outer.size == self.n + 1 and int(self.outer[-1]) == self.inner.size == self.val.size
        self.c = _Compressed(self.n, self.val.size, self.outer.ctypes.data, self.inner.ctypes.data, self.val.ctypes.data)

    @classmethod
    def from_scipy(cls, A):
        """CSC or CSR scipy matrix, taken as stored (no sorting, no duplicate merging)."""
        import scipy.sparse as sp

        if sp.isspmatrix_csr(A):
            return cls(A.shape[0], A.indptr, A.indices, A.data, order="row")
        A = sp.csc_matrix(A)
        return cls(A.shape[0], A.indptr, A.indices, A.data, order="col")


def simple_random(seed: int, n: int) -> np.ndarray:
    out = np.empty(n)
    lib().ref_simple_random(C.c_uint64(seed), C.c_int64(n), _p(out))
    return out


def givens(x, y):
    r, c, s = C.c_double(), C.c_double(), C.c_double()
    lib().ref_givens(x, y, C.byref(r), C.byref(c), C.byref(s))
    return r.value, c.value, s.value


def shifted_qr(H, shift, kind="tridiag"):
    H = _colmajor(H)
    m = H.shape[0]
    R, QtHQ, Q = (np.empty((m, m), order="F") for _ in range(3))
    _check(lib().ref_shifted_qr(0 if kind == "tridiag" else 1, C.c_int64(m), _p(H), C.c_double(shift), _p(R), _p(QtHQ), _p(Q)))
    return R, QtHQ, Q


def double_shift_qr(H, s, t):
    H = _colmajor(H)
    m = H.shape[0]
    QtHQ, Q = (np.empty((m, m), order="F") for _ in range(2))
    _check(lib().ref_double_shift_qr(C.c_int64(m), _p(H), C.c_double(s), C.c_double(t), _p(QtHQ), _p(Q)))
    return QtHQ, Q


def tridiag_eigen(H):
    H = _colmajor(H)
    m = H.shape[0]
    evals = np.empty(m)
    evecs = np.empty((m, m), order="F")
    _check(lib().ref_tridiag_eigen(C.c_int64(m), _p(H), _p(evals), _p(evecs)))
    return evals, evecs


def hess_schur(H):
    H = _colmajor(H)
    m = H.shape[0]
    T, U = (np.empty((m, m), order="F") for _ in range(2))
    _check(lib().ref_hess_schur(C.c_int64(m), _p(H), _p(T), _p(U)))
    return T, U


def hess_eigen(H):
    H = _colmajor(H)
    m = H.shape[0]
    ev = np.empty(2 * m)
    V = np.empty(2 * m * m)
    _check(lib().ref_hess_eigen(C.c_int64(m), _p(H), _p(ev), _p(V)))
    return ev.view(np.complex128), V.view(np.complex128).reshape((m, m), order="F")


def argsort(selection, values):
    values = np.asarray(values)
    n = values.size
    ind = np.empty(n, dtype=np.int64)
    if np.iscomplexobj(values):
        v = np.ascontiguousarray(values, dtype=np.complex128)
        _check(lib().ref_argsort_complex(int(selection), _p(v), C.c_int64(n), _p(ind)))
    else:
        v = _f64(values)
        _check(lib().ref_argsort(int(selection), _p(v), C.c_int64(n), _p(ind)))
    return ind


def spmv(A: Compressed, x, sym=False, uplo="lower"):
    x = _f64(x)
    y = np.empty(A.n)
    _check(lib().ref_spmv(int(bool(sym)), 0 if uplo == "lower" else 1, A.order, C.byref(A.c), _p(x), _p(y)))
    return y


def coeff(A: Compressed, i, j, sym=False):
    out = C.c_double()
    _check(lib().ref_coeff(int(bool(sym)), A.order, C.byref(A.c), C.c_int64(i), C.c_int64(j), C.byref(out)))
    return out.value


def factorize(A: Compressed, m: int, v0=None, mid=None, kind="lanczos"):
    """Lanczos (SparseSymMatProd<Lower, ColMajor>: A must be 'col' ordered) or Arnoldi factorisation."""
    n = A.n
    assert kind != "lanczos" or A.order == 0, "the Lanczos hook binds SparseSymMatProd<double, Lower, ColMajor>"
    mid = m if mid is None else mid
    V = np.empty((n, m), order="F")
    H = np.empty((m, m), order="F")
    f = np.empty(n)
    beta = C.c_double()
    nops = C.c_int64()
    v0a = _f64(v0) if v0 is not None else None
    _check(lib().ref_factorize(0 if kind == "lanczos" else 1, A.order, C.byref(A.c), C.c_int64(m), _p(v0a), C.c_int64(mid), _p(V), _p(H), _p(f),
                               C.byref(beta), C.byref(nops), None))
    return V, H, f, beta.value, nops.value


def lanczos_sample(A: Compressed, m: int):
    """init() + the first m - 1 steps of Lanczos::factorize_from on SparseSymMatProd<double> from the default SimpleRandom(0) residual -- the
    head of a SymEigsSolver solve -- timed inside the library.  Returns (operator applications, seconds)."""
    beta = C.c_double()
    nops = C.c_int64()
    sec = C.c_double()
    _check(lib().ref_factorize(0, A.order, C.byref(A.c), C.c_int64(m), None, C.c_int64(m), None, None, None, C.byref(beta), C.byref(nops), C.byref(sec)))
    return nops.value, sec.value


def sym_eigs(A: Compressed, nev, ncv, selection=0, maxit=1000, tol=1e-10, sorting=3, init_resid=None, uplo="lower", want_vectors=True) -> EigsResult:
    """SymEigsSolver<SparseSymMatProd<double, Uplo, Flags>>: init() / init(resid), compute(), eigenvalues(), eigenvectors()."""
    n = A.n
    res = _RefResult()
    evals = np.zeros(nev)
    evecs = np.zeros(n * nev) if want_vectors else None
    r0 = _f64(init_resid) if init_resid is not None else None
    _check(lib().ref_sym_eigs(0 if uplo == "lower" else 1, A.order, C.byref(A.c), C.c_int64(nev), C.c_int64(ncv), int(selection), C.c_int64(maxit),
                              C.c_double(tol), int(sorting), _p(r0), _p(evals), _p(evecs), C.byref(res)))
    nconv = int(res.nconv)
    vec = evecs.reshape((n, -1), order="F")[:, :nconv].copy() if want_vectors and nconv else None
    return EigsResult(nconv=nconv, niter=int(res.niter), nops=int(res.nops), info=int(res.info), eigenvalues=evals[:nconv].copy(), eigenvectors=vec,
                      reorth_passes=-1, expand_calls=-1, restarts=-1, steps=-1, seconds=float(res.seconds))


def sym_eigs_userop(n, fn, nev, ncv, selection=0, maxit=1000, tol=1e-10, sorting=3, init_resid=None) -> EigsResult:
    CB = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_void_p)

    def tramp(xp, yp, _):
        x = np.ctypeslib.as_array(xp, shape=(n,))
        y = np.ctypeslib.as_array(yp, shape=(n,))
        y[:] = fn(x)

    cb = CB(tramp)
    res = _RefResult()
    evals = np.zeros(nev)
    evecs = np.zeros(n * nev)
    r0 = _f64(init_resid) if init_resid is not None else None
    _check(lib().ref_sym_eigs_userop(C.c_int64(n), cb, None, C.c_int64(nev), C.c_int64(ncv), int(selection), C.c_int64(maxit), C.c_double(tol),
                                     int(sorting), _p(r0), _p(evals), _p(evecs), C.byref(res)))
    nconv = int(res.nconv)
    vec = evecs.reshape((n, -1), order="F")[:, :nconv].copy() if nconv else None
    return EigsResult(nconv=nconv, niter=int(res.niter), nops=int(res.nops), info=int(res.info), eigenvalues=evals[:nconv].copy(), eigenvectors=vec,
                      reorth_passes=-1, expand_calls=-1, restarts=-1, steps=-1, seconds=float(res.seconds))


def gen_eigs(A: Compressed, nev, ncv, selection=0, maxit=1000, tol=1e-10, sorting=0, init_resid=None, want_vectors=True) -> EigsResult:
    """GenEigsSolver<SparseGenMatProd<double, Flags>>."""
    n = A.n
    res = _RefResult()
    evals = np.zeros(2 * nev)
    evecs = np.zeros(2 * n * nev) if want_vectors else None
    r0 = _f64(init_resid) if init_resid is not None else None
    _check(lib().ref_gen_eigs(A.order, C.byref(A.c), C.c_int64(nev), C.c_int64(ncv), int(selection), C.c_int64(maxit), C.c_double(tol), int(sorting),
                              _p(r0), _p(evals), _p(evecs), C.byref(res)))
    nconv = int(res.nconv)
    ev = evals.view(np.complex128)[:nconv].copy()
    vec = evecs.view(np.complex128)[: n * nconv].reshape((n, nconv), order="F").copy() if want_vectors and nconv else None
    return EigsResult(nconv=nconv, niter=int(res.niter), nops=int(res.nops), info=int(res.info), eigenvalues=ev, eigenvectors=vec,
                      reorth_passes=-1, expand_calls=-1, restarts=-1, steps=-1, seconds=float(res.seconds))


# ---------------------------------------------------------------- complex Hermitian path (HermEigsSolver + SparseHermMatProd)
class CompressedZ:
    """Column-major compressed complex matrix (Eigen::SparseMatrix<std::complex<double>>), values interleaved (re, im)."""

    def __init__(self, A):
        import scipy.sparse as sp

        A = sp.csc_matrix(A)
        self.n = int(A.shape[0])
        self.outer = np.ascontiguousarray(A.indptr, dtype=np.int32)
        self.inner = np.ascontiguousarray(A.indices, dtype=np.int32)
        self.val = np.ascontiguousarray(A.data, dtype=np.complex128)
        self.c = _Compressed(self.n, self.val.size, self.outer.ctypes.data, self.inner.ctypes.data, self.val.ctypes.data)


def simple_random_complex(seed: int, n: int) -> np.ndarray:
    out = np.empty(n, dtype=np.complex128)
    lib().ref_simple_random_complex(C.c_uint64(seed), C.c_int64(n), _p(out))
    return out


def herm_spmv(A: CompressedZ, x, uplo="lower"):
    x = np.ascontiguousarray(x, dtype=np.complex128)
    y = np.empty(A.n, dtype=np.complex128)
    _check(lib().ref_herm_spmv(0 if uplo == "lower" else 1, C.byref(A.c), _p(x), _p(y)))
    return y


def _herm_result(res, evals, evecs, n):
    nconv = int(res.nconv)
    vec = evecs[: n * nconv].reshape((n, nconv), order="F").copy() if nconv else None
    return EigsResult(nconv=nconv, niter=int(res.niter), nops=int(res.nops), info=int(res.info), eigenvalues=evals[:nconv].copy(), eigenvectors=vec,
                      reorth_passes=-1, expand_calls=-1, restarts=-1, steps=-1, seconds=float(res.seconds))


def herm_eigs(A: CompressedZ, nev, ncv, selection=0, maxit=1000, tol=1e-10, sorting=3, init_resid=None, uplo="lower") -> EigsResult:
    """HermEigsSolver<SparseHermMatProd<std::complex<double>, Uplo>>."""
    res = _RefResult()
    evals = np.zeros(nev)
    evecs = np.zeros(A.n * nev, dtype=np.complex128)
    r0 = np.ascontiguousarray(init_resid, dtype=np.complex128) if init_resid is not None else None
    _check(lib().ref_herm_eigs(0 if uplo == "lower" else 1, C.byref(A.c), C.c_int64(nev), C.c_int64(ncv), int(selection), C.c_int64(maxit), C.c_double(tol),
                               int(sorting), _p(r0), _p(evals), _p(evecs), C.byref(res)))
    return _herm_result(res, evals, evecs, A.n)


def herm_eigs_userop(n, fn, nev, ncv, selection=0, maxit=1000, tol=1e-10, sorting=3, init_resid=None) -> EigsResult:
    """HermEigsSolver over a user-defined complex OpType: fn maps a complex vector to a complex vector."""
    CB = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_void_p)

    def tramp(xp, yp, _):
        x = np.ctypeslib.as_array(xp, shape=(2 * n,)).view(np.complex128)
        y = np.ctypeslib.as_array(yp, shape=(2 * n,)).view(np.complex128)
        y[:] = fn(x)

    cb = CB(tramp)
    res = _RefResult()
    evals = np.zeros(nev)
    evecs = np.zeros(n * nev, dtype=np.complex128)
    r0 = np.ascontiguousarray(init_resid, dtype=np.complex128) if init_resid is not None else None
    _check(lib().ref_herm_eigs_userop(C.c_int64(n), cb, None, C.c_int64(nev), C.c_int64(ncv), int(selection), C.c_int64(maxit), C.c_double(tol), int(sorting),
                                      _p(r0), _p(evals), _p(evecs), C.byref(res)))
    return _herm_result(res, evals, evecs, n)


# ---------------------------------------------------------------- complex general path (GenEigsSolver with a complex Scalar)
class CompressedZG(CompressedZ):
    """Complex compressed matrix, column-major (default) or row-major as given."""

    def __init__(self, A):
        import scipy.sparse as sp

        if sp.isspmatrix_csr(A):
            self.order = 1
            self.n = int(A.shape[0])
            self.outer = np.ascontiguousarray(A.indptr, dtype=np.int32)
            self.inner = np.ascontiguousarray(A.indices, dtype=np.int32)
            self.val = np.ascontiguousarray(A.data, dtype=np.complex128)
            self.c = _Compressed(self.n, self.val.size, self.outer.ctypes.data, self.inner.ctypes.data, self.val.ctypes.data)
        else:
            super().__init__(A)
            self.order = 0


def _genz_result(res, evals, evecs, n):
    nconv = int(res.nconv)
    ev = evals.view(np.complex128)[:nconv].copy()
    vec = evecs[: n * nconv].reshape((n, nconv), order="F").copy() if nconv else None
    return EigsResult(nconv=nconv, niter=int(res.niter), nops=int(res.nops), info=int(res.info), eigenvalues=ev, eigenvectors=vec,
                      reorth_passes=-1, expand_calls=-1, restarts=-1, steps=-1, seconds=float(res.seconds))


def gen_eigs_complex(A: CompressedZG, nev, ncv, selection=0, maxit=1000, tol=1e-10, sorting=0, init_resid=None) -> EigsResult:
    res = _RefResult()
    evals = np.zeros(2 * nev)
    evecs = np.zeros(A.n * nev, dtype=np.complex128)
    r0 = np.ascontiguousarray(init_resid, dtype=np.complex128) if init_resid is not None else None
    _check(lib().ref_gen_eigs_complex(A.order, C.byref(A.c), C.c_int64(nev), C.c_int64(ncv), int(selection), C.c_int64(maxit), C.c_double(tol), int(sorting),
                                      _p(r0), _p(evals), _p(evecs), C.byref(res)))
    return _genz_result(res, evals, evecs, A.n)


def gen_eigs_complex_userop(n, fn, nev, ncv, selection=0, maxit=1000, tol=1e-10, sorting=0, init_resid=None) -> EigsResult:
    CB = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_void_p)

    def tramp(xp, yp, _):
        x = np.ctypeslib.as_array(xp, shape=(2 * n,)).view(np.complex128)
        y = np.ctypeslib.as_array(yp, shape=(2 * n,)).view(np.complex128)
        y[:] = fn(x)

    cb = CB(tramp)
    res = _RefResult()
    evals = np.zeros(2 * nev)
    evecs = np.zeros(n * nev, dtype=np.complex128)
    r0 = np.ascontiguousarray(init_resid, dtype=np.complex128) if init_resid is not None else None
    _check(lib().ref_gen_eigs_complex_userop(C.c_int64(n), cb, None, C.c_int64(nev), C.c_int64(ncv), int(selection), C.c_int64(maxit), C.c_double(tol),
                                             int(sorting), _p(r0), _p(evals), _p(evecs), C.byref(res)))
    return _genz_result(res, evals, evecs, n)


def hess_eigen_complex(H):
    H = np.asfortranarray(np.asarray(H, dtype=np.complex128))
    m = H.shape[0]
    ev = np.empty(m, dtype=np.complex128)
    V = np.empty((m, m), dtype=np.complex128, order="F")
    _check(lib().ref_hess_eigen_complex(C.c_int64(m), _p(H), _p(ev), _p(V)))
    return ev, V


def shifted_qr_complex(H, shift):
    H = np.asfortranarray(np.asarray(H, dtype=np.complex128))
    m = H.shape[0]
    R_, D, Q = (np.empty((m, m), dtype=np.complex128, order="F") for _ in range(3))
    z = complex(shift)
    _check(lib().ref_shifted_qr_complex(C.c_int64(m), _p(H), C.c_double(z.real), C.c_double(z.imag), _p(R_), _p(D), _p(Q)))
    return R_, D, Q


def givens_complex(x, y):
    x, y = complex(x), complex(y)
    r = (C.c_double * 2)()
    s = (C.c_double * 2)()
    c = C.c_double()
    lib().ref_givens_complex(C.c_double(x.real), C.c_double(x.imag), C.c_double(y.real), C.c_double(y.imag), r, C.byref(c), s)
    return complex(r[0], r[1]), c.value, complex(s[0], s[1])


# ---------------------------------------------------------------- shift-and-invert (SymEigsShiftSolver + SparseSymShiftSolve)
def shift_solve(A: Compressed, sigma, x, uplo="lower"):
    """SparseSymShiftSolve<double, Uplo>(A).set_shift(sigma); perform_op(x) = (A - sigma I)^{-1} x  (A column-major)."""
    assert A.order == 0
    x = _f64(x)
    y = np.empty(A.n)
    _check(lib().ref_shift_solve(0 if uplo == "lower" else 1, C.byref(A.c), C.c_double(sigma), _p(x), _p(y)))
    return y


def sym_shift_eigs(A: Compressed, sigma, nev, ncv, selection=0, maxit=1000, tol=1e-10, sorting=3, init_resid=None, uplo="lower", want_vectors=True) -> EigsResult:
    """SymEigsShiftSolver<SparseSymShiftSolve<double, Uplo>>(op, nev, ncv, sigma)."""
    assert A.order == 0
    n = A.n
    res = _RefResult()
    evals = np.zeros(nev)
    evecs = np.zeros(n * nev) if want_vectors else None
    r0 = _f64(init_resid) if init_resid is not None else None
    _check(lib().ref_sym_shift_eigs(0 if uplo == "lower" else 1, C.byref(A.c), C.c_double(sigma), C.c_int64(nev), C.c_int64(ncv), int(selection), C.c_int64(maxit),
                                    C.c_double(tol), int(sorting), _p(r0), _p(evals), _p(evecs), C.byref(res)))
    nconv = int(res.nconv)
    vec = evecs.reshape((n, -1), order="F")[:, :nconv].copy() if want_vectors and nconv else None
    return EigsResult(nconv=nconv, niter=int(res.niter), nops=int(res.nops), info=int(res.info), eigenvalues=evals[:nconv].copy(), eigenvectors=vec,
                      reorth_passes=-1, expand_calls=-1, restarts=-1, steps=-1, seconds=float(res.seconds))

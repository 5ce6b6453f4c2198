"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes binding of ``oracle/_build/liboracle.so`` (the Eigen-free CPU restatement of the
reference's implicitly-restarted Lanczos/Arnoldi path, see ``dense.hpp`` / ``solver.hpp``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import
this package.  The shipped product (``spectra_b200``) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_STRICT_PATH = os.path.join(_HERE, "_build", "liboracle_strict.so")
_lib = None
_libs = {}
_build_name = "default"

# SortRule (Util/SelectionRule.h:33-58)
LargestMagn, LargestReal, LargestImag, LargestAlge, SmallestMagn, SmallestReal, SmallestImag, SmallestAlge, BothEnds = range(9)
# CompInfo (Util/CompInfo.h:17-30)
Successful, NotComputed, NotConverging, NumericalIssue = range(4)


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (g++ -O2, OpenMP)."""
    srcs = [os.path.join(_HERE, f) for f in ("capi.cpp", "dense.hpp", "solver.hpp", "band.hpp", "Makefile")]
    stale = any((not os.path.exists(p)) or any(os.path.getmtime(s) > os.path.getmtime(p) for s in srcs) for p in (_LIB_PATH, _STRICT_PATH))
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "--no-print-directory"], check=True, capture_output=True)
    return _LIB_PATH


class _Result(C.Structure):
    _fields_ = [
        ("nconv", C.c_int64),
        ("niter", C.c_int64),
        ("nops", C.c_int64),
        ("info", C.c_int32),
        ("reorth_passes", C.c_int64),
        ("expand_calls", C.c_int64),
        ("restarts", C.c_int64),
        ("steps", C.c_int64),
        ("seconds", C.c_double),
    ]


def select_build(name: str) -> str:
    """Choose which build of the restatement the functions below call: "default" (FMA contraction, AVX2: the fast one) or
    "strict" (-ffp-contract=off: every operation rounded once as written, bit-comparable with oracle/_ref).  Returns the previous choice."""
    global _lib, _build_name
    if name not in ("default", "strict"):
        raise ValueError(name)
    prev, _build_name = _build_name, name
    _lib = _libs.get(name)
    return prev


def lib():
    global _lib
    if _lib is None:
        path = _LIB_PATH if _build_name == "default" else _STRICT_PATH
        if not os.path.exists(path):
            build()
        _lib = _libs[_build_name] = C.CDLL(path)
        _lib.oracle_last_error.restype = C.c_char_p
        _lib.oracle_csr_nnz.restype = C.c_int64
        _lib.oracle_gen_sparse_data.restype = C.c_int64
        _lib.oracle_csr_nnz.argtypes = [C.c_void_p]
        _lib.oracle_csr_free.argtypes = [C.c_void_p]
        _lib.oracle_csr_export.argtypes = [C.c_void_p] * 4
        _lib.oracle_csr_set_threads.argtypes = [C.c_void_p, C.c_int]
        _lib.oracle_spmv.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    return _lib


class OracleError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


_EXC = {1: ValueError, 2: RuntimeError, 3: RuntimeError, 4: RuntimeError}


def _check(rc):
    if rc != 0:
        msg = lib().oracle_last_error().decode()
        # 1 = std::invalid_argument, 2 = std::logic_error, 3 = std::runtime_error
        raise OracleError(rc, msg)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def max_threads() -> int:
    return int(lib().oracle_max_threads())


class Csr:
    """Full CSR operator built the way Sparse{Sym,Gen}MatProd interpret an Eigen sparse matrix.

    order: 'col' (Eigen::ColMajor, default) or 'row'; mode: 'gen', 'lower', 'upper'.
    """

    def __init__(self, n, outer, inner, val, order="col", mode="gen"):
        self.n = int(n)
        outer = np.ascontiguousarray(outer, dtype=np.int64)
        inner = np.ascontiguousarray(inner, dtype=np.int32)
        val = _f64(val)
        h = C.c_void_p()
        _check(lib().oracle_csr_create(C.c_int64(self.n), _p(outer), _p(inner), _p(val), {"col": 0, "row": 1}[order],
                                       {"gen": 0, "lower": 1, "upper": 2}[mode], C.byref(h)))
        self.h = h

    @classmethod
    def adopt(cls, n, rowptr, col, val):
        """Full CSR used as is (no triangle expansion / transpose)."""
        self = cls.__new__(cls)
        self.n = int(n)
        rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
        col = np.ascontiguousarray(col, dtype=np.int32)
        val = _f64(val)
        h = C.c_void_p()
        _check(lib().oracle_csr_adopt(C.c_int64(self.n), _p(rowptr), _p(col), _p(val), C.byref(h)))
        self.h = h
        return self

    @classmethod
    def from_scipy(cls, A, mode="gen"):
        import scipy.sparse as sp

        if sp.isspmatrix_csr(A):
            return cls(A.shape[0], A.indptr, A.indices, A.data, "row", mode)
        A = sp.csc_matrix(A)
        return cls(A.shape[0], A.indptr, A.indices, A.data, "col", mode)

    @classmethod
    def from_dense(cls, M, mode="gen"):
        M = np.asarray(M, dtype=np.float64)
        n = M.shape[0]
        outer = np.arange(0, n * n + 1, n, dtype=np.int64)
        inner = np.tile(np.arange(n, dtype=np.int32), n)
        return cls(n, outer, inner, M.T.reshape(-1).copy(), "col", mode)  # column-major storage

    def __del__(self):
        try:
            if self.h:
                lib().oracle_csr_free(self.h)
                self.h = None
        except Exception:
            pass

    @property
    def nnz(self):
        return int(lib().oracle_csr_nnz(self.h))

    def export(self):
        nnz = self.nnz
        rp = np.empty(self.n + 1, np.int64)
        ci = np.empty(nnz, np.int32)
        v = np.empty(nnz, np.float64)
        lib().oracle_csr_export(self.h, _p(rp), _p(ci), _p(v))
        return rp, ci, v

    def to_scipy(self):
        import scipy.sparse as sp

        rp, ci, v = self.export()
        return sp.csr_matrix((v, ci, rp), shape=(self.n, self.n))

    def set_threads(self, t):
        lib().oracle_csr_set_threads(self.h, int(t))

    def spmv(self, x):
        x = _f64(x)
        y = np.empty(self.n)
        lib().oracle_spmv(self.h, _p(x), _p(y))
        return y


def simple_random(seed: int, n: int) -> np.ndarray:
    out = np.empty(n)
    lib().oracle_simple_random(C.c_uint64(seed), C.c_int64(n), _p(out))
    return out


def gen_sparse_data(n: int, prob: float = 0.5):
    """The reference tests' gen_sparse_data (test/SymEigs.cpp:25-42): returns a scipy CSC matrix
    (Eigen::SparseMatrix<double> default storage) that is NOT symmetric; the Sym op reads its lower triangle."""
    import scipy.sparse as sp

    r = np.empty(n * n, np.int32)
    c = np.empty(n * n, np.int32)
    v = np.empty(n * n, np.float64)
    cnt = lib().oracle_gen_sparse_data(C.c_int64(n), C.c_double(prob), _p(r), _p(c), _p(v))
    return sp.csc_matrix((v[:cnt], (r[:cnt], c[:cnt])), shape=(n, n))


def givens(x, y):
    r, c, s = C.c_double(), C.c_double(), C.c_double()
    lib().oracle_givens(C.c_double(x), C.c_double(y), C.byref(r), C.byref(c), C.byref(s))
    return r.value, c.value, s.value


def _colmajor(M):
    return np.asfortranarray(np.asarray(M, dtype=np.float64))


def shifted_qr(H, shift, kind="tridiag"):
    """TridiagQR / UpperHessenbergQR: returns (R, QtHQ, Q) with Q = G1*G2*..."""
    H = _colmajor(H)
    m = H.shape[0]
    R, D, Q = (np.empty((m, m), order="F") for _ in range(3))
    _check(lib().oracle_shifted_qr(0 if kind == "tridiag" else 1, C.c_int64(m), _p(H), C.c_double(shift), _p(R), _p(D), _p(Q)))
    return R, D, Q


def double_shift_qr(H, s, t):
    H = _colmajor(H)
    m = H.shape[0]
    D, Q = (np.empty((m, m), order="F") for _ in range(2))
    _check(lib().oracle_double_shift_qr(C.c_int64(m), _p(H), C.c_double(s), C.c_double(t), _p(D), _p(Q)))
    return D, Q


def tridiag_eigen(H):
    H = _colmajor(H)
    m = H.shape[0]
    ev = np.empty(m)
    V = np.empty((m, m), order="F")
    _check(lib().oracle_tridiag_eigen(C.c_int64(m), _p(H), _p(ev), _p(V)))
    return ev, V


def hess_schur(H):
    H = _colmajor(H)
    m = H.shape[0]
    T, U = (np.empty((m, m), order="F") for _ in range(2))
    _check(lib().oracle_hess_schur(C.c_int64(m), _p(H), _p(T), _p(U)))
    return T, U


def hess_eigen(H):
    H = _colmajor(H)
    m = H.shape[0]
    ev = np.empty(2 * m)
    V = np.empty(2 * m * m)
    _check(lib().oracle_hess_eigen(C.c_int64(m), _p(H), _p(ev), _p(V)))
    evc = ev[0::2] + 1j * ev[1::2]
    Vc = (V[0::2] + 1j * V[1::2]).reshape((m, m), order="F")
    return evc, Vc


def argsort(selection, values):
    values = np.asarray(values)
    n = len(values)
    ind = np.empty(n, np.int64)
    if np.iscomplexobj(values):
        ri = np.empty(2 * n)
        ri[0::2] = values.real
        ri[1::2] = values.imag
        _check(lib().oracle_argsort_complex(int(selection), _p(ri), C.c_int64(n), _p(ind)))
    else:
        v = _f64(values)
        _check(lib().oracle_argsort(int(selection), _p(v), C.c_int64(n), _p(ind)))
    return ind


def factorize(csr: Csr, m: int, v0=None, mid=None, kind="lanczos"):
    """init + factorize_from(1, mid) + factorize_from(mid, m) as test/Arnoldi.cpp does."""
    n = csr.n
    mid = m // 2 if mid is None else mid
    V = np.empty((n, m), order="F")
    H = np.empty((m, m), order="F")
    f = np.empty(n)
    beta = C.c_double()
    nops = C.c_int64()
    stats = np.zeros(4, np.int64)
    v0a = _f64(v0) if v0 is not None else None
    _check(lib().oracle_factorize(0 if kind == "lanczos" else 1, csr.h, C.c_int64(m), _p(v0a), C.c_int64(mid), _p(V), _p(H), _p(f), C.byref(beta),
                                  C.byref(nops), _p(stats)))
    return dict(V=V, H=H, f=f, beta=beta.value, nops=nops.value, reorth_passes=int(stats[0]), expand_calls=int(stats[1]), steps=int(stats[3]))


def sym_restart_prepare(H, beta, nev, selection, tol):
    H = _colmajor(H)
    m = H.shape[0]
    rv, re = np.empty(m), np.empty(m)
    rvec = np.empty((m, nev), order="F")
    conv = np.empty(nev, np.int32)
    nconv, k = C.c_int64(), C.c_int64()
    Q, Hn = (np.empty((m, m), order="F") for _ in range(2))
    _check(lib().oracle_sym_restart_prepare(C.c_int64(m), _p(H), C.c_double(beta), C.c_int64(nev), int(selection), C.c_double(tol), _p(rv), _p(re), _p(rvec),
                                            _p(conv), C.byref(nconv), C.byref(k), _p(Q), _p(Hn)))
    return dict(ritz_val=rv, ritz_est=re, ritz_vec=rvec, conv=conv, nconv=nconv.value, k=k.value, Q=Q, H=Hn)


@dataclass
class EigsResult:
    nconv: int
    niter: int
    nops: int
    info: int
    eigenvalues: np.ndarray
    eigenvectors: np.ndarray | None
    reorth_passes: int
    expand_calls: int
    restarts: int
    steps: int
    seconds: float


def _wrap(res: _Result, evals, evecs, nev, complex_out=False):
    nconv = int(res.nconv)
    if complex_out:
        ev = (evals[0::2] + 1j * evals[1::2])[:nconv]
        V = None
        if evecs is not None:
            n = evecs.size // (2 * nev)
            V = (evecs[0::2] + 1j * evecs[1::2]).reshape((n, nev), order="F")[:, :nconv]
    else:
        ev = evals[:nconv].copy()
        V = evecs[:, :nconv].copy(order="F") if evecs is not None else None
    return EigsResult(nconv, int(res.niter), int(res.nops), int(res.info), ev, V, int(res.reorth_passes), int(res.expand_calls), int(res.restarts),
                      int(res.steps), float(res.seconds))


def sym_eigs(csr: Csr, nev, ncv, selection=LargestMagn, maxit=1000, tol=1e-10, sorting=LargestAlge, init_resid=None, sigma=None, threads=1, op_limit=-1,
             want_vectors=True) -> EigsResult:
    """SymEigsSolver<SparseSymMatProd>: init()/init(resid) + compute(...).  sigma != None applies the
    SymEigsShiftSolver back-transform (the op must then be the shift-solve operator)."""
    n = csr.n
    ncv_eff = min(ncv, n)
    evals = np.zeros(nev)
    evecs = np.zeros((n, nev), order="F") if want_vectors and op_limit < 0 else None
    r0 = _f64(init_resid) if init_resid is not None else None
    res = _Result()
    _check(lib().oracle_sym_eigs(csr.h, C.c_int64(nev), C.c_int64(ncv), int(selection), C.c_int64(maxit), C.c_double(tol), int(sorting), _p(r0),
                                 1 if sigma is not None else 0, C.c_double(sigma or 0.0), int(threads), C.c_int64(op_limit), _p(evals), _p(evecs), C.byref(res)))
    del ncv_eff
    return _wrap(res, evals, evecs, nev)


class BandLu:
    """SparseSymShiftSolve on a banded matrix: set_shift at construction, perform_op = solve (oracle/band.hpp)."""

    def __init__(self, csr: Csr, sigma: float):
        self.n, self.sigma = csr.n, float(sigma)
        h = C.c_void_p()
        _check(lib().oracle_band_create(csr.h, C.c_double(sigma), C.byref(h)))
        self.h = h
        kl, ku = C.c_int64(), C.c_int64()
        lib().oracle_band_info(self.h, C.byref(kl), C.byref(ku))
        self.kl, self.ku = kl.value, ku.value

    def perform_op(self, x):
        x = _f64(x)
        y = np.empty(self.n)
        _check(lib().oracle_band_solve(self.h, _p(x), _p(y)))
        return y

    def __del__(self):
        if getattr(self, "h", None):
            lib().oracle_band_destroy(self.h)
            self.h = None


def sym_shift_eigs(op: BandLu, nev, ncv, selection=LargestMagn, maxit=1000, tol=1e-10, sorting=LargestAlge, init_resid=None, op_limit=-1,
                   want_vectors=True) -> EigsResult:
    """SymEigsShiftSolver<SparseSymShiftSolve> (SymEigsShiftSolver.h:148-196): Lanczos on (A - sigma I)^{-1}, eigenvalues mapped back by 1/nu + sigma."""
    n = op.n
    evals = np.zeros(nev)
    evecs = np.zeros((n, nev), order="F") if want_vectors and op_limit < 0 else None
    r0 = _f64(init_resid) if init_resid is not None else None
    res = _Result()
    _check(lib().oracle_sym_shift_eigs(op.h, C.c_double(op.sigma), C.c_int64(nev), C.c_int64(ncv), int(selection), C.c_int64(maxit), C.c_double(tol), int(sorting),
                                       _p(r0), C.c_int64(op_limit), _p(evals), _p(evecs), C.byref(res)))
    return _wrap(res, evals, evecs, nev)


_USERFN = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_void_p)


def sym_eigs_userop(n, fn, nev, ncv, selection=LargestMagn, maxit=1000, tol=1e-10, sorting=LargestAlge, init_resid=None, sigma=None) -> EigsResult:
    """Same driver with a user-defined OpType (y = fn(x) on numpy vectors)."""

    def tramp(xp, yp, _):
        x = np.ctypeslib.as_array(xp, shape=(n,))
        y = np.ctypeslib.as_array(yp, shape=(n,))
        y[:] = fn(x)

    cb = _USERFN(tramp)
    evals = np.zeros(nev)
    evecs = np.zeros((n, nev), order="F")
    r0 = _f64(init_resid) if init_resid is not None else None
    res = _Result()
    _check(lib().oracle_sym_eigs_userop(C.c_int64(n), cb, None, C.c_int64(nev), C.c_int64(ncv), int(selection), C.c_int64(maxit), C.c_double(tol), int(sorting),
                                        _p(r0), 1 if sigma is not None else 0, C.c_double(sigma or 0.0), _p(evals), _p(evecs), C.byref(res)))
    return _wrap(res, evals, evecs, nev)


def gen_eigs(csr: Csr, nev, ncv, selection=LargestMagn, maxit=1000, tol=1e-10, sorting=LargestMagn, init_resid=None, threads=1, op_limit=-1,
             want_vectors=True) -> EigsResult:
    n = csr.n
    evals = np.zeros(2 * nev)
    evecs = np.zeros(2 * n * nev) if want_vectors and op_limit < 0 else None
    r0 = _f64(init_resid) if init_resid is not None else None
    res = _Result()
    _check(lib().oracle_gen_eigs(csr.h, C.c_int64(nev), C.c_int64(ncv), int(selection), C.c_int64(maxit), C.c_double(tol), int(sorting), _p(r0), int(threads),
                                 C.c_int64(op_limit), _p(evals), _p(evecs), C.byref(res)))
    return _wrap(res, evals, evecs, nev, complex_out=True)

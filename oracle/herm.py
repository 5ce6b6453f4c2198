"""CPU oracle of the complex Hermitian path: HermEigsSolver<SparseHermMatProd<std::complex<double>>>.

TEST INFRASTRUCTURE ONLY -- imported by tests/ (and nothing in spectra_b200/).  A numpy restatement of the reference's control flow
with Scalar = std::complex<double>; citations are relative to /root/reference/include/Spectra/:

  SparseHermMatProd::perform_op   MatOp/SparseHermMatProd.h:83-88   (selfadjointView<Uplo>: one triangle, mirrored conjugated)
  ArnoldiOp<Op, IdentityBOp>      MatOp/internal/ArnoldiOp.h:136-155 (x.dot(y) = x^H y, X.adjoint() * y, norm)
  Arnoldi::init / expand_basis    LinAlg/Arnoldi.h:136-195, 66-115
  Lanczos::factorize_from         LinAlg/Lanczos.h:62-187           (H is a complex matrix whose imaginary parts are rounding noise)
  HermEigsBase                    HermEigsBase.h:105-155 restart (works on m_fac_H.real()), 158-202, 205-224, 229-251, 366-390, 417-470
  SimpleRandom<complex>           Util/SimpleRandom.h:68-77          (re, im drawn consecutively from the one LCG stream)

The real-symmetric m x m restart arithmetic (TridiagEigen, argsort, num_converged, nev_adjusted, the shifted TridiagQR chain) is the
C++ oracle's (`oracle.sym_restart_prepare`, pinned by tests/test_oracle.py); vector work is numpy on complex128 arrays, so the loops
below run once per Lanczos step, not per element.

Pinning (tests/test_oracle.py): dense truth `numpy.linalg.eigh` on the reference's own fixtures (test/HermEigs.cpp:27-50, 118-174)
for every selection rule with the reference's acceptance threshold ||AU - UD||_inf <= 1e-9.  Like everything Eigen-backed, the
reference's bits are not pinned (Eigen is not available here): parity is by those tolerances.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import (BothEnds, LargestAlge, LargestMagn, SmallestAlge, SmallestMagn, _check, _p, lib, simple_random, sym_restart_prepare)

EPS = 2.220446049250313e-16
NEAR0 = 2.2250738585072014e-308 * 10.0


def gen_sparse_data_herm(n: int, prob: float = 0.5):
    """test/HermEigs.cpp:27-50: complex sparse matrix (NOT Hermitian; the op reads its lower triangle), real diagonal.  CSC."""
    import scipy.sparse as sp

    r = np.empty(n * n, np.int32)
    c = np.empty(n * n, np.int32)
    v = np.empty(2 * n * n, np.float64)
    cnt = lib().oracle_gen_sparse_data_herm(C.c_int64(n), C.c_double(prob), _p(r), _p(c), _p(v))
    vals = v[0:2 * cnt:2] + 1j * v[1:2 * cnt:2]
    return sp.csc_matrix((vals, (r[:cnt], c[:cnt])), shape=(n, n))


def herm_full(A, uplo: str = "lower"):
    """selfadjointView<Uplo> of a stored complex matrix as an explicit Hermitian scipy CSR matrix (diagonal taken as real)."""
    import scipy.sparse as sp

    A = sp.csr_matrix(A)
    T = sp.tril(A, -1) if uplo == "lower" else sp.triu(A, 1)
    D = sp.diags(A.diagonal().real.astype(np.complex128))
    return (T + T.conj().T + D).tocsr()


def simple_random_complex(seed: int, n: int) -> np.ndarray:
    """SimpleRandom<std::complex<double>>(seed).random_vec(n) (Util/SimpleRandom.h:68-77, 106-113)"""
    r = simple_random(seed, 2 * n)
    return r[0::2] + 1j * r[1::2]


@dataclass
class HermResult:
    nconv: int
    niter: int
    nops: int
    info: int               # CompInfo: 0 Successful, 2 NotConverging
    eigenvalues: np.ndarray
    eigenvectors: np.ndarray | None
    H: np.ndarray           # last tridiagonal matrix (real)
    reorth_passes: int
    expand_calls: int


class _Lanczos:
    def __init__(self, op, n, m):
        self.op, self.n, self.m = op, n, m
        self.V = np.zeros((n, m), dtype=np.complex128, order="F")
        self.H = np.zeros((m, m), dtype=np.complex128, order="F")
        self.f = np.zeros(n, dtype=np.complex128)
        self.beta = 0.0
        self.k = 0
        self.nops = 0
        self.reorth = 0
        self.expands = 0

    def matvec(self, x):
        self.nops += 1
        return self.op(x)

    # Arnoldi.h:136-195
    def init(self, v0):
        v0norm = np.linalg.norm(v0)
        if v0norm < NEAR0:
            raise ValueError("initial residual vector cannot be zero")
        v = self.matvec(v0)
        vnorm = np.linalg.norm(v)
        v = v0 / v0norm if vnorm < NEAR0 else v / vnorm
        self.V[:, 0] = v
        w = self.matvec(v)
        self.H[:] = 0
        self.H[0, 0] = np.vdot(v, w)
        self.f = w - v * self.H[0, 0]
        if np.abs(self.f).max() < EPS * abs(self.H[0, 0]):
            self.f[:] = 0
            self.beta = 0.0
        else:
            self.beta = float(np.linalg.norm(self.f))
        self.k = 1

    # Arnoldi.h:66-115
    def expand_basis(self, i, seed):
        self.expands += 1
        V = self.V[:, :i]
        for it in range(5):
            rnd = simple_random_complex(seed + 123 * it, self.n)
            f = self.matvec(rnd) if it == 0 else rnd
            Vf = V.conj().T @ f
            f = f - V @ Vf
            fnorm = float(np.linalg.norm(f))
            Vf = V.conj().T @ f
            err = np.abs(Vf).max() if i > 0 else 0.0
            count = 0
            while count < 3 and err >= EPS * fnorm:
                f = f - V @ Vf
                fnorm = float(np.linalg.norm(f))
                Vf = V.conj().T @ f
                err = np.abs(Vf).max()
                count += 1
            self.f, self.beta = f, fnorm
            if err < EPS * fnorm:
                return

    # Lanczos.h:62-187
    def factorize_from(self, from_k, to_m):
        if to_m <= from_k:
            return
        if from_k > self.k:
            raise ValueError("Lanczos: from_k is larger than the current subspace dimension")
        beta_thresh = EPS * np.sqrt(self.n)
        eps_sqrt = np.sqrt(EPS)
        self.H[:, from_k:] = 0
        self.H[from_k:, :from_k] = 0
        for i in range(from_k, to_m):
            restart = self.beta < NEAR0
            if not restart:
                v = self.f / self.beta
                if self.beta < eps_sqrt:
                    restart = abs(np.vdot(self.V[:, i - 1], v)) > eps_sqrt
            if restart:
                self.expand_basis(i, 2 * i)
                v = self.f / self.beta
            self.V[:, i] = v
            self.H[i, i - 1] = 0.0 if restart else self.beta
            self.H[i - 1, i] = self.H[i, i - 1]
            w = self.matvec(v)
            if not restart:
                w = w - self.H[i, i - 1] * self.V[:, i - 1]
            self.H[i, i] = np.vdot(v, w)
            self.f = w - self.H[i, i] * v
            self.beta = float(np.linalg.norm(self.f))
            Vs = self.V[:, :i + 1]
            Vf = Vs.conj().T @ self.f
            err = np.abs(Vf).max()
            count = 0
            while count < 5 and err > EPS * self.beta:
                if self.beta < beta_thresh:
                    self.f[:] = 0
                    self.beta = 0.0
                    break
                self.f = self.f - Vs @ Vf
                self.H[i - 1, i] += Vf[i - 1]
                self.H[i, i - 1] = self.H[i - 1, i]
                self.H[i, i] += Vf[i]
                self.beta = float(np.linalg.norm(self.f))
                Vf = Vs.conj().T @ self.f
                err = np.abs(Vf).max()
                count += 1
                self.reorth += 1
        self.k = to_m

    # Arnoldi.h:320-340 (Q real) + Lanczos::compress_H (H <- Q'HQ, taken from the decomposition)
    def compress(self, Q, Hnew, k):
        m = self.m
        Vs = np.empty((self.n, k + 1), dtype=np.complex128, order="F")
        for i in range(k):
            nnz = m - k + i + 1
            Vs[:, i] = self.V[:, :nnz] @ Q[:nnz, i]
        Vs[:, k] = self.V @ Q[:, k]
        self.V[:, :k + 1] = Vs
        self.H = Hnew.astype(np.complex128)
        self.f = self.f * Q[m - 1, k - 1] + self.V[:, k] * self.H[k, k - 1]
        self.beta = float(np.linalg.norm(self.f))
        self.k = k


def _argsort(rule, vals):
    if rule == LargestAlge:
        return np.argsort(-vals, kind="stable")
    if rule == LargestMagn:
        return np.argsort(-np.abs(vals), kind="stable")
    if rule == SmallestAlge:
        return np.argsort(vals, kind="stable")
    if rule == SmallestMagn:
        return np.argsort(np.abs(vals), kind="stable")
    raise ValueError("unsupported sorting rule")


def herm_eigs(op, n, nev, ncv, selection=LargestMagn, maxit=1000, tol=1e-10, sorting=LargestAlge, init_resid=None, want_vectors=True) -> HermResult:
    """HermEigsSolver(op, nev, ncv): init() / init(resid) + compute(selection, maxit, tol, sorting).  `op` maps a complex128
    vector to A x (use `herm_full(A).dot` for the SparseHermMatProd semantics)."""
    if nev < 1 or nev > n - 1:
        raise ValueError("nev must satisfy 1 <= nev <= n - 1, n is the size of matrix")
    if ncv <= nev or ncv > n:
        raise ValueError("ncv must satisfy nev < ncv <= n, n is the size of matrix")
    if selection not in (LargestMagn, LargestAlge, SmallestMagn, SmallestAlge, BothEnds):
        raise ValueError("unsupported selection rule")
    m = ncv
    fac = _Lanczos(op, n, m)
    fac.init(simple_random_complex(0, n) if init_resid is None else np.asarray(init_resid, dtype=np.complex128))
    fac.factorize_from(1, m)
    niter, nconv = 0, 0
    prep = None
    i = 0
    for i in range(maxit):
        # retrieve_ritzpair + num_converged + nev_adjusted + the shift loop of restart(), on H.real()
        prep = sym_restart_prepare(np.ascontiguousarray(fac.H.real), fac.beta, nev, selection, tol)
        nconv = prep["nconv"]
        if nconv >= nev:
            break
        k = prep["k"]
        if k < m:
            fac.compress(prep["Q"], prep["H"], k)
            fac.factorize_from(k, m)
    else:
        i = maxit
    conv = prep["conv"].astype(bool)
    if i == maxit:
        # the last restart() ended with retrieve_ritzpair(); the convergence flags stay those of the last num_converged()
        prep2 = sym_restart_prepare(np.ascontiguousarray(fac.H.real), fac.beta, nev, selection, tol)
        prep = dict(prep2, conv=prep["conv"])
    niter = i + 1
    rv, rvec = prep["ritz_val"][:nev], prep["ritz_vec"]
    ind = _argsort(sorting, rv)
    rv, rvec, conv = rv[ind], rvec[:, ind], conv[ind]
    evals = rv[conv]
    U = (fac.V @ rvec[:, conv]) if want_vectors else None
    return HermResult(int(min(nev, nconv)), niter, fac.nops, 0 if nconv >= nev else 2, evals, U, fac.H.real.copy(), fac.reorth, fac.expands)


def herm_factorize(op, n, m, v0=None, mid=None):
    """test/Arnoldi.cpp flow for the Hermitian Lanczos factorisation: init, factorize_from(1, mid), factorize_from(mid, m)."""
    fac = _Lanczos(op, n, m)
    fac.init(simple_random_complex(0, n) if v0 is None else np.asarray(v0, dtype=np.complex128))
    mid = m // 2 if mid is None else mid
    fac.factorize_from(1, mid)
    fac.factorize_from(mid, m)
    return dict(V=fac.V, H=fac.H, f=fac.f, beta=fac.beta, nops=fac.nops)


def arnoldi_factorize_complex(op, n, m, v0=None, mid=None):
    """Arnoldi<complex>::init + factorize_from(1, mid) + factorize_from(mid, m) (LinAlg/Arnoldi.h:136-195, 198-295), the complex case of
    test/Arnoldi.cpp:19-85.  `op` maps a complex128 vector to A x for a general complex A."""
    fac = _Lanczos(op, n, m)  # shares init / expand_basis / state with the Hermitian oracle
    fac.init(simple_random_complex(0, n) if v0 is None else np.asarray(v0, dtype=np.complex128))
    mid = m // 2 if mid is None else mid
    beta_thresh = EPS * np.sqrt(n)

    def factorize_from(from_k, to_m):
        if to_m <= from_k:
            return
        fac.H[:, from_k:] = 0
        fac.H[from_k:, :from_k] = 0
        for i in range(from_k, to_m):
            restart = False
            if fac.beta < NEAR0:
                fac.expand_basis(i, 2 * i)
                restart = True
            v = fac.f / fac.beta
            fac.V[:, i] = v
            fac.H[i, i - 1] = 0.0 if restart else fac.beta
            w = fac.matvec(v)
            Vs = fac.V[:, :i + 1]
            h = Vs.conj().T @ w
            fac.H[:i + 1, i] = h
            fac.f = w - Vs @ h
            fac.beta = float(np.linalg.norm(fac.f))
            if fac.beta > 0.717 * np.linalg.norm(h):
                continue
            Vf = Vs.conj().T @ fac.f
            err = np.abs(Vf).max()
            count = 0
            while count < 5 and err > EPS * fac.beta:
                if fac.beta < beta_thresh:
                    fac.f[:] = 0
                    fac.beta = 0.0
                    break
                fac.f = fac.f - Vs @ Vf
                fac.H[:i + 1, i] += Vf
                fac.beta = float(np.linalg.norm(fac.f))
                Vf = Vs.conj().T @ fac.f
                err = np.abs(Vf).max()
                count += 1
                fac.reorth += 1
        fac.k = to_m

    factorize_from(1, mid)
    factorize_from(mid, m)
    return dict(V=fac.V, H=fac.H, f=fac.f, beta=fac.beta, nops=fac.nops)


# =============================================================================================================================
# Complex GenEigsSolver: GenEigsBase.h with Scalar = std::complex<double>
#   Givens<complex>::compute_rotation     LinAlg/Givens.h:218-335 (Algorithm 1 branches) + StableScaling :28-86, real Givens :166-205
#   UpperHessenbergQR<complex>            LinAlg/UpperHessenbergQR.h:136-195 compute, :219-255 matrix_QtHQ (RQ + sI), :383-417 apply_YQ
#   RestartArnoldi<complex>::run          GenEigsBase.h:122-139 (one complex shift at a time)
#   restart / num_converged / nev_adjusted / retrieve_ritzpair / sort_ritzpair / compute   GenEigsBase.h:204-277, 280-404, 501-525
#   compress_V                            LinAlg/Arnoldi.h:320-340 (complex Q)
# The Ritz pairs of the m x m complex Hessenberg matrix come from LAPACK (numpy.linalg.eig) instead of a restatement of
# Eigen::ComplexSchur (third party, UpperHessenbergEigen.h:328-454): same eigenvalues; eigenvectors normalised to unit 2-norm as in
# :383-387, their phase is irrelevant to everything downstream (|last component| enters the convergence test, V*s the result).
# =============================================================================================================================
def _stable_scaling(a, b):
    """StableScaling::run (Givens.h:28-62): a >= b > 0 -> (r, c, s) = (sqrt(a^2 + b^2), a / r, b / r), Taylor branch for tiny b / a"""
    t = b / a
    if t >= 0.1 * EPS ** 0.25:
        r = float(np.hypot(a, b))
        return r, a / r, b / r
    t2 = t * t
    c = 1.0 - t2 * (0.5 - t2 * (0.375 - 0.3125 * t2))
    return a + 0.5 * b * t * (1.0 - t2 * (0.25 - 0.125 * t2)), c, t * c


def _stable_scaling_complex(a, b):
    """StableScaling::run(Complex a, Complex b, a2, tc1, tc2) (Givens.h:64-92): a2 = |a|^2, tc1 = sqrt(1 + t2), tc2 = 1 / tc1, t2 = |b|^2 / |a|^2"""
    b2 = b.real * b.real + b.imag * b.imag
    a2 = a.real * a.real + a.imag * a.imag
    t2 = b2 / a2
    if t2 >= 0.1 * np.sqrt(EPS):
        return a2, float(np.sqrt(1.0 + t2)), float(np.sqrt(a2 / (a2 + b2)))
    return a2, 1.0 + t2 * (0.5 - t2 * (0.125 - 0.0625 * t2)), 1.0 - t2 * (0.5 - t2 * (0.375 - 0.3125 * t2))


def givens_real(x, y):
    """Givens<double>::compute_rotation (Givens.h:166-205): c*x - s*y = r >= 0, s*x + c*y = 0"""
    xs, ys = (1.0 if x >= 0 else -1.0), (1.0 if y >= 0 else -1.0)
    xa, ya = abs(x), abs(y)
    if xa > ya:
        if ya == 0.0:
            return xa, xs, 0.0
        r, c, s = _stable_scaling(xa, ya)
        return r, xs * c, -ys * s
    if xa == 0.0 and ya == 0.0:
        return 0.0, 1.0, 0.0
    if xa == 0.0:
        return ya, 0.0, -ys
    r, s, c = _stable_scaling(ya, xa)
    return r, xs * c, -ys * s


def givens_complex(x, y):
    """Givens<complex>::compute_rotation (Givens.h:218-335): real c, complex s, r with  c*x - s*y = r,  conj(s)*x + c*y = 0"""
    if y == 0:
        return x, 1.0, 0j
    if x == 0:
        rr, sr, si = givens_real(-y.real, -y.imag)
        return complex(rr, 0.0), 0.0, complex(sr, si)
    xn1, yn1 = abs(x.real) + abs(x.imag), abs(y.real) + abs(y.imag)
    if xn1 > yn1:
        x2, tc1, tc2 = _stable_scaling_complex(x, y)
        c = tc2
        return tc1 * x, c, -(c / x2) * (x * np.conj(y))
    rho = np.sqrt(abs(x) ** 2 + abs(y) ** 2)
    xnorm, zr, zi = givens_real(x.real, -x.imag)
    z = complex(zr, zi)
    return rho * z, xnorm / rho, -(z * np.conj(y)) / rho


def hess_qr_complex(H, shift):
    """UpperHessenbergQR<complex>: returns (RQ + shift*I, cos[], sin[]) of H - shift*I = QR"""
    n = H.shape[0]
    R = H.astype(np.complex128).copy()
    R[np.diag_indices(n)] -= shift
    cs, sn = np.zeros(n - 1), np.zeros(n - 1, dtype=np.complex128)
    for i in range(n - 1):
        R[i + 2:, i] = 0
        r, c, s = givens_complex(complex(R[i, i]), complex(R[i + 1, i]))
        cs[i], sn[i] = c, s
        R[i, i], R[i + 1, i] = r, 0
        t0, t1 = R[i, i + 1:].copy(), R[i + 1, i + 1:].copy()
        R[i, i + 1:] = c * t0 - s * t1
        R[i + 1, i + 1:] = np.conj(s) * t0 + c * t1
    RQ = R
    for i in range(n - 1):
        c, s = cs[i], sn[i]
        a, b = RQ[:i + 2, i].copy(), RQ[:i + 2, i + 1].copy()
        RQ[:i + 2, i] = c * a - np.conj(s) * b
        RQ[:i + 2, i + 1] = s * a + c * b
    RQ[np.diag_indices(n)] += shift
    return RQ, cs, sn


def apply_yq_complex(Y, cs, sn):
    for i in range(len(cs)):
        c, s = cs[i], sn[i]
        a, b = Y[:, i].copy(), Y[:, i + 1].copy()
        Y[:, i] = c * a - np.conj(s) * b
        Y[:, i + 1] = s * a + c * b


def _sort_key_complex(rule, v):
    from . import LargestImag, LargestReal, SmallestImag, SmallestReal  # noqa: PLC0415

    return {LargestMagn: -np.abs(v), LargestReal: -v.real, LargestImag: -np.abs(v.imag), SmallestMagn: np.abs(v), SmallestReal: v.real,
            SmallestImag: np.abs(v.imag)}[rule]


@dataclass
class GenResultZ:
    nconv: int
    niter: int
    nops: int
    info: int
    eigenvalues: np.ndarray
    eigenvectors: np.ndarray | None


class _ArnoldiZ(_Lanczos):
    def factorize_from(self, from_k, to_m):  # Arnoldi.h:198-295
        if to_m <= from_k:
            return
        beta_thresh = EPS * np.sqrt(self.n)
        self.H[:, from_k:] = 0
        self.H[from_k:, :from_k] = 0
        for i in range(from_k, to_m):
            restart = False
            if self.beta < NEAR0:
                self.expand_basis(i, 2 * i)
                restart = True
            v = self.f / self.beta
            self.V[:, i] = v
            self.H[i, i - 1] = 0.0 if restart else self.beta
            w = self.matvec(v)
            Vs = self.V[:, :i + 1]
            h = Vs.conj().T @ w
            self.H[:i + 1, i] = h
            self.f = w - Vs @ h
            self.beta = float(np.linalg.norm(self.f))
            if self.beta > 0.717 * np.linalg.norm(h):
                continue
            Vf = Vs.conj().T @ self.f
            err = np.abs(Vf).max()
            count = 0
            while count < 5 and err > EPS * self.beta:
                if self.beta < beta_thresh:
                    self.f[:] = 0
                    self.beta = 0.0
                    break
                self.f = self.f - Vs @ Vf
                self.H[:i + 1, i] += Vf
                self.beta = float(np.linalg.norm(self.f))
                Vf = Vs.conj().T @ self.f
                err = np.abs(Vf).max()
                count += 1
                self.reorth += 1
        self.k = to_m


def gen_eigs_complex(op, n, nev, ncv, selection=LargestMagn, maxit=1000, tol=1e-10, sorting=LargestMagn, init_resid=None) -> GenResultZ:
    """GenEigsSolver<complex op>(op, nev, ncv): init() + compute(selection, maxit, tol, sorting)   (GenEigsBase.h:409-525)"""
    if nev < 1 or nev > n - 2:
        raise ValueError("nev must satisfy 1 <= nev <= n - 2, n is the size of matrix")
    if ncv < nev + 2 or ncv > n:
        raise ValueError("ncv must satisfy nev + 2 <= ncv <= n, n is the size of matrix")
    m = ncv
    fac = _ArnoldiZ(op, n, m)
    fac.init(simple_random_complex(0, n) if init_resid is None else np.asarray(init_resid, dtype=np.complex128))
    fac.factorize_from(1, m)
    eps23 = EPS ** (2.0 / 3.0)

    def retrieve():
        ev, Z = np.linalg.eig(fac.H)
        Z = Z / np.linalg.norm(Z, axis=0)
        # UpperHessenbergEigen<complex>::sortEigenvalues (:389-404): ascending modulus, then the selection rule (stable here)
        o = np.argsort(np.abs(ev), kind="stable")
        ev, Z = ev[o], Z[:, o]
        ind = np.argsort(_sort_key_complex(selection, ev), kind="stable")
        return ev[ind], Z[m - 1, ind], Z[:, ind[:nev]]

    ritz_val, ritz_est, ritz_vec = retrieve()
    nconv, i = 0, 0
    conv = np.zeros(nev, dtype=bool)
    for i in range(maxit):
        thresh = tol * np.maximum(np.abs(ritz_val[:nev]), eps23)
        conv = np.abs(ritz_est[:nev]) * fac.beta < thresh
        nconv = int(conv.sum())
        if nconv >= nev:
            break
        # nev_adjusted (GenEigsBase.h:245-277)
        k = nev + int(np.sum(np.abs(ritz_est[nev:]) < NEAR0))
        k += min(nconv, (m - k) // 2)
        if k == 1 and m >= 6:
            k = m // 2
        elif k == 1 and m > 3:
            k = 2
        k = min(k, m - 2)
        if ritz_val[k - 1].imag != 0 and ritz_val[k - 1] == np.conj(ritz_val[k]):
            k += 1
        # restart(k): one complex shift per unwanted Ritz value, V <- V Q, expand again
        if k < m:
            Q = np.eye(m, dtype=np.complex128)
            for j in range(k, m):
                fac.H, cs, sn = hess_qr_complex(fac.H, ritz_val[j])
                apply_yq_complex(Q, cs, sn)
            Vs = np.empty((n, k + 1), dtype=np.complex128, order="F")
            for c in range(k):
                nnz = m - k + c + 1
                Vs[:, c] = fac.V[:, :nnz] @ Q[:nnz, c]
            Vs[:, k] = fac.V @ Q[:, k]
            fac.V[:, :k + 1] = Vs
            fac.f = fac.f * Q[m - 1, k - 1] + fac.V[:, k] * fac.H[k, k - 1]
            fac.beta = float(np.linalg.norm(fac.f))
            fac.k = k
            fac.factorize_from(k, m)
            ritz_val, ritz_est, ritz_vec = retrieve()
    else:
        i = maxit
    niter = i + 1
    ind = np.argsort(_sort_key_complex(sorting, ritz_val[:nev]), kind="stable")
    rv, rvec, conv = ritz_val[:nev][ind], ritz_vec[:, ind], conv[ind]
    return GenResultZ(int(min(nev, nconv)), niter, fac.nops, 0 if nconv >= nev else 2, rv[conv], fac.V @ rvec[:, conv])


def gen_sparse_data_complex(n: int, prob: float = 0.5):
    """test/ComplexEigs.cpp:20-39: general complex sparse matrix, same engine; real and imaginary part drawn for every entry.  CSC."""
    import scipy.sparse as sp

    r = np.empty(n * n, np.int32)
    c = np.empty(n * n, np.int32)
    v = np.empty(2 * n * n, np.float64)
    cnt = lib().oracle_gen_sparse_data_complex(C.c_int64(n), C.c_double(prob), _p(r), _p(c), _p(v))
    vals = v[0:2 * cnt:2] + 1j * v[1:2 * cnt:2]
    return sp.csc_matrix((vals, (r[:cnt], c[:cnt])), shape=(n, n))

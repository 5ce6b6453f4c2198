"""Parity cases of the complex Hermitian path, written against a backend module `sb` with the API of spectra_b200:
run on a device by tests/test_gpu_layouts_complex.py and on the kernel-logic emulator by tests/test_emu_kernels.py.
Tolerances: operator 1e-13 relative; factorisation identities 1e-12 (test/Arnoldi.cpp); solver ||AU - UD||_inf <= 1e-9
(test/HermEigs.cpp:66-70) and eigenvalues within 1e-10 relative of the oracle (oracle/herm.py)."""
import numpy as np
import scipy.sparse as sp

import oracle as O
from oracle import herm as OH

FIXTURES = {10: (0.5, 3, 6), 100: (0.1, 10, 20), 1000: (0.01, 20, 50)}  # test/HermEigs.cpp:140-174


def operator_case(sb, n=100, fmt="csc", uplo="lower"):
    # SparseHermMatProd::perform_op (MatOp/SparseHermMatProd.h:83-88): one triangle, mirrored conjugated
    rng = np.random.default_rng(n)
    A = OH.gen_sparse_data_herm(n, FIXTURES.get(n, (0.1,))[0])
    A = A.tocsc() if fmt == "csc" else A.tocsr()
    Af = OH.herm_full(A, uplo)
    op = sb.SparseHermMatProd(A, uplo=uplo)
    assert op.rows() == n and op.cols() == n
    for _ in range(2):
        x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        y, y0 = op.perform_op(x), Af @ x
        assert np.abs(y - y0).max() <= 1e-13 * max(1.0, np.abs(y0).max())
    M = rng.standard_normal((n, 3)) + 1j * rng.standard_normal((n, 3))
    assert np.abs(op @ M - Af @ M).max() <= 1e-13 * max(1.0, np.abs(Af @ M).max())
    assert np.array_equal(op.perform_op(x), op.perform_op(x))
    # Hermitian: <x, A x> is real
    assert abs(np.vdot(x, op.perform_op(x)).imag) <= 1e-12 * abs(np.vdot(x, op.perform_op(x)))


def factorization_case(sb, n=200, m=24):
    # test/Arnoldi.cpp:19-85 with a complex Scalar: A V - V H = f e_m', V^H V = I, H real symmetric tridiagonal
    A = sp.random(n, n, density=0.05, random_state=7, format="csc") + 1j * sp.random(n, n, density=0.05, random_state=8, format="csc")
    A = sp.csc_matrix(A)
    Af = OH.herm_full(A)
    op = sb.SparseHermMatProd(A)
    eigs = sb.HermEigsSolver(op, 3, m)
    v0 = OH.simple_random_complex(3, n)
    eigs.init(v0)
    eigs.factorize_from(1, m // 2)
    eigs.factorize_from(m // 2, m)
    fz = eigs.factorization()
    V, H, f = fz["V"], fz["H"], fz["f"]
    E = Af @ V - V @ H
    E[:, -1] -= f
    scale = max(1.0, np.abs(H).max())
    assert np.abs(E).max() <= 1e-12 * scale
    assert np.abs(V.conj().T @ V - np.eye(m)).max() <= 1e-12
    assert abs(np.linalg.norm(f) - fz["beta"]) <= 1e-12 * scale
    assert np.abs(H - H.T).max() == 0.0 and np.abs(np.triu(H, 2)).max() == 0.0
    ref = OH.herm_factorize(Af.dot, n, m, v0=v0)
    assert np.abs(H - ref["H"].real).max() <= 1e-9 * scale


def solver_case(sb, n, selection, check_history=True):
    # test/HermEigs.cpp:52-71, 118-174 (sparse cases)
    prob, k, m = FIXTURES[n]
    A = OH.gen_sparse_data_herm(n, prob)
    Af = OH.herm_full(A)
    op = sb.SparseHermMatProd(A)
    eigs = sb.HermEigsSolver(op, k, m)
    eigs.init()
    nconv = eigs.compute(selection)
    assert eigs.info() == sb.CompInfo.Successful and nconv == k
    ev, U = eigs.eigenvalues(), eigs.eigenvectors()
    assert U.dtype == np.complex128 and U.shape == (n, k)
    assert np.abs(Af @ U - U * ev).max() <= 1e-9
    ref = OH.herm_eigs(Af.dot, n, k, m, selection)
    assert ref.nconv == k
    assert np.abs(ev - ref.eigenvalues).max() <= 1e-10 * np.abs(ref.eigenvalues).max()
    if check_history and ref.niter <= 60:
        # step-for-step agreement with the oracle; long runs (SmallestMagn at n = 1000 needs ~580 restarts) may legitimately drift
        # apart by a restart when a convergence test is decided by the last bits
        assert eigs.num_operations() == ref.nops and eigs.num_iterations() == ref.niter
    assert np.abs(U.conj().T @ U - np.eye(k)).max() <= 1e-9


def argument_checks(sb):
    A = OH.gen_sparse_data_herm(10, 0.5)
    op = sb.SparseHermMatProd(A)
    for nev, ncv in ((0, 6), (10, 11), (3, 3), (3, 11)):  # HermEigsBase.h:267-271
        try:
            sb.HermEigsSolver(op, nev, ncv)
        except sb.InvalidArgument:
            continue
        raise AssertionError((nev, ncv))
    # SymEigsSolver needs a real operator, HermEigsSolver a complex one
    try:
        sb.SymEigsSolver(op, 3, 6)
        raise AssertionError("SymEigsSolver accepted a complex operator")
    except sb.InvalidArgument:
        pass
    real_op = sb.SparseSymMatProd(O.gen_sparse_data(10, 0.5))
    try:
        sb.HermEigsSolver(real_op, 3, 6)
        raise AssertionError("HermEigsSolver accepted a real operator")
    except sb.InvalidArgument:
        pass
    # a zero initial residual is rejected (Arnoldi.h:147-148)
    e = sb.HermEigsSolver(op, 3, 6)
    try:
        e.init(np.zeros(10, dtype=np.complex128))
        raise AssertionError("zero residual accepted")
    except sb.InvalidArgument:
        pass


def user_operator_case(sb, n=60):
    # a user-defined complex operator (the OpType concept with Scalar = std::complex<double>) behind HermEigsSolver
    rng = np.random.default_rng(5)
    B = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    A = (B + B.conj().T) / 2

    class MyOp:
        def rows(self):
            return n

        def perform_op(self, x_in, y_out):
            y_out[:] = A @ x_in

    op = sb.UserOp(MyOp(), complex_scalar=True)
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    assert np.abs(op.perform_op(x) - A @ x).max() <= 1e-13 * np.abs(A @ x).max()
    eigs = sb.HermEigsSolver(op, 5, 20)
    eigs.init()
    assert eigs.compute(sb.SortRule.LargestAlge) == 5
    ev, U = eigs.eigenvalues(), eigs.eigenvectors()
    w = np.linalg.eigvalsh(A)[::-1][:5]
    assert np.abs(ev - w).max() <= 1e-10 * np.abs(w).max()
    assert np.abs(A @ U - U * ev).max() <= 1e-9
    ref = OH.herm_eigs(lambda v: A @ v, n, 5, 20, O.LargestAlge)
    assert eigs.num_operations() == ref.nops


def complex_arnoldi_factorization_case(sb, n=150, m=20):
    # test/Arnoldi.cpp:19-85, complex Arnoldi: A V - V H = f e_m', V^H V = I (1e-12), H upper Hessenberg; H against the oracle
    A = sp.random(n, n, density=0.06, random_state=11, format="csc") + 1j * sp.random(n, n, density=0.06, random_state=12, format="csc")
    A = sp.csc_matrix(A)
    op = sb.SparseHermMatProd(A, uplo="general")  # every stored entry: a general complex operator
    x = np.random.default_rng(1).standard_normal(n) + 0.5j
    assert np.abs(op.perform_op(x) - A @ x).max() <= 1e-13 * np.abs(A @ x).max()
    g = sb.GenEigsSolver(op, 3, m)
    v0 = OH.simple_random_complex(5, n)
    g.init(v0)
    g.factorize_from(1, m // 2)
    g.factorize_from(m // 2, m)
    fz = g.factorization()
    V, H, f = fz["V"], fz["H"], fz["f"]
    E = A @ V - V @ H
    E[:, -1] -= f
    scale = max(1.0, np.abs(H).max())
    assert np.abs(E).max() <= 1e-12 * scale
    assert np.abs(V.conj().T @ V - np.eye(m)).max() <= 1e-12
    assert abs(np.linalg.norm(f) - fz["beta"]) <= 1e-12 * scale
    assert np.abs(np.tril(H, -2)).max() == 0.0 and np.abs(np.diag(H, -1).imag).max() == 0.0  # Hessenberg, real positive sub-diagonal
    ref = OH.arnoldi_factorize_complex(lambda v: A @ v, n, m, v0=v0)
    assert np.abs(H - ref["H"]).max() <= 1e-9 * scale and g.num_operations() == ref["nops"]


COMPLEX_FIXTURES = {10: (0.5, 3, 6), 100: (0.1, 10, 30), 1000: (0.01, 20, 50)}  # test/ComplexEigs.cpp:151-192


def complex_dense_kernels_case(sb, m):
    # test/QR.cpp:177-189 and test/Eigen.cpp with a complex Hessenberg matrix: Q unitary, Q^H H Q = RQ + sI, H Z = Z D to 1e-12
    rng = np.random.default_rng(100 + m)
    H = np.triu(rng.standard_normal((m, m)) + 1j * rng.standard_normal((m, m)), -1)
    mu = complex(*rng.standard_normal(2))
    D, Q = sb.dense.shifted_qr_z(H, mu)
    assert np.abs(Q.conj().T @ Q - np.eye(m)).max() <= 1e-12
    assert np.abs(Q.conj().T @ H @ Q - D).max() <= 1e-12 * m * max(1.0, np.abs(H).max())
    assert np.abs(np.tril(D, -2)).max() == 0.0
    D0, cs, sn = OH.hess_qr_complex(H, mu)
    assert np.abs(D - D0).max() <= 1e-12 * m * max(1.0, np.abs(H).max())
    ev, Z = sb.dense.hess_eigen_z(H)
    assert np.abs(H @ Z - Z * ev).max() <= 1e-12 * m * max(1.0, np.abs(H).max())
    assert np.abs(np.linalg.norm(Z, axis=0) - 1.0).max() <= 1e-12
    w = np.linalg.eigvals(H)
    assert max(np.abs(w - e).min() for e in ev) <= 1e-11 * max(1.0, np.abs(w).max())


def complex_gen_solver_case(sb, n, rule_name, check_history=True):
    # test/ComplexEigs.cpp:41-110 (sparse cases, maxit = 300): ||AU - UD||_inf <= 1e-9; eigenvalues within 1e-10 relative of the oracle
    prob, k, m = COMPLEX_FIXTURES[n]
    A = OH.gen_sparse_data_complex(n, prob)
    op = sb.SparseHermMatProd(A, uplo="general")
    g = sb.GenEigsSolver(op, k, m)
    g.init()
    nconv = g.compute(getattr(sb.SortRule, rule_name), 300)
    assert g.info() == sb.CompInfo.Successful and nconv == k
    ev, Z = g.eigenvalues(), g.eigenvectors()
    assert Z.dtype == np.complex128 and Z.shape == (n, k)
    assert np.abs(A @ Z - Z * ev).max() <= 1e-9
    ref = OH.gen_eigs_complex(A.tocsr().dot, n, k, m, getattr(O, rule_name), 300)
    assert ref.nconv == k
    assert np.abs(np.sort_complex(ev) - np.sort_complex(ref.eigenvalues)).max() <= 1e-10 * np.abs(ref.eigenvalues).max()
    if check_history and ref.niter <= 60:
        assert g.num_operations() == ref.nops and g.num_iterations() == ref.niter


def complex_gen_user_operator_case(sb, n=50):
    # a user-defined general complex operator (OpType concept with Scalar = std::complex<double>) behind GenEigsSolver
    rng = np.random.default_rng(9)
    A = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))

    class MyOp:
        def rows(self):
            return n

        def perform_op(self, x_in, y_out):
            y_out[:] = A @ x_in

    op = sb.UserOp(MyOp(), complex_scalar=True)
    g = sb.GenEigsSolver(op, 4, 16)
    g.init()
    assert g.compute(sb.SortRule.LargestMagn, 300) == 4 and g.info() == sb.CompInfo.Successful
    ev, Z = g.eigenvalues(), g.eigenvectors()
    assert np.abs(A @ Z - Z * ev).max() <= 1e-9
    ref = OH.gen_eigs_complex(lambda v: A @ v, n, 4, 16, O.LargestMagn, 300)
    assert g.num_operations() == ref.nops
    assert np.abs(np.sort_complex(ev) - np.sort_complex(ref.eigenvalues)).max() <= 1e-10 * np.abs(ref.eigenvalues).max()
    # argument checks of GenEigsBase.h:419-423 hold for complex operators too
    for nev, ncv in ((0, 6), (n - 1, n), (3, 4), (3, n + 1)):
        try:
            sb.GenEigsSolver(op, nev, ncv)
        except sb.InvalidArgument:
            continue
        raise AssertionError((nev, ncv))

"""GPU parity tests of the nonsymmetric (Arnoldi) path through the C ABI against the CPU oracle.
Tiers and tolerances follow test/GenEigs.cpp (solver, ||AU-UD||_inf <= 1e-9, maxit 300),
test/Arnoldi.cpp (factorisation, 1e-12), test/QR.cpp / test/Eigen.cpp / test/Schur.cpp (dense, 1e-12)."""
import numpy as np
import pytest
import scipy.sparse as sp

import oracle as O
from helpers import readme_banded, dense_as_csc

pytestmark = pytest.mark.gpu


def _rand_hess(rng, m):
    return np.triu(rng.standard_normal((m, m)), -1)


@pytest.mark.parametrize("m", [2, 3, 6, 20, 60, 64])
def test_hessenberg_qr_device(gpu, m):
    # test/QR.cpp:101-113
    rng = np.random.default_rng(m)
    H = _rand_hess(rng, m)
    for shift in (0.0, 0.41):
        D, Q = gpu.dense.shifted_qr(H, shift, "hess")
        R0, D0, Q0 = O.shifted_qr(H, shift, "hess")
        assert np.abs(Q.T @ Q - np.eye(m)).max() <= 1e-12
        assert np.abs(Q.T @ H @ Q - D).max() <= 1e-12 * m
        assert np.abs(Q - Q0).max() <= 1e-11 and np.abs(D - D0).max() <= 1e-11 * max(1, np.abs(D0).max())


@pytest.mark.parametrize("m", [3, 4, 6, 20, 60, 64])
def test_double_shift_qr_device(gpu, m):
    # test/QR.cpp:131-175
    rng = np.random.default_rng(10 + m)
    H = _rand_hess(rng, m)
    if m >= 20:
        H[7, 6] = 0.0  # split into blocks (DoubleShiftQR.h:351-385)
        H[12, 11] = 1e-300
    s, t = 0.4, 1.3
    D, Q = gpu.dense.double_shift_qr(H, s, t)
    D0, Q0 = O.double_shift_qr(H, s, t)
    assert np.abs(Q.T @ Q - np.eye(m)).max() <= 1e-12
    assert np.abs(Q.T @ H @ Q - D).max() <= 1e-12 * m
    R = Q.T @ (H @ H - s * H + t * np.eye(m))
    assert np.abs(np.tril(R, -1)).max() <= 1e-11 * m
    assert np.abs(Q - Q0).max() <= 1e-10 and np.abs(D - D0).max() <= 1e-10 * max(1, np.abs(D0).max())


@pytest.mark.parametrize("m", [1, 2, 3, 6, 20, 50, 60, 64])
def test_hessenberg_eigen_device(gpu, m):
    # test/Eigen.cpp:26-42 (||HU - UD|| <= 1e-12 scaled), eigenvalues vs the oracle
    rng = np.random.default_rng(20 + m)
    H = _rand_hess(rng, m)
    ev, V = gpu.dense.hess_eigen(H)
    ev0, V0 = O.hess_eigen(H)
    assert np.abs(H @ V - V * ev).max() <= 1e-12 * max(m, 4) * max(1, np.abs(H).max())
    assert np.abs(np.linalg.norm(V, axis=0) - 1).max() <= 1e-12
    assert np.abs(np.sort_complex(ev) - np.sort_complex(ev0)).max() <= 1e-10 * max(1, np.abs(ev0).max())
    # exact conjugate pairs / exact zero imaginary parts (relied upon by GenEigsBase.h:200-201)
    for z in ev:
        if z.imag != 0:
            assert np.any(ev == np.conj(z))


@pytest.mark.parametrize("n,m", [(10, 6), (100, 30), (1000, 50), (3000, 64)])
def test_arnoldi_factorization(gpu, n, m):
    # test/Arnoldi.cpp:19-85
    if n <= 1000:
        A = O.gen_sparse_data(n, {10: 0.5, 100: 0.1, 1000: 0.01}[n])
    else:
        A = sp.random(n, n, density=0.003, random_state=5, format="csc")
    op = gpu.SparseGenMatProd(A)
    eigs = gpu.GenEigsSolver(op, min(3, n - 2), m)
    v0 = O.simple_random(3, n)
    eigs.init(v0)
    eigs.factorize_from(1, m // 2)
    eigs.factorize_from(m // 2, m)
    fz = eigs.factorization()
    V, H, f = fz["V"], fz["H"], fz["f"]
    E = A @ V - V @ H
    E[:, -1] -= f
    scale = max(1.0, np.abs(H).max())
    assert np.abs(E).max() <= 1e-12 * scale
    assert np.abs(V.T @ V - np.eye(m)).max() <= 1e-12
    assert np.abs(np.tril(H, -2)).max() == 0.0
    ref = O.factorize(O.Csr.from_scipy(A), m, v0=v0, kind="arnoldi")
    assert np.abs(H - ref["H"]).max() <= 1e-9 * scale


GEN_CASES = [(10, 0.5, 3, 6), (100, 0.1, 10, 30), (1000, 0.01, 20, 50)]
GEN_RULES = [("LargestMagn", False), ("LargestReal", False), ("LargestImag", False), ("SmallestMagn", True), ("SmallestReal", False), ("SmallestImag", True)]


@pytest.mark.parametrize("n,prob,k,m", GEN_CASES)
@pytest.mark.parametrize("rule,allow_fail", GEN_RULES)
def test_gen_eigs_sparse_reference_cases(gpu, n, prob, k, m, rule, allow_fail):
    # test/GenEigs.cpp:38-107,143-174
    A = O.gen_sparse_data(n, prob)
    eigs = gpu.GenEigsSolver(gpu.SparseGenMatProd(A), k, m)
    eigs.init()
    nconv = eigs.compute(gpu.SortRule[rule], 300)
    ref = O.gen_eigs(O.Csr.from_scipy(A), k, m, int(gpu.SortRule[rule]), 300)
    if allow_fail and eigs.info() != gpu.CompInfo.Successful:
        # test/GenEigs.cpp:47-54: on these selection rules the reference's own test only warns when the solver does not converge within
        # maxit = 300 (whether a run gets there is decided by rounding-level differences in the restart history).  What must still hold:
        # the documented status, the iteration accounting and the pairs that did converge.
        # (pairs flagged as converged by the last convergence test are returned as they stand after the final restart,
        # GenEigsBase.h:514-521, so no residual bound applies to them)
        assert eigs.info() == gpu.CompInfo.NotConverging and eigs.num_iterations() == 301 and nconv < k
        assert len(eigs.eigenvalues()) == nconv
        return
    assert eigs.info() == gpu.CompInfo.Successful and nconv == k
    evals, evecs = eigs.eigenvalues(), eigs.eigenvectors()
    assert np.abs(A @ evecs - evecs * evals).max() <= 1e-9
    assert ref.info == O.Successful
    a = np.sort_complex(np.round(evals, 9))
    b = np.sort_complex(np.round(ref.eigenvalues, 9))
    assert np.abs(np.sort_complex(evals) - np.sort_complex(ref.eigenvalues)).max() <= 1e-9 * np.abs(ref.eigenvalues).max() or np.allclose(a, b)


def test_readme_banded_nonsymmetric(gpu):
    # README.md:146-178
    M = readme_banded(10)
    eigs = gpu.GenEigsSolver(gpu.SparseGenMatProd(sp.csc_matrix(M)), 3, 6)
    eigs.init()
    eigs.compute(gpu.SortRule.LargestMagn)
    assert eigs.info() == gpu.CompInfo.Successful
    true = sorted(np.linalg.eigvals(M), key=lambda z: -abs(z))[:3]
    ev = eigs.eigenvalues()
    assert np.allclose(ev.real, np.real(true), atol=1e-9) and np.abs(ev.imag).max() < 1e-9


def test_gen_argument_checks(gpu):
    op = gpu.SparseGenMatProd(sp.identity(10, format="csc"))
    for nev, ncv in [(0, 5), (9, 12), (3, 4), (3, 11)]:
        with pytest.raises(gpu.InvalidArgument):  # GenEigsBase.h:419-423
            gpu.GenEigsSolver(op, nev, ncv)
    eigs = gpu.GenEigsSolver(op, 3, 6)
    with pytest.raises(gpu.InvalidArgument):
        eigs.init(np.zeros(10))
    eigs.init()
    with pytest.raises(gpu.InvalidArgument):
        eigs.compute(gpu.SortRule.LargestAlge)


def test_gen_eigs_medium_vs_oracle(gpu):
    # BASELINE config C3 shape (G_gen, k = 10, ncv = 30) at a size the oracle finishes in seconds.  A purely random
    # nonsymmetric matrix has its dominant eigenvalues packed on the rim of the circular-law disk (the oracle itself
    # needs > 1000 restarts at this size), so 20 separated diagonal entries are planted; two planted spacings give one
    # fast and one slower-converging case.
    from spectra_b200 import synth

    n = 50000
    rp, ci, v = synth.csr(n, 20, 1, False)
    A0 = sp.csr_matrix((v, ci, rp), shape=(n, n))
    for base, gap in ((3.0, 0.35), (2.0, 0.08)):
        d = np.zeros(n)
        d[:20] = base + gap * np.arange(20)
        A = (A0 + sp.diags(d)).tocsr()
        A.sort_indices()
        op = gpu.SparseGenMatProd(A)
        eigs = gpu.GenEigsSolver(op, 10, 30)
        eigs.init()
        nconv = eigs.compute(gpu.SortRule.LargestMagn)
        assert eigs.info() == gpu.CompInfo.Successful and nconv == 10
        evals, X = eigs.eigenvalues(), eigs.eigenvectors()
        res = np.linalg.norm(A @ X - X * evals, axis=0) / np.abs(evals)
        assert res.max() <= 1e-9
        ref = O.gen_eigs(O.Csr.from_scipy(A), 10, 30, O.LargestMagn, want_vectors=False)
        assert ref.info == O.Successful
        assert np.abs(np.sort_complex(evals) - np.sort_complex(ref.eigenvalues)).max() <= 1e-9 * np.abs(ref.eigenvalues).max()
        assert abs(eigs.num_operations() - ref.nops) <= max(60, ref.nops // 5)


def test_gen_eigs_full_size_properties(gpu):
    # BASELINE config C3: nonsymmetric CSR n = 1e6, nnz/row = 20, k = 10, ncv = 30, LargestMagn (planted separated eigenvalues, see above).
    # Size-independent properties: relative residuals, conjugate-closed spectrum, sortedness, run-to-run reproducibility.
    from spectra_b200 import synth

    n = 1_000_000
    rp, ci, v = synth.csr(n, 20, 1, False)
    d = np.zeros(n)
    d[:20] = 3.0 + 0.35 * np.arange(20)
    A = (sp.csr_matrix((v, ci, rp), shape=(n, n)) + sp.diags(d)).tocsr()
    A.sort_indices()
    op = gpu.SparseGenMatProd(A)
    runs = []
    for _ in range(2):
        eigs = gpu.GenEigsSolver(op, 10, 30)
        eigs.init()
        nconv = eigs.compute(gpu.SortRule.LargestMagn)
        assert eigs.info() == gpu.CompInfo.Successful and nconv == 10
        runs.append((eigs.eigenvalues(), eigs.num_operations(), eigs.num_iterations()))
    evals, X = eigs.eigenvalues(), eigs.eigenvectors()
    assert X.shape == (n, 10)
    res = np.linalg.norm(A @ X - X * evals, axis=0) / np.abs(evals)
    assert res.max() <= 1e-10
    assert np.all(np.diff(np.abs(evals)) <= 1e-12 * np.abs(evals).max())  # LargestMagn ordering (GenEigsBase.h:501-502 default sorting)
    assert np.abs(np.linalg.norm(X, axis=0) - 1).max() <= 1e-12
    # the planted eigenvalues are real up to the perturbation: the top 10 lie near 3 + 0.35 j, j = 19..10
    assert np.abs(np.sort(evals.real)[::-1] - (3.0 + 0.35 * np.arange(19, 9, -1))).max() < 0.5
    assert np.array_equal(runs[0][0], runs[1][0]) and runs[0][1:] == runs[1][1:]


def test_gen_eigs_c3_unplanted_history(gpu):
    # BASELINE config C3 as stated: G_gen(n = 1e6, 20 nnz/row, seed 1), k = 10, ncv = 30, LargestMagn -- NO planted eigenvalues.  The
    # spectrum fills a disc (circular law) whose rim is packed with eigenvalues of nearly equal modulus, so neither the reference nor
    # this solver converges within the default 1000 restarts; what can be pinned is (1) the Arnoldi factorisation itself against the
    # oracle's at full size and (2) the maxit-bounded run: status NotConverging and the oracle's operation / iteration counts
    # (tests/golden/baseline_C3.json, maxit = 40).
    import json
    import os

    from spectra_b200 import synth

    n, k, m = 1_000_000, 10, 30
    rp, ci, v = synth.csr(n, 20, 1, False)
    A = sp.csr_matrix((v, ci, rp), shape=(n, n))
    op = gpu.SparseGenMatProd.from_csr_slab(n, 0, rp, ci, v)
    g = gpu.GenEigsSolver(op, k, m)
    g.init()
    g.factorize_from(1, m)
    fz = g.factorization()
    V, H, f = fz["V"], fz["H"], fz["f"]
    E = A @ V - V @ H
    E[:, -1] -= f
    hs = np.abs(H).max()
    assert np.abs(E).max() <= 1e-12 * max(1.0, hs)                     # A V = V H + f e'   (test/Arnoldi.cpp:70-76)
    assert np.abs(V.T @ V - np.eye(m)).max() <= 1e-12                  # V'V = I
    assert np.abs(V.T @ f).max() <= 1e-12 * max(1.0, np.linalg.norm(f))
    ref = O.factorize(O.Csr.adopt(n, rp, ci, v), m, kind="arnoldi")
    assert np.abs(H - ref["H"]).max() <= 1e-9 * hs                      # same Hessenberg matrix as the CPU oracle from the same start
    assert abs(fz["beta"] - ref["beta"]) <= 1e-9 * max(1.0, ref["beta"])
    del V, E, fz
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "baseline_C3.json")
    with open(path) as fh:
        gold = json.load(fh)
    g2 = gpu.GenEigsSolver(op, k, m)
    g2.init()
    nconv = g2.compute(gpu.SortRule.LargestMagn, gold["maxit"])
    assert g2.info() == gpu.CompInfo.NotConverging and gold["info"] == O.NotConverging
    assert nconv == gold["nconv"] and len(g2.eigenvalues()) == nconv
    assert g2.num_iterations() == gold["niter"]
    assert abs(g2.num_operations() - gold["nops"]) <= 20


@pytest.mark.parametrize("n,k,m", [(100, 10, 30), (1000, 20, 50)])
def test_gen_eigs_against_golden_spectra(gpu, n, k, m):
    import golden_cases as GC

    A = O.gen_sparse_data(n, {100: 0.1, 1000: 0.01}[n])
    g = gpu.GenEigsSolver(gpu.SparseGenMatProd(A), k, m)
    g.init()
    assert g.compute(gpu.SortRule.LargestMagn) == k
    GC.check_gen_values(n, g.eigenvalues(), k)

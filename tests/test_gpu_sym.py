"""GPU parity tests of the symmetric path, mirroring the three tiers of the reference's tests
(operator / factorisation / solver) through the C ABI, checked against the CPU oracle on the same
inputs.  Tolerances: operator 1e-13 relative (summation order differs), factorisation and dense
kernels 1e-12 (test/Arnoldi.cpp, test/QR.cpp, test/Eigen.cpp), solver ||AU-UD||_inf <= 1e-9
(test/SymEigs.cpp:60-64) and eigenvalues within 1e-10 relative of the oracle (north star)."""
import numpy as np
import pytest
import scipy.sparse as sp

import oracle as O
from helpers import EXAMPLE2, cycle_laplacian, dense_as_csc, sym_full

pytestmark = pytest.mark.gpu


# ---------------------------------------------------------------- operator tier
@pytest.mark.parametrize("fmt", ["csc", "csr"])
@pytest.mark.parametrize("uplo", ["lower", "upper"])
def test_sparse_sym_mat_prod(gpu, fmt, uplo):
    # test/SparseSymMatProd.cpp:37-54
    rng = np.random.default_rng(0)
    A = O.gen_sparse_data(100, 0.1)
    A = A.tocsc() if fmt == "csc" else A.tocsr()
    op = gpu.SparseSymMatProd(A, uplo=uplo)
    ref = O.Csr.from_scipy(A, uplo)
    assert op.rows() == 100 and op.cols() == 100
    for _ in range(3):
        x = rng.standard_normal(100)
        y, y0 = op.perform_op(x), ref.spmv(x)
        assert np.abs(y - y0).max() <= 1e-13 * max(1.0, np.abs(y0).max())
    M = rng.standard_normal((100, 7))
    Y0 = ref.to_scipy() @ M
    assert np.abs(op @ M - Y0).max() <= 1e-13 * np.abs(Y0).max()
    assert op(45, 22) == A[45, 22]  # the STORED coefficient (SparseSymMatProd.h:101-104)


@pytest.mark.parametrize("fmt", ["csc", "csr"])
def test_sparse_gen_mat_prod(gpu, fmt):
    # test/SparseGenMatProd.cpp:37-53
    rng = np.random.default_rng(1)
    A = O.gen_sparse_data(100, 0.1)
    A = A.tocsc() if fmt == "csc" else A.tocsr()
    op = gpu.SparseGenMatProd(A)
    x = rng.standard_normal(100)
    y0 = A @ x
    assert np.abs(op.perform_op(x) - y0).max() <= 1e-13 * np.abs(y0).max()
    assert op(45, 22) == A[45, 22]


def test_operator_edge_cases(gpu):
    # empty matrix, empty rows, a long row, int64 outer index
    Z = gpu.SparseSymMatProd(sp.csc_matrix((50, 50)))
    assert np.array_equal(Z.perform_op(np.ones(50)), np.zeros(50))
    rng = np.random.default_rng(2)
    n = 3000
    A = sp.random(n, n, density=0.002, random_state=3, format="csr")
    A = A.tolil()
    A[7, :] = rng.standard_normal(n)  # dense row
    A[11, :] = 0
    A = A.tocsr()
    x = rng.standard_normal(n)
    op = gpu.SparseGenMatProd((n, A.indptr.astype(np.int64), A.indices, A.data, "row"))
    y0 = A @ x
    assert np.abs(op.perform_op(x) - y0).max() <= 1e-12 * np.abs(y0).max()


def test_spmv_large_synthetic(gpu):
    from spectra_b200 import synth

    n = 200000
    rp, ci, v = synth.csr(n, 20, 0, True)
    A = sp.csr_matrix((v, ci, rp), shape=(n, n))
    op = gpu.SparseGenMatProd.from_csr_slab(n, 0, rp, ci, v)
    x = O.simple_random(7, n)
    y0 = A @ x
    assert np.abs(op.perform_op(x) - y0).max() <= 1e-13 * np.abs(y0).max()
    # the symmetric wrapper (lower triangle, CSC storage == CSR arrays of a symmetric matrix) gives the same operator
    op2 = gpu.SparseSymMatProd((n, rp, ci, v, "col"))
    assert op2.nnz == op.nnz
    assert np.abs(op2.perform_op(x) - y0).max() <= 1e-13 * np.abs(y0).max()


# ---------------------------------------------------------------- dense-kernel tier
def _rand_tridiag(rng, m):
    d, e = rng.standard_normal(m), rng.standard_normal(m - 1)
    return np.diag(d) + np.diag(e, 1) + np.diag(e, -1)


@pytest.mark.parametrize("m", [2, 3, 6, 20, 50, 60, 64, 100])
def test_tridiag_eigen_device(gpu, m):
    # test/Eigen.cpp:68-86 tolerance 1e-12; bitwise-close to the oracle
    rng = np.random.default_rng(m)
    T = _rand_tridiag(rng, m)
    ev, Z = gpu.dense.tridiag_eigen(T)
    ev0, Z0 = O.tridiag_eigen(T)
    assert np.abs(T @ Z - Z * ev).max() <= 1e-12 * m
    assert np.abs(Z.T @ Z - np.eye(m)).max() <= 1e-12 * m
    assert np.abs(ev - ev0).max() <= 1e-12 * max(1, np.abs(ev0).max())
    # zero matrix early exit (TridiagEigen.h:142-150)
    ev, Z = gpu.dense.tridiag_eigen(np.zeros((m, m)))
    assert np.array_equal(ev, np.zeros(m)) and np.array_equal(Z, np.eye(m))


@pytest.mark.parametrize("impl", [0, 1], ids=["dmma_tma", "fma"])
@pytest.mark.parametrize("n,m,kk", [(1, 2, 2), (63, 6, 4), (64, 20, 11), (65, 30, 21), (1000, 60, 41), (4099, 64, 64), (300_001, 60, 37), (5000, 50, 1)])
def test_restart_gemm_compress(gpu, impl, n, m, kk):
    # Arnoldi::compress_V (Arnoldi.h:320-340): V[:, :kk] <- V Q[:, :kk]; f <- f Q(m-1,kk-2) + Vnew[:,kk-1] H(kk-1,kk-2); beta = ||f||
    rng = np.random.default_rng(n + m + kk)
    V = rng.standard_normal((n, m))
    Q = np.linalg.qr(rng.standard_normal((m, m)))[0]
    if n % 2:
        Q = np.triu(Q, -(m - kk + 1))  # the band shape of a restart's Q (Arnoldi.h:330): exact zeros below
    H = rng.standard_normal((m, m))
    f = rng.standard_normal(n)
    ref = V @ Q[:, :kk]
    scale = np.abs(V).max() * m
    out = gpu.dense.compress(V, Q, kk, impl=impl)
    assert np.abs(out - ref).max() <= 4e-16 * scale * 4
    if kk >= 2:
        out2, f2, nrm2 = gpu.dense.compress(V, Q, kk, f=f, H=H, impl=impl)
        assert np.array_equal(out2, out)  # deterministic
        fref = f * Q[m - 1, kk - 2] + ref[:, kk - 1] * H[kk - 1, kk - 2]
        assert np.abs(f2 - fref).max() <= 1e-14 * max(1.0, np.abs(fref).max()) * m
        assert abs(nrm2 - fref @ fref) <= 1e-13 * (fref @ fref)


@pytest.mark.parametrize("variant", [0, 1, 2])
def test_givens_rotation_device(gpu, variant):
    # Givens<double>::compute_rotation on the device (Givens.h:166-205, StableScaling :28-86), the reference's own test
    # (test/Givens.cpp:64-99: c x - s y = r and s x + c y = 0 to 1e-12 on 100000 draws from U(-100, 100) with 10 % exact zeros), plus the
    # branches that test never reaches: the Taylor branch (ratio < 0.1 eps^(1/4) = 1.22e-5), both orderings |x| >< |y|, every sign
    # combination, extreme magnitudes.  variant 0 = the reference's formulas, 1 = the rsqrt form the QR kernels call, 2 = makeGivens.
    rng = np.random.default_rng(0)
    nsim = 100_000
    x = np.where(rng.random(nsim) < 0.1, 0.0, rng.uniform(-100, 100, nsim))
    y = np.where(rng.random(nsim) < 0.1, 0.0, rng.uniform(-100, 100, nsim))
    # Taylor branch: ratios from 1e-5 down to 1e-17, both orderings, all signs
    big = rng.uniform(0.5, 2.0, 4000) * rng.choice([-1.0, 1.0], 4000)
    small = big * 10.0 ** rng.uniform(-17, -4.9, 4000) * rng.choice([-1.0, 1.0], 4000)
    x = np.concatenate([x, big, small, [3.0, -3.0, 0.0, 0.0, 0.0, 1e200, 1e-200, -1e150, 5e-324, 1.0]])
    y = np.concatenate([y, small, big, [0.0, 0.0, 4.0, -4.0, 0.0, 1e200, 1e-200, 1e150, 1.0, 5e-324]])
    r, c, s = gpu.dense.givens(x, y, variant)
    scale = np.maximum(1.0, np.hypot(x, y))
    assert np.abs(c * x - s * y - r)[:nsim].max() <= 1e-12 and np.abs((c * x - s * y - r) / scale).max() <= 4e-16 * 4
    assert np.abs((s * x + c * y) / scale).max() <= 4e-16 * 4 and np.abs(s * x + c * y)[:nsim].max() <= 1e-12
    assert np.all(r >= 0.0) and np.abs(c * c + s * s - 1.0).max() <= 1e-15 * 4
    # sign convention of the special cases (Givens.h:176-192; makeGivens: c = sign(p), s = 0 / c = 0, s = -sign(q))
    k = len(x) - 10
    assert (c[k], s[k], r[k]) == (1.0, 0.0, 3.0) and (c[k + 1], s[k + 1], r[k + 1]) == (-1.0, 0.0, 3.0)
    assert (c[k + 2], s[k + 2], r[k + 2]) == (0.0, -1.0, 4.0) and (c[k + 3], s[k + 3], r[k + 3]) == (0.0, 1.0, 4.0)
    assert (c[k + 4], s[k + 4], r[k + 4]) == (1.0, 0.0, 0.0)
    # agreement with the CPU oracle's restatement of Givens.h (variants 0 and 1 implement it, to a few ulp)
    if variant < 2:
        ref = np.array([O.givens(a, b) for a, b in zip(x[nsim:], y[nsim:])])
        ref_main = np.array([O.givens(a, b) for a, b in zip(x[:2000], y[:2000])])
        tol = 4.5e-16 if variant == 0 else 1e-15  # device hypot / rsqrt are within 1 ulp of the correctly rounded host functions
        for got, want in ((r[nsim:], ref[:, 0]), (c[nsim:], ref[:, 1]), (s[nsim:], ref[:, 2]), (r[:2000], ref_main[:, 0]), (c[:2000], ref_main[:, 1]),
                          (s[:2000], ref_main[:, 2])):
            assert np.all(np.abs(got - want) <= tol * np.maximum(np.abs(want), 1e-300)), variant


@pytest.mark.parametrize("m", [2, 3, 6, 20, 60, 64])
def test_tridiag_qr_device(gpu, m):
    # test/QR.cpp:115-129: Q orthogonal, Q'(T - sI) upper triangular, Q'TQ = D (1e-12, scaled)
    rng = np.random.default_rng(100 + m)
    T = _rand_tridiag(rng, m)
    lam = float(np.linalg.eigvalsh(T)[0])
    tn = max(1.0, np.abs(T).max())
    for shift in (0.0, 0.37, lam):
        D, Q = gpu.dense.shifted_qr(T, shift, "tridiag")
        R0, D0, Q0 = O.shifted_qr(T, shift, "tridiag")
        err = dict(
            orth=np.abs(Q.T @ Q - np.eye(m)).max(),
            lowerR=np.abs(np.tril(Q.T @ (T - shift * np.eye(m)), -1)).max() / tn,
            band=np.abs(np.tril(D, -2)).max() + np.abs(D - D.T).max(),
            spectrum=np.abs(np.linalg.eigvalsh(D) - np.linalg.eigvalsh(T)).max() / tn,
            sim=np.abs(Q.T @ T @ Q - D).max() / tn,
            q_vs_oracle=np.abs(Q[:, :m - 1] - Q0[:, :m - 1]).max(),
            d_vs_oracle=np.abs(D - D0).max() / tn,
        )
        print(m, shift, err)
        assert err["orth"] <= 1e-12 * m and err["lowerR"] <= 1e-12 * m and err["band"] == 0.0 and err["spectrum"] <= 1e-12 * m, (m, shift, err)
        if shift != lam:
            # (with an exact-eigenvalue shift the last rotation is decided by rounding noise, R[m-1,m-1] ~ 0, so Q'TQ = D
            #  and the entry-wise agreement with the oracle hold only for the well-determined part)
            assert err["sim"] <= 1e-12 * m, (m, shift, err)
            assert err["q_vs_oracle"] <= 1e-9 and err["d_vs_oracle"] <= 1e-9, (m, shift, err)


@pytest.mark.parametrize("rule", [O.LargestMagn, O.LargestAlge, O.SmallestMagn, O.SmallestAlge, O.BothEnds])
def test_sym_restart_step_device(gpu, rule):
    # one HermEigsBase restart-prepare step vs the oracle on a Lanczos H of the n=1000 fixture
    A = O.gen_sparse_data(1000, 0.01)
    fz = O.factorize(O.Csr.from_scipy(A, "lower"), 50)
    H, beta = fz["H"], fz["beta"]
    a = gpu.dense.sym_restart(H, beta, 20, rule, 1e-10)
    b = O.sym_restart_prepare(H, beta, 20, rule, 1e-10)
    assert a["nconv"] == b["nconv"] and a["k"] == b["k"], (a["nconv"], b["nconv"], a["k"], b["k"])
    hn = np.abs(H).max()
    k = a["k"]
    Q, Hn = a["Q"], a["H"]
    err = dict(
        ritz_val=np.abs(a["ritz_val"] - b["ritz_val"]).max() / hn,
        ritz_est=np.abs(np.abs(a["ritz_est"]) - np.abs(b["ritz_est"])).max(),
        orth=np.abs(Q.T @ Q - np.eye(50)).max(),
        band=np.abs(np.tril(Hn, -2)).max() + np.abs(Hn - Hn.T).max(),
        spectrum=np.abs(np.linalg.eigvalsh(Hn) - np.linalg.eigvalsh(H)).max() / hn,
        sim_lead=np.abs(Q.T @ H @ Q - Hn)[:k + 1, :k + 1].max() / hn,
    )
    # With exact-eigenvalue shifts the individual columns of Q are not forward stable (both implementations are only
    # backward stable), but the restart is defined by what the leading block preserves: span(Q[:, :k]) is the invariant
    # subspace of the wanted Ritz vectors, eig(Hn[:k,:k]) are the wanted Ritz values, and Hn(k, k-1) ~ 0.
    w, S = np.linalg.eigh(H)

    def lead_metrics(Qx, Hx):
        wanted = b["ritz_val"][:k]
        Sw = S[:, [int(np.argmin(np.abs(w - x))) for x in wanted]]
        Qk = Qx[:, :k]
        return dict(subspace=np.linalg.norm(Sw - Qk @ (Qk.T @ Sw), axis=0).max(),
                    lead_eigs=np.abs(np.sort(np.linalg.eigvalsh(Hx[:k, :k])) - np.sort(wanted)).max() / hn,
                    coupling=abs(Hx[k, k - 1]) / hn)

    mine, ref = lead_metrics(Q, Hn), lead_metrics(b["Q"], b["H"])
    print(rule, err, mine, ref)
    assert np.array_equal(a["conv"], b["conv"])
    assert err["ritz_val"] <= 1e-12 and err["ritz_est"] <= 1e-10, err
    assert err["orth"] <= 1e-12 and err["band"] == 0.0 and err["spectrum"] <= 1e-11 and err["sim_lead"] <= 1e-10, err
    assert mine["subspace"] <= 1e-7 and mine["lead_eigs"] <= 1e-12 and mine["coupling"] <= 1e-7, (mine, ref)


# ---------------------------------------------------------------- factorisation tier
@pytest.mark.parametrize("n,m", [(10, 6), (100, 20), (1000, 50), (5000, 64)])
def test_lanczos_factorization(gpu, n, m):
    # test/Arnoldi.cpp:19-85: init, factorize_from(1, m/2), factorize_from(m/2, m); AV - VH = f e', V'V = I (1e-12)
    if n <= 1000:
        A = O.gen_sparse_data(n, {10: 0.5, 100: 0.1, 1000: 0.01}[n])
    else:
        A = sp.random(n, n, density=0.002, random_state=5, format="csc")
    Af = sym_full(A)
    op = gpu.SparseSymMatProd(A)
    eigs = gpu.SymEigsSolver(op, min(3, n - 1), m)
    v0 = O.simple_random(3, n)
    eigs.init(v0)
    eigs.factorize_from(1, m // 2)
    eigs.factorize_from(m // 2, m)
    fz = eigs.factorization()
    V, H, f = fz["V"], fz["H"], fz["f"]
    E = Af @ V - V @ H
    E[:, -1] -= f
    scale = max(1.0, np.abs(H).max())
    assert np.abs(E).max() <= 1e-12 * scale
    assert np.abs(V.T @ V - np.eye(m)).max() <= 1e-12
    assert abs(np.linalg.norm(f) - fz["beta"]) <= 1e-12 * scale
    assert np.abs(H - H.T).max() == 0.0 and np.abs(np.triu(H, 2)).max() == 0.0
    # step-for-step agreement with the oracle's H
    ref = O.factorize(O.Csr.from_scipy(A, "lower"), m, v0=v0)
    assert np.abs(H - ref["H"]).max() <= 1e-9 * scale
    with pytest.raises(gpu.InvalidArgument):
        eigs2 = gpu.SymEigsSolver(op, min(3, n - 1), m)
        eigs2.init(v0)
        eigs2.factorize_from(5, m)  # from_k > current dimension (Lanczos.h:70-75)


# ---------------------------------------------------------------- solver tier
SYM_CASES = [(10, 0.5, 3, 6), (100, 0.1, 10, 20), (1000, 0.01, 20, 50)]
RULES = ["LargestMagn", "LargestAlge", "SmallestMagn", "SmallestAlge", "BothEnds"]


@pytest.mark.parametrize("n,prob,k,m", SYM_CASES)
@pytest.mark.parametrize("rule", RULES)
def test_sym_eigs_sparse_reference_cases(gpu, n, prob, k, m, rule):
    # test/SymEigs.cpp:44-65,133-167
    if n == 1000 and rule == "SmallestMagn":
        pytest.skip("~23k matvecs; interior selection is covered at n <= 100")
    A = O.gen_sparse_data(n, prob)
    Af = sym_full(A)
    op = gpu.SparseSymMatProd(A)
    eigs = gpu.SymEigsSolver(op, k, m)
    eigs.init()
    nconv = eigs.compute(gpu.SortRule[rule])
    assert eigs.info() == gpu.CompInfo.Successful and nconv == k
    evals, evecs = eigs.eigenvalues(), eigs.eigenvectors()
    assert np.abs(Af @ evecs - evecs * evals).max() <= 1e-9
    ref = O.sym_eigs(O.Csr.from_scipy(A, "lower"), k, m, int(gpu.SortRule[rule]))
    assert ref.info == O.Successful
    assert np.abs(evals - ref.eigenvalues).max() <= 1e-10 * np.abs(ref.eigenvalues).max()
    # the restart sequence is the same algorithm: iteration / operation counts agree (allow rounding-induced drift)
    assert abs(eigs.num_operations() - ref.nops) <= max(3 * m, ref.nops // 5)
    st = eigs.stats()
    assert st["kernel_launches"] > 0 and st["spmv_launches"] == eigs.num_operations()


def test_readme_diag_kat(gpu):
    # SymEigsSolver.h:99-126 with the operator given as a sparse diagonal matrix
    A = sp.diags(np.arange(1.0, 11.0)).tocsc()
    eigs = gpu.SymEigsSolver(gpu.SparseSymMatProd(A), 3, 6)
    eigs.init()
    eigs.compute(gpu.SortRule.LargestAlge)
    assert eigs.info() == gpu.CompInfo.Successful
    assert np.allclose(eigs.eigenvalues(), [10, 9, 8], atol=1e-10)
    assert eigs.num_iterations() > 0 and eigs.num_operations() > 0


@pytest.mark.parametrize("k,m", [(3, 6), (5, 12), (6, 12)])
def test_example1_cycle_laplacian(gpu, k, m):
    # test/Example1.cpp (issue #144): repeated eigenvalues exercise expand_basis; tol = 1e-15
    M = cycle_laplacian(20)
    true = np.linalg.eigvalsh(M)
    eigs = gpu.SymEigsSolver(gpu.SparseSymMatProd(dense_as_csc(M)), k, m)
    eigs.init()
    eigs.compute(gpu.SortRule.LargestMagn, 1000, 1e-15, gpu.SortRule.SmallestAlge)
    assert eigs.info() == gpu.CompInfo.Successful
    evals, U = eigs.eigenvalues(), eigs.eigenvectors()
    assert np.abs(M @ U - U * evals).max() <= 1e-9
    assert np.abs(true[-k:] - evals).max() <= 1e-9


@pytest.mark.parametrize("idx", [0, 1, 2])
def test_example2_near_rank_one(gpu, idx):
    # test/Example2.cpp (issue #159): beta < sqrt(eps) restart heuristic, n = 5, nev = 1, ncv = 3
    M = EXAMPLE2[idx]
    eigs = gpu.SymEigsSolver(gpu.SparseSymMatProd(dense_as_csc(M)), 1, 3)
    eigs.init()
    eigs.compute(gpu.SortRule.LargestMagn)
    assert eigs.info() == gpu.CompInfo.Successful
    evals, U = eigs.eigenvalues(), eigs.eigenvectors()
    assert np.abs(M @ U - U * evals).max() <= 1e-8
    assert abs(np.linalg.eigvalsh(M)[-1] - evals[0]) <= 1e-8


def test_example4_zero_matrix_and_null_init(gpu):
    # test/Example4.cpp:59-92
    n = 100
    rng = np.random.default_rng(123)
    v0 = rng.uniform(-1, 1, n)
    eigs = gpu.SymEigsSolver(gpu.SparseSymMatProd(sp.csc_matrix((n, n))), 3, 6)
    eigs.init(v0)
    eigs.compute(gpu.SortRule.LargestAlge)
    assert eigs.info() == gpu.CompInfo.Successful and np.abs(eigs.eigenvalues()).max() <= 1e-8
    U = rng.uniform(-1, 1, (n, n))
    w, Q = np.linalg.eigh(U + U.T)
    w[-1] = 0.0
    A = (Q * w) @ Q.T
    A = (A + A.T) / 2
    eigs = gpu.SymEigsSolver(gpu.SparseSymMatProd(dense_as_csc(A)), 3, 6)
    eigs.init(Q[:, -1].copy())
    eigs.compute(gpu.SortRule.LargestAlge)
    assert eigs.info() == gpu.CompInfo.Successful
    true = np.sort(np.linalg.eigvalsh(A))
    evals, Uv = eigs.eigenvalues(), eigs.eigenvectors()
    assert np.abs(true[-3:][::-1] - evals).max() <= 1e-8
    assert np.abs(A @ Uv - Uv * evals).max() <= 1e-8


def test_argument_and_error_behaviour(gpu):
    op = gpu.SparseSymMatProd(sp.identity(10, format="csc"))
    for nev, ncv in [(0, 5), (10, 12), (3, 3), (3, 11)]:
        with pytest.raises(gpu.InvalidArgument):  # HermEigsBase.h:267-271
            gpu.SymEigsSolver(op, nev, ncv)
    eigs = gpu.SymEigsSolver(op, 3, 6)
    with pytest.raises(gpu.InvalidArgument):  # Arnoldi.h:147-148
        eigs.init(np.zeros(10))
    with pytest.raises(gpu.LogicError):
        gpu.SymEigsSolver(op, 3, 6).compute()
    eigs.init()
    with pytest.raises(gpu.InvalidArgument):  # HermEigsBase.h:231-233
        eigs.compute(gpu.SortRule.LargestAlge, 1000, 1e-10, gpu.SortRule.BothEnds)
    eigs = gpu.SymEigsSolver(op, 3, 6)
    eigs.init()
    with pytest.raises(gpu.InvalidArgument):  # SelectionRule.h:261-262
        eigs.compute(gpu.SortRule.LargestReal)
    # not converging: maxit = 1 on a hard problem -> NotConverging, min(nev, nconv) returned (HermEigsBase.h:387-389)
    A = O.gen_sparse_data(1000, 0.01)
    e2 = gpu.SymEigsSolver(gpu.SparseSymMatProd(A), 20, 50)
    e2.init()
    nconv = e2.compute(gpu.SortRule.SmallestMagn, 1)
    ref = O.sym_eigs(O.Csr.from_scipy(A, "lower"), 20, 50, O.SmallestMagn, 1)
    assert e2.info() == gpu.CompInfo.NotConverging and nconv == ref.nconv
    assert e2.num_iterations() == ref.niter and e2.num_operations() == ref.nops
    assert len(e2.eigenvalues()) == nconv


def test_sym_eigs_medium_vs_oracle_and_arpack(gpu):
    # synthetic G_sym at a size the oracle finishes in seconds
    from scipy.sparse.linalg import eigsh
    from spectra_b200 import synth

    n = 50000
    rp, ci, v = synth.csr(n, 20, 0, True)
    A = sp.csr_matrix((v, ci, rp), shape=(n, n))
    op = gpu.SparseSymMatProd((n, rp, ci, v, "col"))
    eigs = gpu.SymEigsSolver(op, 20, 60)
    eigs.init()
    nconv = eigs.compute(gpu.SortRule.LargestAlge)
    assert eigs.info() == gpu.CompInfo.Successful and nconv == 20
    evals, X = eigs.eigenvalues(), eigs.eigenvectors()
    res = np.linalg.norm(A @ X - X * evals, axis=0) / np.abs(evals)
    assert res.max() <= 1e-10
    ref = O.sym_eigs(O.Csr.adopt(n, rp, ci, v), 20, 60, O.LargestAlge, want_vectors=False)
    assert np.abs(evals - ref.eigenvalues).max() <= 1e-10 * np.abs(ref.eigenvalues).max()
    w = eigsh(A, k=20, which="LA", ncv=60, tol=1e-12, return_eigenvectors=False)
    assert np.abs(np.sort(evals) - np.sort(w)).max() <= 1e-10 * np.abs(w).max()


def baseline_golden(name):
    import json
    import os

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", f"baseline_{name}.json")
    if not os.path.exists(path):
        pytest.skip(f"{path} not generated (tests/golden/make_baseline_golden.py {name})")
    with open(path) as f:
        return json.load(f)


def test_sym_eigs_full_size_properties(gpu):
    # BASELINE config C2: n = 1e6, nnz/row = 20, k = 20, ncv = 60.  Size-independent properties only.
    from spectra_b200 import synth

    n = 1_000_000
    rp, ci, v = synth.csr(n, 20, 0, True)
    A = sp.csr_matrix((v, ci, rp), shape=(n, n))
    op = gpu.SparseGenMatProd.from_csr_slab(n, 0, rp, ci, v)
    eigs = gpu.SymEigsSolver(op, 20, 60)
    eigs.init()
    nconv = eigs.compute(gpu.SortRule.LargestAlge)
    assert eigs.info() == gpu.CompInfo.Successful and nconv == 20
    evals, X = eigs.eigenvalues(), eigs.eigenvectors()
    assert np.all(np.diff(evals) <= 0)  # sorted LargestAlge
    res = np.linalg.norm(A @ X - X * evals, axis=0) / np.abs(evals)
    assert res.max() <= 1e-10
    assert np.abs(X.T @ X - np.eye(20)).max() <= 1e-10
    # parity with the CPU oracle's full solve of the same problem (tests/golden/baseline_C2.json, make_baseline_golden.py):
    # north_star's "eigenvalues within 1e-10 relative", equal nconv, and the same restart history up to rounding-level decisions
    g = baseline_golden("C2")
    ref = np.array(g["eigenvalues"])
    assert nconv == g["nconv"] and len(evals) == len(ref)
    assert np.abs(evals - ref).max() <= 1e-10 * np.abs(ref).max()
    assert np.abs(evals - ref).max() / np.abs(ref).min() <= 1e-10
    assert abs(eigs.num_iterations() - g["niter"]) <= 2 and abs(eigs.num_operations() - g["nops"]) <= 2 * 40
    # run-to-run bit reproducibility (fixed-order reductions)
    eigs.init()
    eigs.compute(gpu.SortRule.LargestAlge)
    assert np.array_equal(evals, eigs.eigenvalues())


def test_sym_eigs_c4_size_properties(gpu):
    # BASELINE config C4 on one GPU (the bench workload): n = 1e7, nnz/row = 20, k = 20, ncv = 60.  Size-independent properties:
    # north_star's ||Ax - lambda x|| / |lambda| <= 1e-10, orthonormality, ordering; the operator is checked through linearity and
    # symmetry (x'Ay = y'Ax) on the device result.
    from spectra_b200 import synth

    n = 10_000_000
    rp, ci, v = synth.csr(n, 20, 0, True)
    op = gpu.SparseGenMatProd.from_csr_slab(n, 0, rp, ci, v)
    A = sp.csr_matrix((v, ci, rp), shape=(n, n))
    rng = np.random.default_rng(7)
    x, y = rng.standard_normal(n), rng.standard_normal(n)
    Ax, Ay = op.perform_op(x), op.perform_op(y)
    assert abs(y @ Ax - x @ Ay) <= 1e-10 * np.linalg.norm(x) * np.linalg.norm(Ay)
    assert np.abs(op.perform_op(2.0 * x - 3.0 * y) - (2.0 * Ax - 3.0 * Ay)).max() <= 1e-12 * np.abs(Ax).max() * 10
    assert np.abs(Ax[:: n // 1000] - (A @ x)[:: n // 1000]).max() <= 1e-12 * np.abs(Ax).max()
    eigs = gpu.SymEigsSolver(op, 20, 60)
    eigs.init()
    nconv = eigs.compute(gpu.SortRule.LargestAlge)
    assert eigs.info() == gpu.CompInfo.Successful and nconv == 20
    evals, X = eigs.eigenvalues(), eigs.eigenvectors()
    assert np.all(np.diff(evals) <= 0)
    res = np.linalg.norm(A @ X - X * evals, axis=0) / np.abs(evals)
    assert res.max() <= 1e-10
    assert np.abs(X.T @ X - np.eye(20)).max() <= 1e-10
    # parity with the CPU oracle's full solve of the bench workload (tests/golden/baseline_C4.json; hours of CPU time, generated once)
    g = baseline_golden("C4")
    ref = np.array(g["eigenvalues"])
    assert nconv == g["nconv"] and len(evals) == len(ref)
    assert np.abs(evals - ref).max() / np.abs(ref).min() <= 1e-10
    # (a 170-restart history is decided by rounding-level differences many times over: the counts agree to a few per cent, not exactly)
    assert abs(eigs.num_operations() - g["nops"]) <= 0.1 * g["nops"] and abs(eigs.num_iterations() - g["niter"]) <= 0.1 * g["niter"]


def test_sweep_modes_agree(gpu, monkeypatch):
    # The three ways the host can sequence a factorisation -- one round trip per Lanczos step (SB200_SWEEP=0), a whole sweep enqueued behind
    # the device-side abort flag (default), and the opt-in overlapped sweep (two-part correction pass, head blocks of the next operator
    # application on a second stream) -- run the same kernels on the same data: identical histories, eigenvalues equal to rounding
    # (bitwise for the first two, which also use the same reduction order).  Two column blocks, so that the overlapped variant applies.
    from spectra_b200 import synth

    n = 200_000
    rp, ci, v = synth.csr(n, 20, 9, True)
    monkeypatch.setenv("SB200_XSLICE_MB", "0.8")  # 1.6 MB of x -> 2 column blocks
    op = gpu.SparseGenMatProd.from_csr_slab(n, 0, rp, ci, v)
    assert op.spmv_layout()["col_blocks"] == 2
    out = {}
    for name, env in (("sweep", {}), ("per_step", {"SB200_SWEEP": "0"}), ("overlap", {"SB200_OVERLAP": "1"})):
        for k, val in env.items():
            monkeypatch.setenv(k, val)
        e = gpu.SymEigsSolver(op, 10, 30)
        e.init()
        assert e.compute(gpu.SortRule.LargestAlge) == 10
        out[name] = (e.eigenvalues(), e.num_operations(), e.num_iterations(), e.stats()["host_syncs"])
        for k in env:
            monkeypatch.delenv(k)
    assert out["sweep"][1:3] == out["per_step"][1:3] == out["overlap"][1:3]
    assert np.array_equal(out["sweep"][0], out["per_step"][0])
    assert np.abs(out["overlap"][0] - out["sweep"][0]).max() <= 1e-12 * np.abs(out["sweep"][0]).max()
    assert out["sweep"][3] < out["per_step"][3] / 4  # far fewer host synchronisations
    ref = O.sym_eigs(O.Csr.adopt(n, rp, ci, v), 10, 30, O.LargestAlge, want_vectors=False, threads=4)
    assert np.abs(out["sweep"][0] - ref.eigenvalues).max() <= 1e-10 * np.abs(ref.eigenvalues).max()


def test_column_blocked_operator_and_solver(gpu, monkeypatch):
    # Large operands are stored as column blocks so that each SpMV pass gathers from an L2-resident slice of x
    # (csr_build.cu: split_column_blocks).  Force the blocked layout at a small size and compare with the unblocked one.
    from spectra_b200 import synth

    n = 200_000
    rp, ci, v = synth.csr(n, 20, 5, True)
    A = sp.csr_matrix((v, ci, rp), shape=(n, n))
    x = O.simple_random(11, n)
    y0 = A @ x
    monkeypatch.setenv("SB200_XSLICE_MB", "0.25")  # 1.6 MB of x -> 7 column blocks
    op = gpu.SparseGenMatProd.from_csr_slab(n, 0, rp, ci, v)
    assert np.abs(op.perform_op(x) - y0).max() <= 1e-13 * np.abs(y0).max()
    op_sym = gpu.SparseSymMatProd((n, rp, ci, v, "col"))
    assert np.abs(op_sym.perform_op(x) - y0).max() <= 1e-13 * np.abs(y0).max()
    eigs = gpu.SymEigsSolver(op, 10, 30)
    eigs.init()
    eigs.compute(gpu.SortRule.LargestAlge)
    ev_blocked, nops_blocked = eigs.eigenvalues(), eigs.num_operations()
    monkeypatch.setenv("SB200_XSLICE_MB", "1000")
    op1 = gpu.SparseGenMatProd.from_csr_slab(n, 0, rp, ci, v)
    e1 = gpu.SymEigsSolver(op1, 10, 30)
    e1.init()
    e1.compute(gpu.SortRule.LargestAlge)
    assert eigs.info() == gpu.CompInfo.Successful and e1.info() == gpu.CompInfo.Successful
    assert nops_blocked == e1.num_operations()
    assert np.abs(ev_blocked - e1.eigenvalues()).max() <= 1e-12 * np.abs(ev_blocked).max()
    # a nonsymmetric operator through the blocked Arnoldi step head
    rpg, cig, vg = synth.csr(n, 20, 6, False)
    G = sp.csr_matrix((vg, cig, rpg), shape=(n, n))
    monkeypatch.setenv("SB200_XSLICE_MB", "0.25")
    opg = gpu.SparseGenMatProd.from_csr_slab(n, 0, rpg, cig, vg)
    yg = G @ x
    assert np.abs(opg.perform_op(x) - yg).max() <= 1e-13 * np.abs(yg).max()
    g = gpu.GenEigsSolver(opg, 3, 12)
    g.init()
    g.factorize_from(1, 12)
    fz = g.factorization()
    E = G @ fz["V"] - fz["V"] @ fz["H"]
    E[:, -1] -= fz["f"]
    assert np.abs(E).max() <= 1e-12 * max(1.0, np.abs(fz["H"]).max())


@pytest.mark.parametrize("n,ranks,chunks", [(200_003, 3, 4), (1000, 8, 2), (10, 4, 2), (65_537, 2, 5)])
def test_chunked_operand_layout(gpu, monkeypatch, n, ranks, chunks):
    # Row-sharded operators keep their columns in "chunk-major" order (column block c = the c-th part of every rank's slab,
    # ids remapped to positions in the c-th partial all-gather) so that SpMV block c overlaps all-gather c+1.  The test hook
    # SB200_FORCE_CHUNK_RANKS lays a single-GPU operator out the same way; results must not change.
    if n >= 1000:
        from spectra_b200 import synth

        rp, ci, v = synth.csr(n, 20, 3, True)
        A = sp.csr_matrix((v, ci, rp), shape=(n, n))
        k, m = 10, 30
    else:
        A = sym_full(O.gen_sparse_data(n, 0.5)).tocsr()
        A.sort_indices()
        rp, ci, v = A.indptr.astype(np.int64), A.indices, A.data
        k, m = 3, 6
    x = O.simple_random(5, n)
    y0 = A @ x
    monkeypatch.delenv("SB200_FORCE_CHUNK_RANKS", raising=False)
    e0 = gpu.SymEigsSolver(gpu.SparseGenMatProd.from_csr_slab(n, 0, rp, ci, v), k, m)
    e0.init()
    e0.compute(gpu.SortRule.LargestAlge)
    monkeypatch.setenv("SB200_FORCE_CHUNK_RANKS", str(ranks))
    monkeypatch.setenv("SB200_AG_CHUNKS", str(chunks))
    op = gpu.SparseGenMatProd.from_csr_slab(n, 0, rp, ci, v)
    assert np.abs(op.perform_op(x) - y0).max() <= 1e-13 * max(1.0, np.abs(y0).max())
    e1 = gpu.SymEigsSolver(op, k, m)
    e1.init()
    e1.compute(gpu.SortRule.LargestAlge)
    assert e0.info() == gpu.CompInfo.Successful and e1.info() == gpu.CompInfo.Successful
    assert e1.num_operations() == e0.num_operations()
    assert np.abs(e1.eigenvalues() - e0.eigenvalues()).max() <= 1e-12 * np.abs(e0.eigenvalues()).max()
    U = e1.eigenvectors()
    assert np.abs(A @ U - U * e1.eigenvalues()).max() <= 1e-9


def test_user_defined_operator(gpu):
    # the reference's OpType concept (SymEigsSolver.h:99-126): a user class with rows()/cols()/perform_op
    class MyDiagonalTen:
        def rows(self):
            return 10

        def cols(self):
            return 10

        def perform_op(self, x_in, y_out):
            y_out[:] = x_in * np.arange(1, 11)

    op = gpu.UserOp(MyDiagonalTen())
    eigs = gpu.SymEigsSolver(op, 3, 6)
    eigs.init()
    eigs.compute(gpu.SortRule.LargestAlge)
    assert eigs.info() == gpu.CompInfo.Successful
    assert np.allclose(eigs.eigenvalues(), [10, 9, 8], atol=1e-10)
    ref = O.sym_eigs_userop(10, lambda x: x * np.arange(1, 11), 3, 6, selection=O.LargestAlge)
    assert eigs.num_operations() == ref.nops and eigs.num_iterations() == ref.niter
    # a callable operator gives the same answer as the device-resident wrapper of the same matrix
    A = O.gen_sparse_data(1000, 0.01)
    Af = sym_full(A)
    e1 = gpu.SymEigsSolver(gpu.UserOp(lambda x: Af @ x, n=1000), 20, 50)
    e1.init()
    e1.compute(gpu.SortRule.LargestAlge)
    e2 = gpu.SymEigsSolver(gpu.SparseSymMatProd(A), 20, 50)
    e2.init()
    e2.compute(gpu.SortRule.LargestAlge)
    assert e1.info() == gpu.CompInfo.Successful
    assert np.abs(e1.eigenvalues() - e2.eigenvalues()).max() <= 1e-11 * np.abs(e2.eigenvalues()).max()
    U = e1.eigenvectors()
    assert np.abs(Af @ U - U * e1.eigenvalues()).max() <= 1e-9
    # nonsymmetric user operator through GenEigsSolver
    G = O.gen_sparse_data(100, 0.1).tocsr()
    g = gpu.GenEigsSolver(gpu.UserOp(lambda x: G @ x, n=100), 10, 30)
    g.init()
    g.compute(gpu.SortRule.LargestMagn, 300)
    assert g.info() == gpu.CompInfo.Successful
    Z = g.eigenvectors()
    assert np.abs(G @ Z - Z * g.eigenvalues()).max() <= 1e-9
    # an exception raised inside the Python operator surfaces from the call that ran it
    calls = {"n": 0}

    def failing(x):
        calls["n"] += 1
        if calls["n"] == 4:
            raise KeyError("operator failed")
        return x * np.arange(1, 11)

    bad = gpu.SymEigsSolver(gpu.UserOp(failing, n=10), 3, 6)
    with pytest.raises(KeyError):
        bad.init()
        bad.compute(gpu.SortRule.LargestAlge)


@pytest.mark.parametrize("n", [10, 100, 1000])
def test_sym_eigs_against_golden_spectra(gpu, n):
    # tests/golden/kat_spectra.json: dense-LAPACK spectra of the reference's own fixtures (make_golden.py)
    import golden_cases as GC

    prob = {10: 0.5, 100: 0.1, 1000: 0.01}[n]
    k, m = GC.KM[n]
    A = O.gen_sparse_data(n, prob)
    op = gpu.SparseSymMatProd(A)
    for rule, srule in ((O.LargestAlge, gpu.SortRule.LargestAlge), (O.SmallestAlge, gpu.SortRule.SmallestAlge), (O.LargestMagn, gpu.SortRule.LargestMagn)):
        e = gpu.SymEigsSolver(op, k, m)
        e.init()
        assert e.compute(srule) == k
        GC.check_sym_values("sym", n, rule, e.eigenvalues())

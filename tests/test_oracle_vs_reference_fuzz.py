"""Property-based pinning of the restatement on the reference's own code (hypothesis): random and degenerate inputs through both,
`np.array_equal` on everything (both sides are -ffp-contract=off builds: same operations in the same order => same bits).
See tests/test_oracle_vs_reference.py for what oracle/_ref is.  CPU only."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

import oracle as O
from oracle import ref as R

pytestmark = pytest.mark.skipif(not R.available(), reason="oracle/_ref is not built and /root/reference is not present")
SETTINGS = dict(max_examples=400, deadline=None, derandomize=True, suppress_health_check=[HealthCheck.too_slow])


@pytest.fixture(autouse=True, scope="module")
def _strict_restatement():
    prev = O.select_build("strict")
    yield
    O.select_build(prev)


def _eq(a, b):
    return np.array_equal(np.asarray(a), np.asarray(b))


@st.composite
def tridiagonal(draw):
    m = draw(st.integers(2, 40))
    rng = np.random.default_rng(draw(st.integers(0, 2**31 - 1)))
    d = rng.standard_normal(m) * 10.0 ** draw(st.integers(-3, 3))
    e = rng.standard_normal(m - 1) * 10.0 ** draw(st.integers(-12, 2))
    kind = draw(st.sampled_from(["plain", "zeros", "tiny", "repeated", "graded"]))
    if kind == "zeros":
        e[rng.integers(0, m - 1, max(1, (m - 1) // 3))] = 0.0        # decoupled blocks (converged Ritz pairs)
    elif kind == "tiny":
        e[rng.integers(0, m - 1, max(1, (m - 1) // 3))] *= 1e-18     # below the deflation thresholds
    elif kind == "repeated":
        d[:] = d[0]
    elif kind == "graded":
        d *= np.logspace(0, -10, m)
    return np.diag(d) + np.diag(e, 1) + np.diag(e, -1)


@settings(**SETTINGS)
@given(T=tridiagonal(), shift_pick=st.integers(0, 3))
def test_fuzz_tridiagonal_kernels(T, shift_pick):
    m = T.shape[0]
    for a, b in zip(R.tridiag_eigen(T), O.tridiag_eigen(T)):
        assert _eq(a, b)
    w = np.linalg.eigvalsh(T)
    shift = [0.0, float(w[0]), float(w[-1]), float(w[m // 2]) * (1 + 1e-9)][shift_pick]
    for a, b in zip(R.shifted_qr(T, shift, "tridiag"), O.shifted_qr(T, shift, "tridiag")):
        assert _eq(a, b)
    for a, b in zip(R.shifted_qr(T, shift, "hess"), O.shifted_qr(T, shift, "hess")):
        assert _eq(a, b)


@st.composite
def hessenberg(draw):
    m = draw(st.integers(3, 30))
    rng = np.random.default_rng(draw(st.integers(0, 2**31 - 1)))
    H = np.triu(rng.standard_normal((m, m)), -1) * 10.0 ** draw(st.integers(-2, 2))
    kind = draw(st.sampled_from(["plain", "split", "symmetric", "rotation", "tiny_sub"]))
    if kind == "split":
        H[rng.integers(1, m), :][...] = H[rng.integers(1, m), :]
        i = int(rng.integers(1, m))
        H[i, i - 1] = 0.0
    elif kind == "symmetric":
        H = np.triu(H, -1)
        H = np.tril(H, 1)
        H = (H + H.T) / 2  # tridiagonal symmetric: real spectrum
    elif kind == "rotation":
        H = np.triu(H, -1)
        for i in range(0, m - 1, 2):  # 2x2 blocks with complex pairs
            H[i, i] = H[i + 1, i + 1] = 0.3
            H[i, i + 1], H[i + 1, i] = 1.0, -1.0
    elif kind == "tiny_sub":
        H[np.arange(1, m), np.arange(0, m - 1)] *= 1e-14
    return H


@settings(**SETTINGS)
@given(H=hessenberg(), s=st.floats(-3, 3), t=st.floats(0, 9))
def test_fuzz_hessenberg_kernels(H, s, t):
    for a, b in zip(R.hess_schur(H), O.hess_schur(H)):
        assert _eq(a, b)
    (rev, rV), (oev, oV) = R.hess_eigen(H), O.hess_eigen(H)
    assert _eq(rev, oev) and _eq(rV, oV)
    for a, b in zip(R.double_shift_qr(H, s, t), O.double_shift_qr(H, s, t)):
        assert _eq(a, b)
    for a, b in zip(R.shifted_qr(H, s, "hess"), O.shifted_qr(H, s, "hess")):
        assert _eq(a, b)


@settings(max_examples=150, deadline=None, derandomize=True, suppress_health_check=[HealthCheck.too_slow])
@given(n=st.integers(6, 60), seed=st.integers(0, 2**31 - 1), rule=st.sampled_from([O.LargestMagn, O.LargestAlge, O.SmallestAlge, O.BothEnds]),
       kind=st.sampled_from(["dense", "low_rank", "repeated", "diag"]))
def test_fuzz_symmetric_solves_bit_identical(n, seed, rule, kind):
    # complete SymEigsSolver solves over ONE user-defined operator on both sides: random, rank-deficient (expand_basis / beta = 0 paths),
    # repeated-eigenvalue and diagonal matrices; every step, every restart, every rare path must agree to the bit
    rng = np.random.default_rng(seed)
    if kind == "dense":
        M = rng.standard_normal((n, n))
        M = M + M.T
    elif kind == "low_rank":
        r = max(1, n // 4)
        B = rng.standard_normal((n, r))
        M = B @ B.T
    elif kind == "repeated":
        Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
        w = np.repeat(rng.standard_normal(n // 3 + 1), 3)[:n]
        M = (Q * w) @ Q.T
        M = (M + M.T) / 2
    else:
        M = np.diag(rng.integers(1, 5, n).astype(float))
    k = int(rng.integers(1, max(2, n // 3)))
    m = int(min(n, max(k + 2, 2 * k + 1)))
    fn = lambda x: M @ x  # noqa: E731
    r = R.sym_eigs_userop(n, fn, k, m, selection=rule, maxit=200)
    o = O.sym_eigs_userop(n, fn, k, m, selection=rule, maxit=200)
    assert (r.info, r.nconv, r.niter, r.nops) == (o.info, o.nconv, o.niter, o.nops)
    assert _eq(r.eigenvalues, o.eigenvalues)
    if r.nconv:
        assert _eq(r.eigenvectors, o.eigenvectors)


@settings(max_examples=150, deadline=None, derandomize=True, suppress_health_check=[HealthCheck.too_slow])
@given(n=st.integers(8, 80), seed=st.integers(0, 2**31 - 1),
       rule=st.sampled_from([O.LargestMagn, O.LargestReal, O.LargestImag, O.SmallestMagn, O.SmallestReal, O.SmallestImag]), density=st.floats(0.05, 0.6))
def test_fuzz_general_solves_bit_identical(n, seed, rule, density):
    # complete GenEigsSolver solves: row-major storage with ascending columns makes the reference's product sum like the restatement's
    import scipy.sparse as sp

    rng = np.random.default_rng(seed)
    A = sp.random(n, n, density, random_state=seed, format="csr") + sp.diags(rng.standard_normal(n))
    A = sp.csr_matrix(A)
    A.sort_indices()
    k = int(rng.integers(1, max(2, n // 4)))
    m = int(min(n, max(k + 3, 2 * k + 2)))
    r = R.gen_eigs(R.Compressed.from_scipy(A), k, m, rule, 100)
    o = O.gen_eigs(O.Csr(n, A.indptr, A.indices, A.data, order="row", mode="gen"), k, m, rule, 100)
    assert (r.info, r.nconv, r.niter, r.nops) == (o.info, o.nconv, o.niter, o.nops)
    assert _eq(r.eigenvalues, o.eigenvalues)
    if r.nconv:
        assert _eq(r.eigenvectors, o.eigenvectors)

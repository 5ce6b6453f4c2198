"""Shared fixtures of the parity tests: the portable known-answer inputs of the reference's tests."""
import numpy as np
import scipy.sparse as sp


def cycle_laplacian(n):
    """test/Example1.cpp:18-32"""
    L = np.zeros((n, n))
    for i in range(n):
        L[i, i] = 1
        L[i, (i + n - 1) % n] = -0.5
        L[i, (i + 1) % n] = -0.5
    return L


# test/Example2.cpp:51-84 (issue #159): near rank-1 matrices
EXAMPLE2 = [
    np.array([
        [15.035447086947079479, 3.932587856183598677, -4.848070276813470542, -8.027254936523050904, -2.865327349780228231],
        [3.932587856183598677, 1.028585791773944732, -1.268034278346991263, -2.099564123322002035, -0.749439073848281425],
        [-4.848070276813470542, -1.268034278346991263, 1.563224909309606855, 2.588329820664053864, 0.923903910371237535],
        [-8.027254936523050904, -2.099564123322002035, 2.588329820664053864, 4.285660509016328222, 1.529765824738644411],
        [-2.865327349780228231, -0.749439073848281425, 0.923903910371237535, 1.529765824738644411, 0.546049663433429209]]),
    np.array([
        [0.6118330552, -3.058379358, 1.329013596, 2.601267208, 1.072783220],
        [-3.058379358, 15.28796821, -6.643360824, -13.00299463, -5.362538075],
        [1.329013596, -6.643360824, 2.886861251, 5.650429406, 2.330281884],
        [2.601267208, -13.00299463, 5.650429406, 11.05953826, 4.561041261],
        [1.072783220, -5.362538075, 2.330281884, 4.561041261, 1.881009576]]),
    np.array([
        [17.7699571312182, 10.7033479738827, -19.1658731825582, -4.20053658859459, -11.1426294187651],
        [10.7033479738827, 6.44692933157151, -11.5441477084849, -2.53010203979439, -6.71152097511499],
        [-19.1658731825582, -11.5441477084849, 20.6714451890590, 4.53050904744533, 12.0179368348118],
        [-4.20053658859459, -2.53010203979439, 4.53050904744533, 0.992940360059961, 2.63394122006329],
        [-11.1426294187651, -6.71152097511499, 12.0179368348118, 2.63394122006329, 6.98697185632535]]),
]


def readme_banded(n=10):
    """README.md:146-178: 1 on the diagonal, 2 below, 3 above."""
    M = np.zeros((n, n))
    for i in range(n):
        M[i, i] = 1.0
        if i > 0:
            M[i - 1, i] = 3.0
        if i < n - 1:
            M[i + 1, i] = 2.0
    return M


def dense_as_csc(M):
    """A dense matrix in Eigen's compressed ColMajor layout (every entry stored)."""
    return sp.csc_matrix(M + 0.0) if False else sp.csc_matrix((np.asfortranarray(M).reshape(-1, order="F"),
                                                                    np.tile(np.arange(M.shape[0], dtype=np.int32), M.shape[1]),
                                                                    np.arange(0, M.size + 1, M.shape[0], dtype=np.int32)), shape=M.shape)


def sym_full(A_lower_source):
    """selfadjointView<Lower> of a (possibly non-symmetric) stored matrix as a dense/sparse symmetric matrix."""
    L = sp.tril(sp.csr_matrix(A_lower_source), 0)
    D = sp.diags(L.diagonal())
    return (L + L.T - D).tocsr()


def stencil_matrix(dims, full=True, seed=1, diag_shift=0.3):
    """A symmetric mesh-like matrix in natural (lexicographic) ordering: random weights on the edges of a grid graph whose nodes are
    coupled to all neighbours within Chebyshev distance 1 (`full`: 9-point in 2-D, 27-point in 3-D) or to the axis neighbours only
    (5- / 7-point), random diagonal.  Half-bandwidth: the stride of the slowest index (+ lower strides + 1 when `full`).  CSC."""
    import itertools

    n = int(np.prod(dims))
    idx = np.arange(n).reshape(dims)
    rng = np.random.default_rng(seed)
    rows, cols = [], []
    offs = [o for o in itertools.product(*[(-1, 0, 1)] * len(dims)) if any(o) and (full or sum(abs(x) for x in o) == 1)]
    for o in offs:
        src = tuple(slice(max(0, -d), dims[a] - max(0, d)) for a, d in enumerate(o))
        dst = tuple(slice(max(0, d), dims[a] - max(0, -d)) for a, d in enumerate(o))
        rows.append(idx[src].ravel())
        cols.append(idx[dst].ravel())
    r, c = np.concatenate(rows), np.concatenate(cols)
    A = sp.csr_matrix((rng.uniform(-1, 1, r.size), (r, c)), shape=(n, n))
    A = (A + A.T) / 2 + sp.diags(rng.uniform(-1, 1, n) + diag_shift)
    return A.tocsc()

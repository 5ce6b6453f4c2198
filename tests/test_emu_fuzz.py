"""Property-based checks of the operator layer on the kernel-logic emulator (hypothesis): random compressed matrices -- empty rows and
columns, dense rows, duplicate entries, unsorted inner indices, 32- and 64-bit outer indices, both storage orders, every matrix mode --
through csr_build.cu (count / scan / fill / sort, column blocks, sliced layout) and the SpMV kernels, against a dense numpy model of
the reference's semantics (SparseGenMatProd.h:82-87, SparseSymMatProd.h:83-88 and SparseHermMatProd.h:83-88: selfadjointView<Uplo>
reads one triangle and mirrors it; duplicates of an uncompressed-but-valid input add up)."""
import os

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

import test_gpu_layouts_complex as X


def _dense_model(n, outer, inner, vals, order, mode):
    M = np.zeros((n, n), dtype=vals.dtype)
    for o in range(n):
        for p in range(outer[o], outer[o + 1]):
            i, j = (inner[p], o) if order == "col" else (o, inner[p])
            M[i, j] += vals[p]
    if mode == "gen":
        return M
    T = np.tril(M, -1) if mode in ("lower", "hlower") else np.triu(M, 1)
    D = np.diag(np.diag(M).real) if mode.startswith("h") else np.diag(np.diag(M))
    return T + (T.conj().T if mode.startswith("h") else T.T) + D


@st.composite
def compressed_matrix(draw, complex_values=False):
    n = draw(st.integers(1, 70))
    nnz_per = draw(st.lists(st.integers(0, min(n, 9)), min_size=n, max_size=n))
    if draw(st.booleans()):
        nnz_per[draw(st.integers(0, n - 1))] = n  # one dense outer slice
    outer = np.concatenate([[0], np.cumsum(nnz_per)]).astype(np.int64 if draw(st.booleans()) else np.int32)
    total = int(outer[-1])
    seed = draw(st.integers(0, 2**31 - 1))
    rng = np.random.default_rng(seed)
    inner = rng.integers(0, n, total).astype(np.int32)  # duplicates and unsorted order on purpose
    vals = rng.standard_normal(total)
    if complex_values:
        vals = vals + 1j * rng.standard_normal(total)
    order = draw(st.sampled_from(["col", "row"]))
    return n, outer, inner, vals, order


FMT = st.sampled_from([{}, {"SB200_SPMV_FORMAT": "sell", "SB200_SELL_MAX_FILL": "1000"}, {"SB200_FORCE_CHUNK_RANKS": "2", "SB200_AG_CHUNKS": "3"},
                       {"SB200_SPMV_FORMAT": "sell", "SB200_SELL_MAX_FILL": "1000", "SB200_FORCE_CHUNK_RANKS": "3", "SB200_AG_CHUNKS": "2"}])
SETTINGS = dict(max_examples=200, deadline=None, derandomize=True, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])


@settings(**SETTINGS)
@given(mat=compressed_matrix(), mode=st.sampled_from(["gen", "lower", "upper"]), fmt=FMT, xseed=st.integers(0, 1000))
def test_fuzz_real_operator(emu, mat, mode, fmt, xseed):
    n, outer, inner, vals, order = mat
    A = _dense_model(n, outer, inner, vals, order, mode)
    with X.env(**fmt):
        op = emu.SparseGenMatProd((n, outer, inner, vals, order)) if mode == "gen" else emu.SparseSymMatProd((n, outer, inner, vals, order), uplo=mode)
    x = np.random.default_rng(xseed).standard_normal(n)
    y, y0 = op.perform_op(x), A @ x
    assert np.abs(y - y0).max() <= 1e-12 * max(1.0, np.abs(y0).max(), np.abs(A).sum(axis=1).max() * np.abs(x).max())
    assert op.nnz == _stored_count(n, outer, inner, order, mode)


def _stored_count(n, outer, inner, order, mode):
    cnt = 0
    for o in range(n):
        for p in range(outer[o], outer[o + 1]):
            i, j = (inner[p], o) if order == "col" else (o, inner[p])
            if mode == "gen":
                cnt += 1
            elif (mode.endswith("lower") and i >= j) or (mode.endswith("upper") and i <= j):
                cnt += 1 if i == j else 2
    return cnt


@settings(**SETTINGS)
@given(mat=compressed_matrix(complex_values=True), mode=st.sampled_from(["hlower", "hupper", "gen"]), xseed=st.integers(0, 1000))
def test_fuzz_complex_operator(emu, mat, mode, xseed):
    n, outer, inner, vals, order = mat
    A = _dense_model(n, outer, inner, vals, order, mode)
    op = emu.SparseHermMatProd((n, outer, inner, vals, order), uplo={"hlower": "lower", "hupper": "upper", "gen": "general"}[mode])
    rng = np.random.default_rng(xseed)
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    y, y0 = op.perform_op(x), A @ x
    assert np.abs(y - y0).max() <= 1e-12 * max(1.0, np.abs(y0).max(), np.abs(A).sum(axis=1).max() * np.abs(x).max())
    assert op.nnz == _stored_count(n, outer, inner, order, mode)
    if mode != "gen":
        assert abs(np.vdot(x, y).imag) <= 1e-10 * max(1.0, abs(np.vdot(x, y)))  # Hermitian operator: x^H A x is real


@st.composite
def small_sym_problem(draw):
    n = draw(st.integers(6, 40))
    kind = draw(st.sampled_from(["random", "low_rank", "diag_repeated", "sparse"]))
    rng = np.random.default_rng(draw(st.integers(0, 2**31 - 1)))
    if kind == "random":
        M = rng.standard_normal((n, n))
    elif kind == "low_rank":
        u = rng.standard_normal((n, 2))
        M = u @ u.T  # rank 2: the Krylov space is exhausted early -> expand_basis / restart heuristics (Lanczos.h:99-121)
    elif kind == "diag_repeated":
        M = np.diag(rng.integers(1, 4, n).astype(float))  # repeated eigenvalues (test/Example1.cpp's difficulty)
    else:
        M = rng.standard_normal((n, n)) * (rng.random((n, n)) < 0.15)
    k = draw(st.integers(1, max(1, min(5, n - 2))))
    m = draw(st.integers(k + 1, min(n, k + 12)))
    rule = draw(st.sampled_from(["LargestAlge", "SmallestAlge", "LargestMagn", "BothEnds"]))
    return M, k, m, rule, kind


@settings(max_examples=40, deadline=None, derandomize=True, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])
@given(prob=small_sym_problem())
def test_fuzz_sym_solver_follows_oracle(emu, prob):
    # whole pipeline (init, factorisation with its restart heuristics, device restart kernel, compress) against the CPU oracle on random and
    # degenerate inputs: same convergence outcome, same operation count, same eigenvalues
    import oracle as O
    from helpers import dense_as_csc

    M, k, m, rule, kind = prob
    A = dense_as_csc(M)
    ref = O.sym_eigs(O.Csr.from_scipy(A, "lower"), k, m, getattr(O, rule), 200)
    op = emu.SparseSymMatProd(A)
    e = emu.SymEigsSolver(op, k, m)
    e.init()
    nconv = e.compute(getattr(emu.SortRule, rule), 200)
    scale = max(1.0, np.abs(M).max() * M.shape[0])
    if kind in ("random", "sparse"):
        # generic inputs: the device follows the oracle step for step
        assert nconv == ref.nconv and int(e.info()) == ref.info
        assert e.num_operations() == ref.nops and e.num_iterations() == ref.niter
        assert np.abs(e.eigenvalues() - ref.eigenvalues).max(initial=0.0) <= 1e-10 * scale
    else:
        # exhausted Krylov spaces (low rank, repeated eigenvalues): the residual is rounding noise, and whether ||f|| falls below
        # eps*sqrt(n) (Lanczos.h:163-168) or triggers expand_basis depends on the summation order -- in the reference as much as
        # here -- so only the outcome is compared: what converged are eigenvalues of the matrix, and as many as the oracle found
        Ml = np.tril(M) + np.tril(M, -1).T
        w = np.linalg.eigvalsh(Ml)
        for ev in e.eigenvalues():
            assert np.abs(w - ev).min() <= 1e-9 * scale
        assert nconv == ref.nconv

"""GPU tests of the sliced (SELL-32) operand layout and of the complex scalar paths.  First device run: round 2, call 1
(profiles/r2_layout_complex_tests.log, 94 passed); since then part of the default `-m gpu` suite.

Covered: the sliced layout + lane-per-row SpMV kernels (the default layout; csr_build.cu build_sell_layout, spmv.cu sell_plain_kernel /
sell_step_kernel) against the CSR-vector kernels (SB200_SPMV_FORMAT=csr) and SciPy at the operator, factorisation and solver tiers; the
gather microbenchmark; the complex Hermitian path (SparseHermMatProd / HermEigsSolver, user-defined complex operators); the complex
GenEigsSolver (complex Arnoldi, the one-warp complex restart kernel of dense_gen_z.cu, complex restart GEMM); the C++ shim flows for
complex and float scalars."""
import contextlib
import os

import numpy as np
import pytest
import scipy.sparse as sp

import oracle as O
from helpers import sym_full

pytestmark = [pytest.mark.gpu]


@contextlib.contextmanager
def env(**kw):
    old = {k: os.environ.get(k) for k in kw}
    try:
        for k, v in kw.items():
            os.environ[k] = str(v)
        yield
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _rand_csr(n, density, seed):
    A = sp.random(n, n, density=density, random_state=seed, format="csr")
    A.data -= 0.5
    return A


@pytest.mark.parametrize("threads", [256, 512, 1024])
@pytest.mark.parametrize("n,density", [(1, 1.0), (31, 0.3), (32, 0.3), (33, 0.3), (1000, 0.01), (1023, 0.02), (1024, 0.02), (1025, 0.02), (5000, 0.004), (40_000, 0.0005)])
def test_sell_spmv_matches_csr_and_scipy(gpu, threads, n, density):
    rng = np.random.default_rng(n)
    A = _rand_csr(n, density, n)
    x = rng.standard_normal(n)
    y0 = A @ x
    with env(SB200_SPMV_FORMAT="csr"):
        op_csr = gpu.SparseGenMatProd(A)
    assert op_csr.spmv_layout()["format"] == "csr"
    y_csr = op_csr.perform_op(x)
    with env(SB200_SPMV_FORMAT="sell", SB200_SELL_THREADS=threads, SB200_SELL_MAX_FILL=100):
        op = gpu.SparseGenMatProd(A)
    lay = op.spmv_layout()
    assert lay["format"] == "sell" and lay["stored_entries"] >= A.nnz and lay["stored_entries"] % 32 == 0
    y = op.perform_op(x)
    scale = max(1.0, np.abs(y0).max())
    assert np.abs(y - y0).max() <= 1e-13 * scale
    assert np.abs(y - y_csr).max() <= 1e-13 * scale
    assert np.array_equal(y, op.perform_op(x))  # run-to-run bit reproducibility
    M = rng.standard_normal((n, 3))
    assert np.abs(op @ M - A @ M).max() <= 1e-13 * max(1.0, np.abs(A @ M).max())


def test_sell_rejects_wasteful_padding(gpu):
    # one dense row among very short ones: the sliced layout would store far more than max_fill x nnz -> CSR kernels stay in charge
    n = 3000
    rng = np.random.default_rng(2)
    A = sp.random(n, n, density=0.0005, random_state=3, format="lil")
    A[7, :] = rng.standard_normal(n)
    A = A.tocsr()
    with env(SB200_SPMV_FORMAT="sell"):
        op = gpu.SparseGenMatProd(A)
    assert op.spmv_layout()["format"] == "csr"
    x = rng.standard_normal(n)
    assert np.abs(op.perform_op(x) - A @ x).max() <= 1e-12 * np.abs(A @ x).max()
    # accepted when the caller raises the limit; a dense row is handled by one lane (slow, correct)
    with env(SB200_SPMV_FORMAT="sell", SB200_SELL_MAX_FILL=1000):
        op2 = gpu.SparseGenMatProd(A)
    assert op2.spmv_layout()["format"] == "sell"
    assert np.abs(op2.perform_op(x) - A @ x).max() <= 1e-12 * np.abs(A @ x).max()


def test_sell_empty_rows_and_empty_matrix(gpu):
    with env(SB200_SPMV_FORMAT="sell", SB200_SELL_MAX_FILL=100):
        Z = gpu.SparseSymMatProd(sp.csc_matrix((50, 50)))
        assert np.array_equal(Z.perform_op(np.ones(50)), np.zeros(50))
        n = 2500
        A = _rand_csr(n, 0.002, 5).tolil()
        A[11, :] = 0
        A[1024:1100, :] = 0
        A = A.tocsr()
        op = gpu.SparseGenMatProd(A)
    x = np.random.default_rng(0).standard_normal(n)
    y = op.perform_op(x)
    assert np.abs(y - A @ x).max() <= 1e-13 * np.abs(A @ x).max()
    assert y[11] == 0.0 and np.all(y[1024:1100] == 0.0)


@pytest.mark.parametrize("n,force_ranks,chunks", [(400_000, 0, 0), (100_003, 2, 3), (65_537, 3, 4)])
def test_sell_with_column_blocks_and_chunk_layout(gpu, n, force_ranks, chunks):
    # column blocks (SB200_XSLICE_MB=1: 3 MB of x -> 4 blocks, accumulate kernels) and the chunk-major layout of sharded operators
    from spectra_b200 import synth

    rp, ci, v = synth.csr(n, 20, 0, True)
    A = sp.csr_matrix((v, ci, rp), shape=(n, n))
    kw = dict(SB200_SPMV_FORMAT="sell", SB200_XSLICE_MB=1)
    if force_ranks:
        kw.update(SB200_FORCE_CHUNK_RANKS=force_ranks, SB200_AG_CHUNKS=chunks)
    with env(**kw):
        op = gpu.SparseGenMatProd.from_csr_slab(n, 0, rp, ci, v)
    lay = op.spmv_layout()
    assert lay["format"] == "sell" and lay["col_blocks"] >= 2
    assert lay["stored_entries"] <= 1.3 * A.nnz
    x = O.simple_random(7, n)
    y0 = A @ x
    assert np.abs(op.perform_op(x) - y0).max() <= 1e-13 * np.abs(y0).max()


@pytest.mark.parametrize("threads", [256, 512, 1024])
def test_sell_lanczos_factorization(gpu, threads):
    # test/Arnoldi.cpp:19-85 through the fused step kernel of the sliced layout
    n, m = 3000, 30
    A = sp.random(n, n, density=0.004, random_state=5, format="csc")
    Af = sym_full(A)
    with env(SB200_SPMV_FORMAT="sell", SB200_SELL_THREADS=threads, SB200_SELL_MAX_FILL=100, SB200_FORCE_CHUNK_RANKS=2, SB200_AG_CHUNKS=3):
        op = gpu.SparseSymMatProd(A)
    assert op.spmv_layout()["format"] == "sell" and op.spmv_layout()["col_blocks"] >= 2
    s = gpu.SymEigsSolver(op, 10, m)
    v0 = O.simple_random(3, n)
    s.init(v0)
    s.factorize_from(1, m // 2)
    s.factorize_from(m // 2, m)
    fz = s.factorization()
    V, H, f = fz["V"], fz["H"], fz["f"]
    R = Af @ V - V @ H
    R[:, -1] -= f
    scale = max(1.0, np.abs(H).max())
    assert np.abs(R).max() <= 1e-12 * scale
    assert np.abs(V.T @ V - np.eye(m)).max() <= 1e-12
    assert abs(np.linalg.norm(f) - fz["beta"]) <= 1e-12 * scale
    assert np.abs(H - H.T).max() == 0.0 and np.abs(np.triu(H, 2)).max() == 0.0
    ref = O.factorize(O.Csr.from_scipy(A, "lower"), m, v0=v0)
    assert np.abs(H - ref["H"]).max() <= 1e-9 * scale


@pytest.mark.parametrize("which", ["sym", "gen"])
def test_sell_solver_matches_oracle(gpu, which):
    A = O.gen_sparse_data(1000, 0.01)  # test/SymEigs.cpp:157-167
    with env(SB200_SPMV_FORMAT="sell", SB200_SELL_MAX_FILL=100):
        op = gpu.SparseSymMatProd(A) if which == "sym" else gpu.SparseGenMatProd(A)
    assert op.spmv_layout()["format"] == "sell"
    if which == "sym":
        eigs = gpu.SymEigsSolver(op, 20, 50)
        eigs.init()
        nconv = eigs.compute(gpu.SortRule.LargestAlge)
        assert eigs.info() == gpu.CompInfo.Successful and nconv == 20
        ev, U = eigs.eigenvalues(), eigs.eigenvectors()
        ref = O.sym_eigs(O.Csr.from_scipy(A, "lower"), 20, 50, O.LargestAlge)
        Af = O.Csr.from_scipy(A, "lower").to_scipy()
        assert np.abs(Af @ U - U * ev).max() <= 1e-9
        assert np.abs(ev - ref.eigenvalues).max() <= 1e-10 * np.abs(ref.eigenvalues).max()
    else:
        g = gpu.GenEigsSolver(op, 10, 30)
        g.init()
        g.compute(gpu.SortRule.LargestMagn, 300)
        assert g.info() == gpu.CompInfo.Successful
        ev, Z = g.eigenvalues(), g.eigenvectors()
        assert np.abs(A @ Z - Z * ev).max() <= 1e-9


def test_sell_mid_size_solve_against_oracle(gpu):
    from spectra_b200 import synth

    n = 50_000
    rp, ci, v = synth.csr(n, 20, 0, True)
    A = sp.csr_matrix((v, ci, rp), shape=(n, n))
    with env(SB200_SPMV_FORMAT="sell"):
        op = gpu.SparseSymMatProd((n, rp, ci, v, "col"))
    assert op.spmv_layout()["format"] == "sell"
    eigs = gpu.SymEigsSolver(op, 10, 30)
    eigs.init()
    nconv = eigs.compute(gpu.SortRule.LargestAlge)
    assert nconv == 10
    ev, U = eigs.eigenvalues(), eigs.eigenvectors()
    ref = O.sym_eigs(O.Csr.adopt(n, rp, ci, v), 10, 30, O.LargestAlge, want_vectors=False)
    assert np.abs(ev - ref.eigenvalues).max() <= 1e-10 * np.abs(ref.eigenvalues).max()
    r = np.linalg.norm(A @ U - U * ev, axis=0) / np.abs(ev)
    assert r.max() <= 1e-10


def test_gather_microbenchmark_runs(gpu):
    r = gpu.bench_gather(1_000_000, 10_000_000, 2)
    assert r["ms"] > 0 and abs(r["checksum"] - 1.0) < 1e-12


# ---------------------------------------------------------------- complex Hermitian path
import herm_cases as HC  # noqa: E402


@pytest.mark.parametrize("fmt,uplo", [("csc", "lower"), ("csr", "lower"), ("csc", "upper"), ("csr", "upper")])
@pytest.mark.parametrize("n", [10, 100, 1000])
def test_sparse_herm_mat_prod(gpu, n, fmt, uplo):
    HC.operator_case(gpu, n, fmt, uplo)


def test_herm_lanczos_factorization(gpu):
    HC.factorization_case(gpu)
    HC.factorization_case(gpu, n=5000, m=63)


@pytest.mark.parametrize("selection", [O.LargestMagn, O.LargestAlge, O.SmallestMagn, O.SmallestAlge, O.BothEnds])
@pytest.mark.parametrize("n", [10, 100, 1000])
def test_herm_eigs_reference_cases(gpu, n, selection):
    # test/HermEigs.cpp:140-174
    HC.solver_case(gpu, n, selection)


def test_herm_argument_checks(gpu):
    HC.argument_checks(gpu)


def test_herm_user_operator(gpu):
    HC.user_operator_case(gpu)


def test_complex_arnoldi_factorization(gpu):
    HC.complex_arnoldi_factorization_case(gpu)
    HC.complex_arnoldi_factorization_case(gpu, n=20_000, m=40)


def test_complex_gen_user_operator(gpu):
    HC.complex_gen_user_operator_case(gpu)


@pytest.mark.parametrize("m", [2, 3, 6, 20, 50, 63])
def test_complex_dense_kernels(gpu, m):
    HC.complex_dense_kernels_case(gpu, m)


@pytest.mark.parametrize("rule", ["LargestMagn", "LargestReal", "LargestImag", "SmallestReal"])
@pytest.mark.parametrize("n", [10, 100, 1000])
def test_complex_gen_eigs_reference_cases(gpu, n, rule):
    # test/ComplexEigs.cpp:151-192
    HC.complex_gen_solver_case(gpu, n, rule)


def test_herm_shim_reference_flow_on_gpu(gpu):
    # test/HermEigs.cpp's sparse flow through the C++ shim headers against the CUDA library
    import subprocess

    import test_cpp_shim as TS

    exe = TS._compile_herm(os.path.dirname(gpu.lib_path()), os.path.basename(gpu.lib_path())[3:-3])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ALL PASSED" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]


def test_float_shim_flow_on_gpu(gpu):
    import subprocess

    import test_cpp_shim as TS

    exe = TS._compile_herm(os.path.dirname(gpu.lib_path()), os.path.basename(gpu.lib_path())[3:-3], "test_shim_float.cpp")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ALL PASSED" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]

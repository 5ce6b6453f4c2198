"""Kernel-logic checks WITHOUT a GPU: the product's own .cu sources compiled for the CPU against the CUDA execution model of
tools/cuda_emu (fibers for the threads of a CTA, real barriers / shuffles / shared memory / atomics; see cuda_emu.h), driven through
the same C ABI and Python mirror as on a device and compared with the CPU oracle.

What this does and does not show: indexing, barrier placement, shuffle patterns, reduction orders, host sequencing and the
reference's control flow are exercised statement by statement; hardware behaviour (memory model, PTX paths such as the TMA/DMMA
restart GEMM, performance) is not -- the `-m gpu` suite remains the parity gate.  The emulator is test infrastructure like oracle/:
nothing in spectra_b200/ can load it.

The cases are small twins of the GPU tests (sizes chosen so that the file runs in a couple of minutes)."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

import herm_cases as HC
import oracle as O
import test_gpu_layouts_complex as X
from helpers import sym_full


# ---------------------------------------------------------------- default product path (verified on a B200 in round 1)
def test_emu_default_operator_and_solver(emu):
    A = O.gen_sparse_data(100, 0.1)
    op = emu.SparseSymMatProd(A)
    ref = O.Csr.from_scipy(A, "lower")
    x = np.random.default_rng(0).standard_normal(100)
    assert np.abs(op.perform_op(x) - ref.spmv(x)).max() <= 1e-13 * np.abs(ref.spmv(x)).max()
    eigs = emu.SymEigsSolver(op, 10, 20)
    eigs.init()
    assert eigs.compute(emu.SortRule.LargestAlge) == 10 and eigs.info() == emu.CompInfo.Successful
    r = O.sym_eigs(ref, 10, 20, O.LargestAlge)
    assert np.abs(eigs.eigenvalues() - r.eigenvalues).max() <= 1e-10 * np.abs(r.eigenvalues).max()
    assert eigs.num_operations() == r.nops and eigs.num_iterations() == r.niter
    U = eigs.eigenvectors()
    assert np.abs(ref.to_scipy() @ U - U * eigs.eigenvalues()).max() <= 1e-9


def test_emu_default_gen_solver(emu):
    A = O.gen_sparse_data(100, 0.1)
    g = emu.GenEigsSolver(emu.SparseGenMatProd(A), 6, 20)
    g.init()
    g.compute(emu.SortRule.LargestMagn, 300)
    assert g.info() == emu.CompInfo.Successful
    ev, Z = g.eigenvalues(), g.eigenvectors()
    assert np.abs(A @ Z - Z * ev).max() <= 1e-9
    r = O.gen_eigs(O.Csr.from_scipy(A, "gen"), 6, 20, O.LargestMagn, 300)
    assert g.num_operations() == r.nops


# ---------------------------------------------------------------- sliced layout (SB200_SPMV_FORMAT=sell)
@pytest.mark.parametrize("threads", [256, 512, 1024])
@pytest.mark.parametrize("n,density", [(1, 1.0), (33, 0.3), (1025, 0.02), (5000, 0.004)])
def test_emu_sell_spmv(emu, threads, n, density):
    X.test_sell_spmv_matches_csr_and_scipy(emu, threads, n, density)


def test_emu_sell_padding_and_empty_rows(emu):
    X.test_sell_rejects_wasteful_padding(emu)
    X.test_sell_empty_rows_and_empty_matrix(emu)


def test_emu_sell_chunk_layout(emu):
    X.test_sell_with_column_blocks_and_chunk_layout(emu, 65_537, 3, 4)


@pytest.mark.parametrize("threads", [256, 1024])
def test_emu_sell_fused_step_factorization(emu, threads):
    X.test_sell_lanczos_factorization(emu, threads)


def test_emu_sell_persistent_grid(emu):
    # operands beyond 16384 windows per rank use a persistent grid that strides over the windows; forced here at a small size
    n = 13_000  # 13 windows > 12 resident CTAs of 256 threads on the 2 emulated SMs
    A = sp.random(n, n, density=0.0008, random_state=4, format="csr")
    x = np.random.default_rng(4).standard_normal(n)
    with X.env(SB200_SELL_PERSISTENT=1, SB200_SPMV_FORMAT="sell", SB200_SELL_THREADS=256, SB200_SELL_MAX_FILL=100):
        op = emu.SparseGenMatProd(A)
        assert op.spmv_layout()["format"] == "sell"
        assert np.abs(op.perform_op(x) - A @ x).max() <= 1e-13 * np.abs(A @ x).max()
        X.test_sell_lanczos_factorization(emu, 512)


def test_emu_sell_solver(emu):
    A = O.gen_sparse_data(300, 0.03)
    with X.env(SB200_SPMV_FORMAT="sell", SB200_SELL_MAX_FILL=100):
        ops = emu.SparseSymMatProd(A)
    assert ops.spmv_layout()["format"] == "sell"
    e = emu.SymEigsSolver(ops, 8, 24)
    e.init()
    assert e.compute(emu.SortRule.LargestAlge) == 8
    r = O.sym_eigs(O.Csr.from_scipy(A, "lower"), 8, 24, O.LargestAlge)
    assert np.abs(e.eigenvalues() - r.eigenvalues).max() <= 1e-10 * np.abs(r.eigenvalues).max()
    assert e.num_operations() == r.nops
    # a nonsymmetric operator through the sliced Arnoldi step head
    A = O.gen_sparse_data(100, 0.1)
    with X.env(SB200_SPMV_FORMAT="sell", SB200_SELL_MAX_FILL=100):
        op = emu.SparseGenMatProd(A)
    assert op.spmv_layout()["format"] == "sell"
    g = emu.GenEigsSolver(op, 6, 20)
    g.init()
    g.compute(emu.SortRule.LargestMagn, 300)
    assert g.info() == emu.CompInfo.Successful
    ev, Z = g.eigenvalues(), g.eigenvectors()
    assert np.abs(A @ Z - Z * ev).max() <= 1e-9


# ---------------------------------------------------------------- complex Hermitian path
@pytest.mark.parametrize("fmt,uplo", [("csc", "lower"), ("csr", "upper")])
def test_emu_herm_operator(emu, fmt, uplo):
    HC.operator_case(emu, 100, fmt, uplo)
    HC.operator_case(emu, 10, fmt, uplo)


def test_emu_herm_factorization(emu):
    HC.factorization_case(emu)




def test_emu_arnoldi_sweep_hands_rare_paths_back(emu):
    # Arnoldi sweep mode (solver_gen.cu): low-rank operators drive the residual to zero inside a sweep (f_zeroed -> FacCtl::abort -> host tail),
    # a full-rank one runs whole sweeps on the device; operation and iteration counts must be the oracle's in both regimes
    rng = np.random.default_rng(3)
    for n, r, k, m in ((30, 5, 3, 10), (40, 3, 2, 12), (25, 25, 4, 12)):
        M = rng.standard_normal((n, r)) @ rng.standard_normal((r, n)) if r < n else rng.standard_normal((n, n))
        A = sp.csc_matrix(M)
        g = emu.GenEigsSolver(emu.SparseGenMatProd(A), k, m)
        g.init()
        nconv = g.compute(emu.SortRule.LargestMagn, 300)
        ref = O.gen_eigs(O.Csr.from_scipy(A), k, m, O.LargestMagn, 300)
        assert nconv == ref.nconv and g.num_operations() == ref.nops and g.num_iterations() == ref.niter and int(g.info()) == ref.info
        if nconv:
            assert np.abs(np.sort_complex(g.eigenvalues()) - np.sort_complex(ref.eigenvalues)).max() <= 1e-9 * np.abs(ref.eigenvalues).max()


@pytest.mark.parametrize("fmt", ["sell"])  # the default layout; the opt-in overlap mode (SB200_OVERLAP=1, measured slower on the device) over the CSR-vector kernels ran in round 2
def test_emu_overlapped_sweep_column_blocks(emu, fmt):
    # natural single-rank layout with several column blocks: the sweep runs the correction pass in two row ranges and starts the head
    # blocks of the next operator application on a second stream in between (solver_sym.cu, overlap_capable); same history and
    # eigenvalues as the plain sweep and as the oracle
    A = O.gen_sparse_data(2600, 0.004)
    ref = O.sym_eigs(O.Csr.from_scipy(A, "lower"), 6, 20, O.LargestAlge)
    out = {}
    for name in ("overlap", "plain_sweep"):
        # the variants are distinguished by the layout the operator is created with: 0.005 MB slices -> 4 column blocks (overlapped sweep);
        # one block -> plain sweep
        with X.env(SB200_SPMV_FORMAT=fmt, SB200_SELL_MAX_FILL=100, SB200_XSLICE_MB=0.005 if name == "overlap" else 1000):
            op = emu.SparseSymMatProd(A)
        assert op.spmv_layout()["col_blocks"] == (4 if name == "overlap" else 1)
        e = emu.SymEigsSolver(op, 6, 20)
        e.init()
        with X.env(SB200_OVERLAP=1 if name == "overlap" else 0):
            assert e.compute(emu.SortRule.LargestAlge) == 6
        out[name] = (e.eigenvalues(), e.num_operations(), e.num_iterations(), e.stats()["host_syncs"])
    for name, (ev, nops, niter, syncs) in out.items():
        assert nops == ref.nops and niter == ref.niter, name
        assert np.abs(ev - ref.eigenvalues).max() <= 1e-10 * np.abs(ref.eigenvalues).max(), name
    assert np.abs(out["overlap"][0] - out["plain_sweep"][0]).max() <= 1e-12 * np.abs(ref.eigenvalues).max()


@pytest.mark.parametrize("selection", [O.LargestMagn, O.LargestAlge, O.SmallestAlge, O.BothEnds])
@pytest.mark.parametrize("n", [10, 100])
def test_emu_herm_solver(emu_order, n, selection):
    HC.solver_case(emu_order, n, selection)


def test_emu_herm_solver_smallest_magnitude(emu):
    HC.solver_case(emu, 10, O.SmallestMagn)


def test_emu_herm_argument_checks(emu):
    HC.argument_checks(emu)


def test_emu_herm_user_operator(emu):
    HC.user_operator_case(emu)


def test_emu_complex_arnoldi_factorization(emu_order):
    HC.complex_arnoldi_factorization_case(emu_order)


@pytest.mark.parametrize("m", [2, 3, 6, 20, 50, 63])
def test_emu_complex_dense_kernels(emu_order, m):
    HC.complex_dense_kernels_case(emu_order, m)


@pytest.mark.parametrize("rule", ["LargestMagn", "LargestReal", "LargestImag", "SmallestReal"])
def test_emu_complex_gen_solver(emu, rule):
    HC.complex_gen_solver_case(emu, 10, rule)


def test_emu_complex_gen_user_operator(emu):
    HC.complex_gen_user_operator_case(emu)


@pytest.mark.parametrize("rule", ["LargestMagn", "LargestReal", "LargestImag", "SmallestReal"])
def test_emu_complex_gen_solver_n100(emu_order, rule):
    HC.complex_gen_solver_case(emu_order, 100, rule)


def test_emu_complex_gen_solver_reverse_order(emu_order):
    HC.complex_gen_solver_case(emu_order, 10, "LargestMagn")


# ---------------------------------------------------------------- scheduling-order independence (race detection)
def test_emu_order_dense_kernels(emu_order):
    # the small dense device kernels in both fiber orders; m = 61, 62, 64 of the Hessenberg eigen-decomposition exercise the
    # overflow-rescaling branch whose missing __syncwarp() this check uncovered
    import test_gpu_gen as G
    import test_gpu_sym as S

    for m in (6, 50, 61, 62, 64):
        G.test_hessenberg_eigen_device(emu_order, m)
    for m in (6, 60):
        G.test_hessenberg_qr_device(emu_order, m)
        G.test_double_shift_qr_device(emu_order, m)
        S.test_tridiag_eigen_device(emu_order, m)
        S.test_tridiag_qr_device(emu_order, m)


def test_emu_order_solvers(emu_order):
    test_emu_default_operator_and_solver(emu_order)
    HC.solver_case(emu_order, 10, O.LargestAlge)
    X.test_sell_lanczos_factorization(emu_order, 512)


# ---------------------------------------------------------------- golden known-answer spectra through the kernels
@pytest.mark.parametrize("n", [10, 100])
def test_emu_against_golden_spectra(emu, n):
    import golden_cases as GC
    from oracle import herm as OH

    prob = {10: 0.5, 100: 0.1}[n]
    k, m = GC.KM[n]
    A = O.gen_sparse_data(n, prob)
    for rule, srule in ((O.LargestAlge, emu.SortRule.LargestAlge), (O.SmallestAlge, emu.SortRule.SmallestAlge))[:2 if n == 10 else 1]:
        e = emu.SymEigsSolver(emu.SparseSymMatProd(A), k, m)
        e.init()
        e.compute(srule)
        GC.check_sym_values("sym", n, rule, e.eigenvalues())
        h = emu.HermEigsSolver(emu.SparseHermMatProd(OH.gen_sparse_data_herm(n, prob)), k, m)
        h.init()
        h.compute(srule)
        GC.check_sym_values("herm", n, rule, h.eigenvalues())


# ---------------------------------------------------------------- row-sharded solver: ranks emulated as threads of one process
def _sharded_solve(emu, n, P, rp, ci, v, k, m, kind="sym", fmt_env=None):
    """One thread per rank; the emulated communicator (emu_build.py, comm.cu stand-in) rendezvouses the collectives in-process."""
    import threading

    from spectra_b200_emu import dist

    uid = emu.Comm.unique_id()
    out, errs = [None] * P, []

    def worker(rank):
        try:
            comm = emu.Comm(rank, P, uid)
            row0, nrows = dist.slab_range(n, rank, P)
            # global row pointers of the slab + the full col / value arrays (upload_csr_slab reads them from rowptr_local[0] on)
            op = emu.SparseGenMatProd.from_csr_slab(n, row0, rp[row0:row0 + nrows + 1], ci, v, comm=comm)
            if kind == "sym":
                e = emu.SymEigsSolver(op, k, m)
                e.init()
                nconv = e.compute(emu.SortRule.LargestAlge)
                out[rank] = dict(nconv=nconv, ev=e.eigenvalues(), nops=e.num_operations(), niter=e.num_iterations(), U=e.eigenvectors(local=True).copy(),
                                 Ufull=e.eigenvectors().copy(), row0=row0, nrows=nrows, layout=op.spmv_layout())
            else:
                # Arnoldi factorisation tier (test/Arnoldi.cpp) on the sharded operator: no restart, so the emulated run stays short
                g = emu.GenEigsSolver(op, k, m)
                g.init()
                g.factorize_from(1, m)
                fz = g.factorization()
                out[rank] = dict(V=fz["V"].copy(), H=fz["H"].copy(), f=fz["f"].copy(), beta=fz["beta"], nops=g.num_operations(), row0=row0, nrows=nrows)
        except BaseException as ex:  # noqa: BLE001
            import traceback

            errs.append((rank, traceback.format_exc()))

    with X.env(**(fmt_env or {})):
        ts = [threading.Thread(target=worker, args=(r,)) for r in range(P)]
        for t in ts:
            t.start()
        for t in ts:
            t.join(timeout=900)
    assert not errs, errs
    assert all(o is not None for o in out), "a rank did not finish (deadlocked collective?)"
    return out


@pytest.mark.parametrize("P,fmt,peer", [(2, "csr", 1), (3, "csr", 1), (2, "sell", 1), (3, "sell", 1), (2, "sell", 0), (3, "csr", 0)])
def test_emu_row_sharded_sym_solver(emu, P, fmt, peer):
    # 1-D row partition, all-gather of the SpMV operand in chunks, all-reduce of the dot products; every rank must run the
    # same iteration (identical operation counts), reproduce the single-rank eigenvalues, and hold its rows of the eigenvectors
    from spectra_b200_emu import synth

    n, k, m = 901, 5, 16
    rp, ci, v = synth.csr(n, 12, 3, True)
    A = sp.csr_matrix((v, ci, rp), shape=(n, n))
    # peer = 1: peer-memory mode (residual rows written into every rank's operand buffer by the correction pass, one-shot mailbox
    # all-reduce, peer.cu); peer = 0: the NCCL-style collectives (all-gather + all-reduce), the fallback when peers cannot be mapped
    fmt_env = dict(SB200_SPMV_FORMAT="sell", SB200_SELL_MAX_FILL=100) if fmt == "sell" else dict(SB200_SPMV_FORMAT="csr")
    fmt_env["SB200_PEER"] = str(peer)
    if P == 3 and peer:
        fmt_env["SB200_OVERLAP"] = "1"  # the opt-in two-part correction pass / early head blocks, in peer mode
    res = _sharded_solve(emu, n, P, rp, ci, v, k, m, "sym", fmt_env)
    ref = O.sym_eigs(O.Csr.adopt(n, rp, ci, v), k, m, O.LargestAlge, want_vectors=False)
    for r, o in enumerate(res):
        assert o["layout"]["format"] == fmt
        assert o["nconv"] == k and o["nops"] == res[0]["nops"] and o["niter"] == res[0]["niter"]
        assert np.array_equal(o["ev"], res[0]["ev"])  # bitwise identical across ranks (deterministic collectives + redundant restart kernel)
        assert np.abs(o["ev"] - ref.eigenvalues).max() <= 1e-10 * np.abs(ref.eigenvalues).max()
        assert np.array_equal(o["Ufull"][o["row0"]:o["row0"] + o["nrows"]], o["U"])
    U = res[0]["Ufull"]
    assert np.abs(A @ U - U * res[0]["ev"]).max() <= 1e-9
    assert res[0]["nops"] == ref.nops


def test_emu_row_sharded_arnoldi_factorization(emu):
    from spectra_b200_emu import synth

    n, k, m = 400, 3, 12
    rp, ci, v = synth.csr(n, 12, 4, False)
    A = sp.csr_matrix((v, ci, rp), shape=(n, n))
    res = _sharded_solve(emu, n, 2, rp, ci, v, k, m, "gen")
    res_nccl = _sharded_solve(emu, n, 2, rp, ci, v, k, m, "gen", dict(SB200_PEER="0"))
    assert np.array_equal(res[0]["H"], res_nccl[0]["H"])  # peer-memory mode and the collective path sum in the same (rank) order
    assert np.array_equal(res[0]["H"], res[1]["H"]) and res[0]["beta"] == res[1]["beta"] and res[0]["nops"] == res[1]["nops"]
    V = np.vstack([o["V"] for o in res])
    f = np.concatenate([o["f"] for o in res])
    H = res[0]["H"]
    E = A @ V - V @ H
    E[:, -1] -= f
    scale = max(1.0, np.abs(H).max())
    assert np.abs(E).max() <= 1e-12 * scale                      # A V - V H = f e_m'   (test/Arnoldi.cpp:60-75)
    assert np.abs(V.T @ V - np.eye(m)).max() <= 1e-12
    ref = O.factorize(O.Csr.adopt(n, rp, ci, v), m, kind="arnoldi")
    assert np.abs(H - ref["H"]).max() <= 1e-9 * scale


# ---------------------------------------------------------------- wide-band shift-solve route (round 2; GPU twins: tests/test_gpu_shift.py)
def test_emu_shift_solve_mesh_route(emu):
    # block-tridiagonal elimination with grid-wide block kernels on 9-point (2-D) and 27-point (3-D) stencil matrices, the split block
    # products, a singular shift, and a complete SymEigsShiftSolver solve against ARPACK
    import test_gpu_shift as S

    S.test_shift_solve_operator_mesh(emu, (7, 6, 5), True, "thomas")
    S.test_shift_solve_mesh_split_products(emu, (12, 9), 3)
    S.test_shift_solve_mesh_singular_shift(emu)
    # an operator stored in several column blocks (large n on a device; forced here): every CSR kernel of the route runs once per block
    with X.env(SB200_XSLICE_MB="0.002"):
        S.test_shift_solve_operator_mesh(emu, (40, 35), False, "thomas")
    S.test_shift_solve_mesh_inverse_variants(emu, (5, 8, 9), True, "blocked")
    S.test_shift_solve_mesh_inverse_variants(emu, (3, 2), True, "blocked")
    S.test_shift_solve_mesh_merged_blocks(emu, (40, 9), True, 30)
    S.test_sym_shift_eigs_mesh_vs_arpack(emu, (9, 8, 7), True, "thomas", 4, 14)

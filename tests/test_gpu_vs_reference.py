"""GPU results next to THE REFERENCE'S OWN CODE, live: oracle/_ref/libspectra_ref.so (yixuan/spectra's headers compiled in the development
container over oracle/eigen_standin, `make -C oracle ref`) travels to the GPU box with the snapshot, so the device path can be compared with
the reference itself and not only with the restatement (which tests/test_oracle_vs_reference.py shows to be bit-identical to it).
Tolerances are the ones the restatement-based GPU tests use.  Skipped when the library did not travel."""
import json
import os

import numpy as np
import pytest
import scipy.sparse as sp

import oracle as O
from oracle import ref as R
from helpers import readme_banded, stencil_matrix, sym_full

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not os.path.exists(R._LIB_PATH), reason="oracle/_ref/libspectra_ref.so is not in the snapshot")]

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SYM_CASES = [(10, 0.5, 3, 6), (100, 0.1, 10, 20), (1000, 0.01, 20, 50)]
GEN_CASES = [(10, 0.5, 3, 6), (100, 0.1, 10, 30), (1000, 0.01, 20, 50)]


def test_reference_library_loaded():
    assert R.version().startswith("spectra 1.2.0")


@pytest.mark.parametrize("n,prob,k,m", SYM_CASES)
@pytest.mark.parametrize("rule", ["LargestMagn", "LargestAlge", "SmallestMagn", "SmallestAlge", "BothEnds"])
def test_sym_eigs_vs_reference(gpu, n, prob, k, m, rule):
    # test/SymEigs.cpp:44-65,133-167: SymEigsSolver<SparseSymMatProd<double>> on the device and in the reference
    if n == 1000 and rule == "SmallestMagn":
        pytest.skip("~23k matvecs; interior selection is covered at n <= 100")
    A = O.gen_sparse_data(n, prob)
    eigs = gpu.SymEigsSolver(gpu.SparseSymMatProd(A), k, m)
    eigs.init()
    nconv = eigs.compute(gpu.SortRule[rule])
    ref = R.sym_eigs(R.Compressed.from_scipy(A), k, m, int(gpu.SortRule[rule]))
    assert ref.info == O.Successful and eigs.info() == gpu.CompInfo.Successful and nconv == ref.nconv == k
    evals, U = eigs.eigenvalues(), eigs.eigenvectors()
    assert np.abs(evals - ref.eigenvalues).max() <= 1e-10 * np.abs(ref.eigenvalues).max()
    assert abs(eigs.num_operations() - ref.nops) <= max(3 * m, ref.nops // 5)
    # eigenvectors: same subspace, column by column up to sign
    Af = sym_full(A)
    assert np.abs(Af @ U - U * evals).max() <= 1e-9
    if rule in ("LargestAlge", "SmallestAlge"):
        d = np.abs(np.sum(U * ref.eigenvectors, axis=0))
        assert np.abs(d - 1.0).max() <= 1e-8


@pytest.mark.parametrize("n,prob,k,m", GEN_CASES)
@pytest.mark.parametrize("rule", ["LargestMagn", "LargestReal", "SmallestReal"])
def test_gen_eigs_vs_reference(gpu, n, prob, k, m, rule):
    # test/GenEigs.cpp:38-107: GenEigsSolver<SparseGenMatProd<double>>
    A = O.gen_sparse_data(n, prob)
    eigs = gpu.GenEigsSolver(gpu.SparseGenMatProd(A), k, m)
    eigs.init()
    nconv = eigs.compute(gpu.SortRule[rule], 300)
    ref = R.gen_eigs(R.Compressed.from_scipy(A), k, m, int(gpu.SortRule[rule]), 300)
    assert ref.info == O.Successful and eigs.info() == gpu.CompInfo.Successful and nconv == ref.nconv == k
    evals = eigs.eigenvalues()
    a = np.sort_complex(np.round(evals, 9))
    b = np.sort_complex(np.round(ref.eigenvalues, 9))
    assert np.abs(np.sort_complex(evals) - np.sort_complex(ref.eigenvalues)).max() <= 1e-9 * np.abs(ref.eigenvalues).max() or np.allclose(a, b)


def test_readme_example_vs_reference(gpu):
    # README.md:146-178
    M = sp.csc_matrix(readme_banded(10))
    eigs = gpu.GenEigsSolver(gpu.SparseGenMatProd(M), 3, 6)
    eigs.init()
    eigs.compute(gpu.SortRule.LargestMagn)
    ref = R.gen_eigs(R.Compressed.from_scipy(M), 3, 6, O.LargestMagn)
    assert eigs.info() == gpu.CompInfo.Successful and ref.info == O.Successful
    assert np.abs(eigs.eigenvalues() - ref.eigenvalues).max() <= 1e-10 * np.abs(ref.eigenvalues).max()
    assert eigs.num_operations() == ref.nops and eigs.num_iterations() == ref.niter


@pytest.mark.parametrize("kind,n,m", [("lanczos", 1000, 50), ("lanczos", 5000, 64), ("arnoldi", 1000, 50)])
def test_factorization_vs_reference(gpu, kind, n, m):
    # test/Arnoldi.cpp flow: the device's H next to the reference's from the same start vector
    A = O.gen_sparse_data(n, 0.01) if n <= 1000 else sp.random(n, n, density=0.002, random_state=5, format="csc")
    v0 = O.simple_random(3, n)
    if kind == "lanczos":
        eigs = gpu.SymEigsSolver(gpu.SparseSymMatProd(A), 3, m)
    else:
        eigs = gpu.GenEigsSolver(gpu.SparseGenMatProd(A), 3, m)
    eigs.init(v0)
    eigs.factorize_from(1, m // 2)
    eigs.factorize_from(m // 2, m)
    H = eigs.factorization()["H"]
    _, Href, _, beta, _ = R.factorize(R.Compressed.from_scipy(sp.csc_matrix(A)), m, v0=v0, mid=m // 2, kind=kind)
    scale = max(1.0, np.abs(Href).max())
    assert np.abs(H - Href).max() <= 1e-9 * scale
    assert abs(eigs.factorization()["beta"] - beta) <= 1e-9 * max(1.0, beta)


@pytest.mark.parametrize("n", [10, 100, 1000])
def test_herm_eigs_vs_reference(gpu, n):
    # test/HermEigs.cpp:140-174: HermEigsSolver<SparseHermMatProd<std::complex<double>>>
    from oracle import herm as OH

    prob, k, m = {10: (0.5, 3, 6), 100: (0.1, 10, 20), 1000: (0.01, 20, 50)}[n]
    A = OH.gen_sparse_data_herm(n, prob)
    for rule in ("LargestAlge", "LargestMagn", "BothEnds"):
        eigs = gpu.HermEigsSolver(gpu.SparseHermMatProd(A), k, m)
        eigs.init()
        nconv = eigs.compute(gpu.SortRule[rule])
        ref = R.herm_eigs(R.CompressedZ(A), k, m, int(gpu.SortRule[rule]))
        assert ref.info == O.Successful and eigs.info() == gpu.CompInfo.Successful and nconv == ref.nconv == k
        assert np.abs(eigs.eigenvalues() - ref.eigenvalues).max() <= 1e-10 * np.abs(ref.eigenvalues).max()


@pytest.mark.parametrize("n", [10, 100])
def test_complex_gen_eigs_vs_reference(gpu, n):
    # test/ComplexEigs.cpp:151-192: GenEigsSolver<SparseGenMatProd<std::complex<double>>>
    from oracle import herm as OH

    prob, k, m = {10: (0.5, 3, 6), 100: (0.1, 10, 30)}[n]
    A = OH.gen_sparse_data_complex(n, prob)
    for rule in ("LargestMagn", "LargestReal"):
        g = gpu.GenEigsSolver(gpu.SparseHermMatProd(A, uplo="general"), k, m)
        g.init()
        nconv = g.compute(gpu.SortRule[rule], 300)
        ref = R.gen_eigs_complex(R.CompressedZG(A.tocsr()), k, m, int(gpu.SortRule[rule]), 300)
        assert ref.info == O.Successful and g.info() == gpu.CompInfo.Successful and nconv == ref.nconv == k
        assert np.abs(np.sort_complex(g.eigenvalues()) - np.sort_complex(ref.eigenvalues)).max() <= 1e-10 * np.abs(ref.eigenvalues).max()


@pytest.mark.parametrize("case", ["banded", "mesh", "fixture"])
def test_sym_shift_eigs_vs_reference(gpu, case):
    # SymEigsShiftSolver<SparseSymShiftSolve<double>> (BASELINE config 5's solver): the device against the reference's own driver
    # (set_shift, lambda = 1 / nu + sigma, sorting) over the stand-in's band LU -- banded (block cyclic reduction on the device), a 27-point
    # stencil, and the reference's random sparse fixture (test/SymEigsShift.cpp:148-186)
    if case == "banded":
        from spectra_b200 import synth

        n, sigma, k, m = 20_000, 0.5, 10, 30
        rp, ci, v = synth.band_csr(n, 15, 0, 0.0)
        A = sp.csr_matrix((v, ci, rp), shape=(n, n)).tocsc()
    elif case == "mesh":
        A, sigma, k, m = stencil_matrix((9, 8, 7), True, seed=7), 0.11, 4, 14
    else:
        A, sigma, k, m = sym_full(O.gen_sparse_data(100, 0.1)).tocsc(), 10.0, 10, 20
    op = gpu.SparseSymShiftSolve(sp.tril(A).tocsc())
    eigs = gpu.SymEigsShiftSolver(op, k, m, sigma)
    eigs.init()
    nconv = eigs.compute(gpu.SortRule.LargestMagn)
    ref = R.sym_shift_eigs(R.Compressed.from_scipy(sp.csc_matrix(A)), sigma, k, m, O.LargestMagn, want_vectors=False)
    assert ref.info == O.Successful and eigs.info() == gpu.CompInfo.Successful and nconv == ref.nconv == k
    evals = eigs.eigenvalues()
    assert np.abs(np.sort(evals) - np.sort(ref.eigenvalues)).max() <= 1e-10 * max(1.0, np.abs(ref.eigenvalues).max())
    assert abs(eigs.num_operations() - ref.nops) <= max(30, ref.nops // 5)


def test_c2_full_size_vs_reference_golden(gpu):
    # BASELINE config C2 against the reference's own complete solve of it (tests/golden/reference_C2.json, make_reference_golden.py)
    path = os.path.join(GOLDEN, "reference_C2.json")
    if not os.path.exists(path):
        pytest.skip("tests/golden/reference_C2.json not generated")
    if os.environ.get("SB200_TEST_BACKEND") == "emu":
        pytest.skip("n = 1e6 is a device-sized case")
    from spectra_b200 import synth

    g = json.load(open(path))
    n = g["n"]
    rp, ci, v = synth.csr(n, g["d"], g["seed"], True)
    eigs = gpu.SymEigsSolver(gpu.SparseGenMatProd.from_csr_slab(n, 0, rp, ci, v), g["nev"], g["ncv"])
    eigs.init()
    nconv = eigs.compute(gpu.SortRule.LargestAlge)
    ref = np.array(g["eigenvalues"])
    assert eigs.info() == gpu.CompInfo.Successful and nconv == g["nconv"]
    assert np.abs(eigs.eigenvalues() - ref).max() <= 1e-10 * np.abs(ref).max()
    assert abs(eigs.num_operations() - g["nops"]) <= 0.05 * g["nops"]

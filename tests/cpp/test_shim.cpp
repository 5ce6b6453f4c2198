// Drop-in check of the header-only C++ shim (include/Spectra/*): the call sequence of the reference's
// own tests (test/SymEigs.cpp:44-65,133-167 and test/GenEigs.cpp:38-71) against the GPU library.
// Eigen-free: the fixture generator below is the reference tests' gen_sparse_data (std::default_random_engine
// seeded 0), assembled into compressed ColMajor arrays by hand.
#include <Spectra/GenEigsSolver.h>
#include <Spectra/MatOp/SparseGenMatProd.h>
#include <Spectra/MatOp/SparseSymMatProd.h>
#include <Spectra/MatOp/SparseSymShiftSolve.h>
#include <Spectra/SymEigsShiftSolver.h>
#include <Spectra/SymEigsSolver.h>

#include <cmath>
#include <complex>
#include <cstdio>
#include <random>
#include <vector>

using namespace Spectra;

struct Csc
{
    int n;
    std::vector<int> outer, inner;
    std::vector<double> val;
};

static Csc gen_sparse_data(int n, double prob)
{
    std::default_random_engine gen;
    gen.seed(0);
    std::uniform_real_distribution<double> distr(0.0, 1.0);
    std::vector<std::vector<std::pair<int, double>>> cols(n);
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++)
            if (distr(gen) < prob)
                cols[j].push_back({i, distr(gen) - 0.5});
    Csc A;
    A.n = n;
    A.outer.push_back(0);
    for (int j = 0; j < n; j++)
    {
        for (auto& e : cols[j])
        {
            A.inner.push_back(e.first);
            A.val.push_back(e.second);
        }
        A.outer.push_back((int) A.inner.size());
    }
    return A;
}

// y = selfadjointView<Lower>(A) * x on the host (reference semantics)
static void sym_lower_mv(const Csc& A, const double* x, double* y)
{
    for (int i = 0; i < A.n; i++)
        y[i] = 0;
    for (int j = 0; j < A.n; j++)
        for (int p = A.outer[j]; p < A.outer[j + 1]; p++)
        {
            const int i = A.inner[p];
            if (i < j)
                continue;
            y[i] += A.val[p] * x[j];
            if (i != j)
                y[j] += A.val[p] * x[i];
        }
}

static int fails = 0;
#define REQUIRE(cond)                                              \
    do                                                             \
    {                                                              \
        if (!(cond))                                               \
        {                                                          \
            std::printf("REQUIRE failed: %s (line %d)\n", #cond, __LINE__); \
            fails++;                                               \
        }                                                          \
    } while (0)

static void run_sym(const Csc& A, int k, int m, SortRule rule)
{
    SparseSymMatProd<double> op(A.n, A.outer.data(), A.inner.data(), A.val.data());
    SymEigsSolver<SparseSymMatProd<double>> eigs(op, k, m);
    eigs.init();
    const int nconv = (int) eigs.compute(rule);
    REQUIRE(eigs.info() == CompInfo::Successful);
    REQUIRE(nconv == k);
    auto evals = eigs.eigenvalues();
    auto evecs = eigs.eigenvectors();
    REQUIRE(evecs.rows() == A.n && evecs.cols() == nconv);
    std::vector<double> y(A.n);
    double err = 0;
    for (int c = 0; c < nconv; c++)
    {
        sym_lower_mv(A, &evecs(0, c), y.data());
        for (int i = 0; i < A.n; i++)
            err = std::max(err, std::fabs(y[i] - evecs(i, c) * evals[c]));
    }
    std::printf("sym n=%d rule=%d nconv=%d niter=%d nops=%d ||AU-UD||_inf=%.3e\n", A.n, (int) rule, nconv, (int) eigs.num_iterations(),
                (int) eigs.num_operations(), err);
    REQUIRE(err <= 1e-9);
    // operator tier: op * M and op(i, j)
    double x0 = op(45 % A.n, 22 % A.n);
    (void) x0;
    // template-argument variants (SparseSymMatProd.h:30): the same compressed arrays read as RowMajor hold the transpose, whose
    // Upper triangle is this matrix' Lower one; 64-bit StorageIndex.  Same symmetric matrix => same eigenvalues and op counts.
    std::vector<long long> outer64(A.outer.begin(), A.outer.end()), inner64(A.inner.begin(), A.inner.end());
    using OpV = SparseSymMatProd<double, SPECTRA_B200_UPPER, SPECTRA_B200_ROWMAJOR, long long>;
    OpV opv(A.n, outer64.data(), inner64.data(), A.val.data());
    SymEigsSolver<OpV> eigv(opv, k, m);
    eigv.init();
    eigv.compute(rule);
    REQUIRE(eigv.info() == CompInfo::Successful);
    REQUIRE(eigv.num_operations() == eigs.num_operations());
    auto ev2 = eigv.eigenvalues();
    for (int c = 0; c < nconv; c++)
        REQUIRE(ev2[c] == evals[c]);
    REQUIRE(opv(22 % A.n, 45 % A.n) == x0);
}

static void run_gen(const Csc& A, int k, int m, SortRule rule)
{
    SparseGenMatProd<double> op(A.n, A.outer.data(), A.inner.data(), A.val.data());
    GenEigsSolver<SparseGenMatProd<double>> eigs(op, k, m);
    eigs.init();
    const int nconv = (int) eigs.compute(rule, 300);
    REQUIRE(eigs.info() == CompInfo::Successful);
    auto evals = eigs.eigenvalues();
    auto evecs = eigs.eigenvectors();
    double err = 0;
    for (int c = 0; c < nconv; c++)
    {
        std::vector<std::complex<double>> y(A.n, 0.0);
        for (int j = 0; j < A.n; j++)
            for (int p = A.outer[j]; p < A.outer[j + 1]; p++)
                y[A.inner[p]] += A.val[p] * evecs(j, c);
        for (int i = 0; i < A.n; i++)
            err = std::max(err, std::abs(y[i] - evecs(i, c) * evals[c]));
    }
    std::printf("gen n=%d rule=%d nconv=%d nops=%d ||AU-UD||_inf=%.3e\n", A.n, (int) rule, nconv, (int) eigs.num_operations(), err);
    REQUIRE(err <= 1e-9);
}

// The user-defined operator of the reference's documentation (SymEigsSolver.h:99-126): M = diag(1, 2, ..., 10)
class MyDiagonalTen
{
public:
    using Scalar = double;
    Index rows() const { return 10; }
    Index cols() const { return 10; }
    void perform_op(const double* x_in, double* y_out) const
    {
        for (Index i = 0; i < rows(); i++)
            y_out[i] = x_in[i] * (i + 1);
    }
};

// The user-defined shift-solve operator of the reference's documentation (SymEigsShiftSolver.h:104-146)
class MyDiagonalTenShiftSolve
{
    double sigma_ = 0.0;

public:
    using Scalar = double;
    Index rows() const { return 10; }
    Index cols() const { return 10; }
    void set_shift(double sigma) { sigma_ = sigma; }
    void perform_op(const double* x_in, double* y_out) const
    {
        for (Index i = 0; i < rows(); i++)
            y_out[i] = x_in[i] / (i + 1 - sigma_);
    }
};

// symmetric band matrix (half-bandwidth b) in compressed ColMajor, lower triangle only; a(i,j) = cos-hash, strong diagonal i
static Csc band_lower(int n, int b)
{
    Csc A;
    A.n = n;
    A.outer.push_back(0);
    for (int j = 0; j < n; j++)
    {
        for (int i = j; i < n && i <= j + b; i++)
        {
            A.inner.push_back(i);
            A.val.push_back(i == j ? 0.01 * (j + 1) : 0.3 * std::cos(1.7 * i + 0.3 * j) / (1 + i - j));
        }
        A.outer.push_back((int) A.inner.size());
    }
    return A;
}

static void run_shift(int n, int b, int k, int m, double sigma)
{
    Csc A = band_lower(n, b);
    SparseSymShiftSolve<double> op(A.n, A.outer.data(), A.inner.data(), A.val.data());
    SymEigsShiftSolver<SparseSymShiftSolve<double>> eigs(op, k, m, sigma);
    eigs.init();
    const Index nconv = eigs.compute(SortRule::LargestMagn);
    REQUIRE(eigs.info() == CompInfo::Successful);
    REQUIRE(nconv == k);
    auto ev = eigs.eigenvalues();
    auto U = eigs.eigenvectors();
    // test/SymEigsShift.cpp:72-76: ||A U - U D||_inf <= 1e-9
    double err = 0.0;
    std::vector<double> y(n);
    for (Index c = 0; c < U.cols(); c++)
    {
        sym_lower_mv(A, &U(0, c), y.data());
        for (int i = 0; i < n; i++)
            err = std::max(err, std::fabs(y[i] - ev[c] * U(i, c)));
    }
    REQUIRE(err <= 1e-9);
    // the solve itself: (A - sigma I) y = x
    std::vector<double> x(n), z(n), t(n);
    for (int i = 0; i < n; i++)
        x[i] = std::sin(0.1 * i) + 0.5;
    op.perform_op(x.data(), z.data());
    sym_lower_mv(A, z.data(), t.data());
    double rs = 0.0, xs = 0.0;
    for (int i = 0; i < n; i++)
    {
        const double r = t[i] - sigma * z[i] - x[i];
        rs += r * r;
        xs += x[i] * x[i];
    }
    REQUIRE(std::sqrt(rs / xs) <= 1e-11);
    std::printf("shift-invert n=%d b=%d sigma=%g: nconv=%d nops=%d ev0=%.12f residual=%.2e solve=%.2e\n", n, b, sigma, (int) nconv, (int) eigs.num_operations(),
                ev[0], err, std::sqrt(rs / xs));
}

int main()
{
    {
        MyDiagonalTenShiftSolve op;
        SymEigsShiftSolver<MyDiagonalTenShiftSolve> eigs(op, 3, 6, 3.14);
        eigs.init();
        eigs.compute(SortRule::LargestMagn);
        REQUIRE(eigs.info() == CompInfo::Successful);
        auto ev = eigs.eigenvalues();
        REQUIRE(ev.size() == 3);
        REQUIRE(std::fabs(ev[0] - 4.0) < 1e-10 && std::fabs(ev[1] - 3.0) < 1e-10 && std::fabs(ev[2] - 2.0) < 1e-10);
        std::printf("user shift-solve op diag(1..10), sigma 3.14: %.12f %.12f %.12f\n", ev[0], ev[1], ev[2]);
    }
#ifdef SB200_SHIM_TEST_SMALL
    run_shift(120, 3, 4, 12, 1.0);
#else
    run_shift(500, 3, 5, 20, 1.0);
#endif
#ifndef SB200_SHIM_TEST_SMALL  // the emulated CPU run (tests/test_cpp_shim.py) keeps the small cases only
    run_shift(20000, 15, 10, 30, 100.005);
#endif
#ifndef SB200_SHIM_TEST_SMALL
    {
        // a large pattern that is not banded is rejected with the reference's exception type
        Csc A = gen_sparse_data(3000, 0.002);
        bool thrown = false;
        try
        {
            SparseSymShiftSolve<double> bad(A.n, A.outer.data(), A.inner.data(), A.val.data());
        }
        catch (const std::invalid_argument&)
        {
            thrown = true;
        }
        REQUIRE(thrown);
    }
#endif
    {
        MyDiagonalTen op;
        SymEigsSolver<MyDiagonalTen> eigs(op, 3, 6);
        eigs.init();
        eigs.compute(SortRule::LargestAlge);
        REQUIRE(eigs.info() == CompInfo::Successful);
        auto ev = eigs.eigenvalues();
        REQUIRE(ev.size() == 3);
        REQUIRE(std::fabs(ev[0] - 10.0) < 1e-10 && std::fabs(ev[1] - 9.0) < 1e-10 && std::fabs(ev[2] - 8.0) < 1e-10);
        std::printf("user op diag(1..10): %.12f %.12f %.12f nops=%d\n", ev[0], ev[1], ev[2], (int) eigs.num_operations());
    }
    const struct
    {
        int n;
        double p;
        int k, m, mg;
    } cases[] = {{10, 0.5, 3, 6, 6}, {100, 0.1, 10, 20, 30}, {1000, 0.01, 20, 50, 50}};
    for (auto& c : cases)
    {
#ifdef SB200_SHIM_TEST_SMALL
        if (c.n > 100)
            continue;
#endif
        Csc A = gen_sparse_data(c.n, c.p);
        for (SortRule r : {SortRule::LargestMagn, SortRule::LargestAlge, SortRule::SmallestAlge, SortRule::BothEnds})
            run_sym(A, c.k, c.m, r);
        for (SortRule r : {SortRule::LargestMagn, SortRule::LargestReal, SortRule::LargestImag, SortRule::SmallestReal})
            run_gen(A, c.k, c.mg, r);
    }
    // exceptions keep the reference's types
    {
        Csc A = gen_sparse_data(10, 0.5);
        SparseSymMatProd<double> op(A.n, A.outer.data(), A.inner.data(), A.val.data());
        bool thrown = false;
        try
        {
            SymEigsSolver<SparseSymMatProd<double>> bad(op, 10, 12);
        }
        catch (const std::invalid_argument&)
        {
            thrown = true;
        }
        REQUIRE(thrown);
        SymEigsSolver<SparseSymMatProd<double>> eigs(op, 3, 6);
        std::vector<double> zero(10, 0.0);
        thrown = false;
        try
        {
            eigs.init(zero.data());
        }
        catch (const std::invalid_argument&)
        {
            thrown = true;
        }
        REQUIRE(thrown);
    }
    std::printf(fails ? "FAILED (%d)\n" : "ALL PASSED\n", fails);
    return fails ? 1 : 0;
}

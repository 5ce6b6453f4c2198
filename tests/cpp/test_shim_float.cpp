// Scalar = float through the shim (the reference's wrappers are templated on Scalar, SparseSymMatProd.h:30, SparseGenMatProd.h:29):
// float matrices / vectors at the boundary, fp64 arithmetic on the device.  Flow of test/SymEigs.cpp / test/GenEigs.cpp on the
// reference's gen_sparse_data fixture (n = 100), plus a user-defined float operator (SymEigsSolver.h:99-126).
#include <Spectra/GenEigsSolver.h>
#include <Spectra/MatOp/DenseGenMatProd.h>
#include <Spectra/MatOp/DenseSymMatProd.h>
#include <Spectra/MatOp/SparseGenMatProd.h>
#include <Spectra/MatOp/SparseSymMatProd.h>
#include <Spectra/SymEigsSolver.h>

#include <cmath>
#include <complex>
#include <cstdio>
#include <random>
#include <type_traits>
#include <vector>

using namespace Spectra;

struct Csc
{
    int n;
    std::vector<int> outer, inner;
    std::vector<float> val;
};

// test/SymEigs.cpp:25-42 (values narrowed to float)
static Csc gen_sparse_data(int n, double prob)
{
    std::vector<std::vector<std::pair<int, float>>> cols(n);
    std::default_random_engine gen;
    gen.seed(0);
    std::uniform_real_distribution<double> distr(0.0, 1.0);
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++)
            if (distr(gen) < prob)
                cols[j].push_back({i, (float) (distr(gen) - 0.5)});
    Csc A;
    A.n = n;
    A.outer.assign(n + 1, 0);
    for (int j = 0; j < n; j++)
    {
        A.outer[j + 1] = A.outer[j] + (int) cols[j].size();
        for (auto& e : cols[j])
        {
            A.inner.push_back(e.first);
            A.val.push_back(e.second);
        }
    }
    return A;
}

static int failures = 0;
#define CHECK(cond)                                                        \
    do                                                                     \
    {                                                                      \
        if (!(cond))                                                       \
        {                                                                  \
            std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
            failures++;                                                    \
        }                                                                  \
    } while (0)

// y = selfadjointView<Lower>(A) x  /  y = A x, in double on the host
static void sym_matvec(const Csc& A, const float* x, double* y)
{
    for (int i = 0; i < A.n; i++)
        y[i] = 0;
    for (int j = 0; j < A.n; j++)
        for (int p = A.outer[j]; p < A.outer[j + 1]; p++)
        {
            const int i = A.inner[p];
            if (i < j)
                continue;
            y[i] += (double) A.val[p] * x[j];
            if (i != j)
                y[j] += (double) A.val[p] * x[i];
        }
}

// user-defined float operator: a diagonal matrix 1..n (SymEigsSolver.h:99-126)
struct DiagOp
{
    using Scalar = float;
    int n;
    int rows() const { return n; }
    int cols() const { return n; }
    void perform_op(const float* x_in, float* y_out) const
    {
        for (int i = 0; i < n; i++)
            y_out[i] = x_in[i] * (float) (i + 1);
    }
};

int main()
{
    const int n = 100;
    const Csc A = gen_sparse_data(n, 0.1);
    {
        SparseSymMatProd<float> op(n, A.outer.data(), A.inner.data(), A.val.data());
        static_assert(std::is_same<SparseSymMatProd<float>::Scalar, float>::value, "Scalar");
        std::vector<float> x(n), y(n);
        std::vector<double> y0(n);
        for (int i = 0; i < n; i++)
            x[i] = (float) std::sin(i + 1.0);
        op.perform_op(x.data(), y.data());
        sym_matvec(A, x.data(), y0.data());
        double err = 0;
        for (int i = 0; i < n; i++)
            err = std::max(err, std::fabs(y[i] - y0[i]));
        CHECK(err <= 1e-5);
        SymEigsSolver<SparseSymMatProd<float>> eigs(op, 10, 20);
        eigs.init();
        const Index nconv = eigs.compute(SortRule::LargestAlge, 1000, 1e-6f);
        CHECK(eigs.info() == CompInfo::Successful && nconv == 10);
        const auto evals = eigs.eigenvalues();
        const auto evecs = eigs.eigenvectors();
        static_assert(std::is_same<std::decay<decltype(evals[0])>::type, float>::value, "float eigenvalues");
        static_assert(std::is_same<std::decay<decltype(evecs(0, 0))>::type, float>::value, "float eigenvectors");
        CHECK(evals.size() == 10 && evecs.rows() == n && evecs.cols() == 10);
        double res = 0;
        for (Index c = 0; c < evecs.cols(); c++)
        {
            sym_matvec(A, evecs.data() + c * n, y0.data());
            for (int i = 0; i < n; i++)
                res = std::max(res, std::fabs(y0[i] - (double) evecs(i, c) * evals[c]));
        }
        std::printf("sym float: nconv=%d nops=%d ||AU-UD||_inf=%.3e\n", (int) nconv, (int) eigs.num_operations(), res);
        CHECK(res <= 1e-5);  // float storage of U and D
        for (Index c = 1; c < evals.size(); c++)
            CHECK(evals[c - 1] >= evals[c]);
    }
    {
        SparseGenMatProd<float> op(n, A.outer.data(), A.inner.data(), A.val.data());
        GenEigsSolver<SparseGenMatProd<float>> eigs(op, 6, 20);
        eigs.init();
        const Index nconv = eigs.compute(SortRule::LargestMagn, 300, 1e-6f);
        CHECK(eigs.info() == CompInfo::Successful && nconv == 6);
        const auto evals = eigs.eigenvalues();
        const auto evecs = eigs.eigenvectors();
        static_assert(std::is_same<std::decay<decltype(evals[0])>::type, std::complex<float>>::value, "complex<float> eigenvalues");
        double res = 0;
        for (Index c = 0; c < evecs.cols(); c++)
        {
            std::vector<std::complex<double>> y(n, 0.0);
            for (int j = 0; j < n; j++)
                for (int p = A.outer[j]; p < A.outer[j + 1]; p++)
                    y[A.inner[p]] += (double) A.val[p] * std::complex<double>(evecs(j, c));
            for (int i = 0; i < n; i++)
                res = std::max(res, std::abs(y[i] - std::complex<double>(evecs(i, c)) * std::complex<double>(evals[c])));
        }
        std::printf("gen float: nconv=%d nops=%d ||AU-UD||_inf=%.3e\n", (int) nconv, (int) eigs.num_operations(), res);
        CHECK(res <= 1e-5);
    }
    {
        DiagOp op{10};
        SymEigsSolver<DiagOp> eigs(op, 3, 6);
        eigs.init();
        const Index nconv = eigs.compute(SortRule::LargestAlge, 1000, 1e-6f);
        CHECK(eigs.info() == CompInfo::Successful && nconv == 3);
        const auto evals = eigs.eigenvalues();
        CHECK(std::fabs(evals[0] - 10.f) < 1e-4f && std::fabs(evals[1] - 9.f) < 1e-4f && std::fabs(evals[2] - 8.f) < 1e-4f);  // SymEigsSolver.h:99-126
    }
    {
        // the reference's first README example (README.md:90-125): dense symmetric M = A + A', DenseSymMatProd, 3 largest eigenvalues
        const int nd = 10;
        std::vector<double> M(nd * nd);
        std::default_random_engine gen(7);
        std::uniform_real_distribution<double> distr(-1.0, 1.0);
        std::vector<double> Ar(nd * nd);
        for (double& a : Ar)
            a = distr(gen);
        for (int i = 0; i < nd; i++)
            for (int j = 0; j < nd; j++)
                M[i + j * nd] = Ar[i + j * nd] + Ar[j + i * nd];
        DenseSymMatProd<double> op(nd, M.data());
        SymEigsSolver<DenseSymMatProd<double>> eigs(op, 3, 6);
        eigs.init();
        const Index nconv = eigs.compute(SortRule::LargestAlge);
        CHECK(eigs.info() == CompInfo::Successful && nconv == 3);
        const auto evals = eigs.eigenvalues();
        const auto evecs = eigs.eigenvectors();
        double res = 0;
        for (Index c = 0; c < 3; c++)
            for (int i = 0; i < nd; i++)
            {
                double y = 0;
                for (int j = 0; j < nd; j++)
                    y += M[i + j * nd] * evecs(j, c);
                res = std::max(res, std::fabs(y - evecs(i, c) * evals[c]));
            }
        std::printf("dense sym README example: %.10f %.10f %.10f ||AU-UD||_inf=%.3e\n", evals[0], evals[1], evals[2], res);
        CHECK(res <= 1e-9);
        CHECK(std::fabs(op(2, 5) - M[2 + 5 * nd]) == 0.0);
        DenseGenMatProd<double> gop(nd, Ar.data());
        GenEigsSolver<DenseGenMatProd<double>> geigs(gop, 2, 6);
        geigs.init();
        geigs.compute(SortRule::LargestMagn, 300);
        CHECK(geigs.info() == CompInfo::Successful);
    }
    std::printf(failures ? "FAILED (%d)\n" : "ALL PASSED\n", failures);
    return failures ? 1 : 0;
}

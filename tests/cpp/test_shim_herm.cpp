// The flow of the reference's test/HermEigs.cpp (sparse cases, :27-71, :140-163) written against the B200 shim headers:
// the same gen_sparse_data (std::default_random_engine seeded 0), SparseHermMatProd + HermEigsSolver, every selection rule,
// acceptance ||AU - UD||_inf <= 1e-9.  No Eigen: the fixture is assembled into CSC arrays by hand.
#include <Spectra/GenEigsSolver.h>
#include <Spectra/HermEigsSolver.h>
#include <Spectra/MatOp/SparseGenMatProd.h>
#include <Spectra/MatOp/SparseHermMatProd.h>

#include <cmath>
#include <complex>
#include <cstdio>
#include <random>
#include <vector>

using namespace Spectra;
using cd = std::complex<double>;

struct Csc
{
    int n;
    std::vector<int> outer, inner;
    std::vector<cd> val;
};

static Csc gen_sparse_data(int n, double prob)
{
    std::vector<std::vector<std::pair<int, cd>>> cols(n);
    std::default_random_engine gen;
    gen.seed(0);
    std::uniform_real_distribution<double> distr(0.0, 1.0);
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++)
            if (distr(gen) < prob)
            {
                const double re = distr(gen) - 0.5;
                const double im = (i == j) ? 0.0 : (distr(gen) - 0.5);
                cols[j].push_back({i, cd(re, im)});  // rows arrive in ascending order
            }
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

// y = selfadjointView<Lower>(A) * x on the host
static void herm_matvec(const Csc& A, const cd* x, cd* y)
{
    for (int i = 0; i < A.n; i++)
        y[i] = 0;
    for (int j = 0; j < A.n; j++)
        for (int p = A.outer[j]; p < A.outer[j + 1]; p++)
        {
            const int i = A.inner[p];
            if (i < j)
                continue;
            if (i == j)
                y[i] += A.val[p].real() * x[j];
            else
            {
                y[i] += A.val[p] * x[j];
                y[j] += std::conj(A.val[p]) * x[i];
            }
        }
}

// test/ComplexEigs.cpp:20-39: general complex sparse matrix (re and im drawn for every entry)
static Csc gen_sparse_data_complex(int n, double prob)
{
    std::vector<std::vector<std::pair<int, cd>>> cols(n);
    std::default_random_engine gen;
    gen.seed(0);
    std::uniform_real_distribution<double> distr(0.0, 1.0);
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++)
            if (distr(gen) < prob)
            {
                const double re = distr(gen) - 0.5;
                const double im = distr(gen) - 0.5;
                cols[j].push_back({i, cd(re, im)});
            }
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
#define CHECK(cond)                                                  \
    do                                                               \
    {                                                                \
        if (!(cond))                                                 \
        {                                                            \
            std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
            failures++;                                              \
        }                                                            \
    } while (0)

static void run_case(int n, double prob, int k, int m)
{
    const Csc A = gen_sparse_data(n, prob);
    SparseHermMatProd<cd> op(n, A.outer.data(), A.inner.data(), A.val.data());
    CHECK(op.rows() == n && op.cols() == n);
    const SortRule rules[] = {SortRule::LargestMagn, SortRule::LargestAlge, SortRule::SmallestMagn, SortRule::SmallestAlge, SortRule::BothEnds};
    for (SortRule rule : rules)
    {
        if (n >= 100 && rule == SortRule::SmallestMagn)
            continue;  // converges after > 1000 matrix operations
        HermEigsSolver<SparseHermMatProd<cd>> eigs(op, k, m);
        eigs.init();
        const Index nconv = eigs.compute(rule);
        CHECK(eigs.info() == CompInfo::Successful);
        CHECK(nconv == k);
        const auto evals = eigs.eigenvalues();
        const auto evecs = eigs.eigenvectors();
        CHECK(evals.size() == k && evecs.rows() == n && evecs.cols() == k);
        double err = 0.0;
        std::vector<cd> y(n);
        for (Index c = 0; c < evecs.cols(); c++)
        {
            herm_matvec(A, evecs.data() + c * n, y.data());
            for (int i = 0; i < n; i++)
                err = std::max(err, std::abs(y[i] - evecs(i, c) * evals[c]));
        }
        std::printf("n=%d rule=%d nconv=%d niter=%d nops=%d ||AU-UD||_inf=%.3e\n", n, (int) rule, (int) nconv, (int) eigs.num_iterations(), (int) eigs.num_operations(), err);
        CHECK(err <= 1e-9);
    }
}

// user-defined complex operator (the OpType concept, SymEigsSolver.h:99-114, with Scalar = std::complex<double>):
// A = D + u u^H with D = diag(1..n) real and a fixed complex u -- Hermitian, applied without forming the matrix
struct RankOneUpdateOp
{
    using Scalar = cd;
    int n;
    std::vector<cd> u;
    explicit RankOneUpdateOp(int n_) : n(n_), u(n_)
    {
        for (int i = 0; i < n; i++)
            u[i] = cd(std::cos(0.7 * i), std::sin(1.3 * i)) / std::sqrt((double) n);
    }
    int rows() const { return n; }
    int cols() const { return n; }
    void perform_op(const cd* x, cd* y) const
    {
        cd dot = 0;
        for (int i = 0; i < n; i++)
            dot += std::conj(u[i]) * x[i];
        for (int i = 0; i < n; i++)
            y[i] = (double) (i + 1) * x[i] + u[i] * dot;
    }
};

// the flow of test/ComplexEigs.cpp:41-110 (sparse case): GenEigsSolver<SparseGenMatProd<std::complex<double>>>, maxit = 300
static void run_complex_gen(int n, double prob, int k, int m)
{
    const Csc A = gen_sparse_data_complex(n, prob);
    SparseGenMatProd<cd> op(n, A.outer.data(), A.inner.data(), A.val.data());
    for (SortRule rule : {SortRule::LargestMagn, SortRule::LargestReal, SortRule::SmallestReal})
    {
        GenEigsSolver<SparseGenMatProd<cd>> eigs(op, k, m);
        eigs.init();
        const Index nconv = eigs.compute(rule, 300);
        CHECK(eigs.info() == CompInfo::Successful && nconv == k);
        const auto evals = eigs.eigenvalues();
        const auto evecs = eigs.eigenvectors();
        double err = 0.0;
        for (Index c = 0; c < evecs.cols(); c++)
        {
            std::vector<cd> y(n, cd(0));
            for (int j = 0; j < n; j++)
                for (int p = A.outer[j]; p < A.outer[j + 1]; p++)
                    y[A.inner[p]] += A.val[p] * evecs(j, c);
            for (int i = 0; i < n; i++)
                err = std::max(err, std::abs(y[i] - evecs(i, c) * evals[c]));
        }
        std::printf("complex gen n=%d rule=%d nconv=%d nops=%d ||AU-UD||_inf=%.3e\n", n, (int) rule, (int) nconv, (int) eigs.num_operations(), err);
        CHECK(err <= 1e-9);
    }
}

int main()
{
    run_complex_gen(10, 0.5, 3, 6);
    run_complex_gen(100, 0.1, 10, 30);
    {
        RankOneUpdateOp op(40);
        HermEigsSolver<RankOneUpdateOp> eigs(op, 4, 12);
        eigs.init();
        const Index nconv = eigs.compute(SortRule::LargestAlge);
        CHECK(eigs.info() == CompInfo::Successful && nconv == 4);
        const auto evals = eigs.eigenvalues();
        const auto evecs = eigs.eigenvectors();
        double err = 0.0;
        std::vector<cd> y(40);
        for (Index c = 0; c < evecs.cols(); c++)
        {
            op.perform_op(evecs.data() + c * 40, y.data());
            for (int i = 0; i < 40; i++)
                err = std::max(err, std::abs(y[i] - evecs(i, c) * evals[c]));
        }
        std::printf("user complex op D + uu^H: lambda_max=%.12f ||AU-UD||_inf=%.3e\n", evals[0], err);
        CHECK(err <= 1e-9);
        CHECK(evals[0] > 40.0 && evals[0] < 41.0 + 1e-9);  // interlacing: lambda_max(D) <= lambda_max(D + uu^H) <= lambda_max(D) + |u|^2
    }
    run_case(10, 0.5, 3, 6);
    run_case(100, 0.1, 10, 20);
    bool threw = false;
    try
    {
        const Csc A = gen_sparse_data(10, 0.5);
        SparseHermMatProd<cd> op(10, A.outer.data(), A.inner.data(), A.val.data());
        HermEigsSolver<SparseHermMatProd<cd>> eigs(op, 10, 11);  // nev > n - 1 (HermEigsBase.h:267-271)
    }
    catch (const std::invalid_argument&)
    {
        threw = true;
    }
    CHECK(threw);
    std::printf(failures ? "FAILED (%d)\n" : "ALL PASSED\n", failures);
    return failures ? 1 : 0;
}

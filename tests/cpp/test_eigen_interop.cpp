// One user program, two libraries.  This file uses ONLY the reference's public API with Eigen types, the way the reference's README and
// tests do (README.md:146-178 sparse example, test/SymEigs.cpp, test/GenEigs.cpp, test/HermEigs.cpp flows).  It is compiled twice:
//   g++ -I oracle/eigen_standin -I /root/reference/include ...                      -> the reference (header-only)
//   g++ -I oracle/eigen_standin -I include ... -lspectra_b200 (or the emulator .so) -> this repository's drop-in shim
// and the two binaries must print the same results (tests/test_cpp_shim.py compares them line by line; where /root/reference is absent the
// reference's output is the committed tests/golden/eigen_interop_reference.txt).  <Eigen/...> is the stand-in of oracle/eigen_standin
// (Eigen 3.4 is not installed in this image); with the real Eigen on the include path the same file compiles unchanged.
#include <Eigen/Core>
#include <Eigen/SparseCore>
#include <Spectra/SymEigsSolver.h>
#include <Spectra/GenEigsSolver.h>
#include <Spectra/HermEigsSolver.h>
#include <Spectra/MatOp/SparseSymMatProd.h>
#include <Spectra/MatOp/SparseGenMatProd.h>
#include <Spectra/MatOp/SparseHermMatProd.h>

#include <complex>
#include <cstdio>
#include <stdexcept>
#include <vector>

using namespace Spectra;

// deterministic entries (a plain LCG: no dependence on the C library's rand())
struct Lcg
{
    unsigned long long s;
    explicit Lcg(unsigned long long seed) : s(seed) {}
    double uniform()  // (-0.5, 0.5)
    {
        s = s * 6364136223846793005ULL + 1442695040888963407ULL;
        return double(s >> 11) / 9007199254740992.0 - 0.5;
    }
};

template <typename Vec>
static void print_vec(const char* key, const Vec& v)
{
    std::printf("%s:", key);
    for (Eigen::Index i = 0; i < v.size(); i++)
        std::printf(" %.17g", double(v[i]));
    std::printf("\n");
}
static void print_cvec(const char* key, const Eigen::VectorXcd& v)
{
    std::printf("%s:", key);
    for (Eigen::Index i = 0; i < v.size(); i++)
        std::printf(" %.17g %.17g", v[i].real(), v[i].imag());
    std::printf("\n");
}

// README.md:146-178 -- note that M is left in UNCOMPRESSED mode (reserve + insert, no makeCompressed())
static void readme_sparse_general()
{
    const int n = 10;
    Eigen::SparseMatrix<double> M(n, n);
    M.reserve(Eigen::VectorXi::Constant(n, 3));
    for (int i = 0; i < n; i++)
    {
        M.insert(i, i) = 1.0;
        if (i > 0)
            M.insert(i - 1, i) = 3.0;
        if (i < n - 1)
            M.insert(i + 1, i) = 2.0;
    }
    SparseGenMatProd<double> op(M);
    GenEigsSolver<SparseGenMatProd<double>> eigs(op, 3, 6);
    eigs.init();
    int nconv = int(eigs.compute(SortRule::LargestMagn));
    std::printf("readme.info: %d\nreadme.nconv: %d\n", int(eigs.info()), nconv);
    Eigen::VectorXcd evalues;
    if (eigs.info() == CompInfo::Successful)
        evalues = eigs.eigenvalues();
    print_cvec("readme.evalues", evalues);
    std::printf("readme.coeff: %.17g %.17g %.17g\n", op(3, 4), op(4, 3), op(0, 9));
}

template <int Uplo, int Flags>
static void symmetric_case(const char* tag, int n, int nnz_per_col, int k, int m, SortRule rule)
{
    // a general sparse matrix of which the operator reads one triangle (test/SymEigs.cpp:30-52)
    std::vector<Eigen::Triplet<double>> trip;
    Lcg g(12345);
    for (int j = 0; j < n; j++)
    {
        trip.emplace_back(j, j, g.uniform() * 4.0);
        for (int t = 0; t < nnz_per_col; t++)
        {
            const int i = int((g.uniform() + 0.5) * n) % n;
            trip.emplace_back(i, j, g.uniform());  // duplicates are summed by setFromTriplets
        }
    }
    Eigen::SparseMatrix<double, Flags> A(n, n);
    A.setFromTriplets(trip.begin(), trip.end());
    using Op = SparseSymMatProd<double, Uplo, Flags>;
    Op op(A);
    SymEigsSolver<Op> eigs(op, k, m);
    eigs.init();
    const int nconv = int(eigs.compute(rule));
    std::printf("%s.info: %d\n%s.nconv: %d\n", tag, int(eigs.info()), tag, nconv);
    std::printf("%s.counts: %d %d\n", tag, int(eigs.num_iterations()), int(eigs.num_operations()));
    Eigen::VectorXd evals = eigs.eigenvalues();
    Eigen::MatrixXd evecs = eigs.eigenvectors();
    print_vec((std::string(tag) + ".evalues").c_str(), evals);
    // ||A U - U D||_inf through the operator's own operator* (SparseSymMatProd.h:93-96)
    Eigen::MatrixXd AU = op * evecs;
    double err = 0.0;
    for (Eigen::Index j = 0; j < evecs.cols(); j++)
        for (Eigen::Index i = 0; i < evecs.rows(); i++)
            err = std::max(err, std::abs(AU(i, j) - evecs(i, j) * evals[j]));
    std::printf("%s.residual_ok: %d\n", tag, err <= 1e-9 ? 1 : 0);
    std::printf("%s.coeff: %.17g %.17g %.17g\n", tag, op(5, 5), op(7, 7), op(0, 1));
}

static void hermitian_case()
{
    const int n = 60, k = 5, m = 15;
    std::vector<Eigen::Triplet<std::complex<double>>> trip;
    Lcg g(777);
    for (int j = 0; j < n; j++)
    {
        trip.emplace_back(j, j, std::complex<double>(g.uniform() * 4.0, 0.0));
        for (int t = 0; t < 4; t++)
        {
            const int i = int((g.uniform() + 0.5) * n) % n;
            if (i != j)
                trip.emplace_back(i, j, std::complex<double>(g.uniform(), g.uniform()));
        }
    }
    Eigen::SparseMatrix<std::complex<double>> A(n, n);
    A.setFromTriplets(trip.begin(), trip.end());
    using Op = SparseHermMatProd<std::complex<double>>;
    Op op(A);
    HermEigsSolver<Op> eigs(op, k, m);
    eigs.init();
    const int nconv = int(eigs.compute(SortRule::LargestAlge));
    std::printf("herm.info: %d\nherm.nconv: %d\n", int(eigs.info()), nconv);
    Eigen::VectorXd evals = eigs.eigenvalues();
    print_vec("herm.evalues", evals);
    Eigen::MatrixXcd U = eigs.eigenvectors();
    std::printf("herm.vec_shape: %d %d\n", int(U.rows()), int(U.cols()));
}

static void argument_errors()
{
    Eigen::SparseMatrix<double> I(10, 10);
    for (int i = 0; i < 10; i++)
        I.insert(i, i) = 1.0;
    I.makeCompressed();
    SparseSymMatProd<double> op(I);
    const int bad[4][2] = {{0, 5}, {10, 12}, {3, 3}, {3, 11}};
    std::printf("errors:");
    for (auto& b : bad)
    {
        try
        {
            SymEigsSolver<SparseSymMatProd<double>> eigs(op, b[0], b[1]);
            std::printf(" none");
        }
        catch (const std::invalid_argument&)
        {
            std::printf(" invalid_argument");
        }
    }
    try
    {
        SymEigsSolver<SparseSymMatProd<double>> eigs(op, 3, 6);
        Eigen::VectorXd zero = Eigen::VectorXd::Zero(10);
        eigs.init(zero.data());
        std::printf(" none");
    }
    catch (const std::invalid_argument&)
    {
        std::printf(" invalid_argument");
    }
    std::printf("\n");
}

int main()
{
    readme_sparse_general();
    symmetric_case<Eigen::Lower, Eigen::ColMajor>("sym_lower_col", 400, 6, 8, 24, SortRule::LargestAlge);
    symmetric_case<Eigen::Upper, Eigen::RowMajor>("sym_upper_row", 300, 5, 6, 20, SortRule::BothEnds);
    symmetric_case<Eigen::Upper, Eigen::ColMajor>("sym_upper_col", 200, 4, 4, 16, SortRule::SmallestAlge);
    hermitian_case();
    argument_errors();
    return 0;
}

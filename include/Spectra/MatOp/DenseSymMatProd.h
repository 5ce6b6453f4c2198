// B200 shim of Spectra/MatOp/DenseSymMatProd.h:20-105 (the reference's default OpType of SymEigsSolver) and of
// MatOp/DenseGenMatProd.h: the dense matrix is uploaded once as a full compressed matrix (every entry stored) and multiplied by the
// same sm_100a CSR SpMV kernel as the sparse wrappers -- a dense n x n operand is simply a CSR with n entries per row (one warp per
// row).  Dense operators are outside the hot-path scope of this build; they are provided so that code written
// against the reference's defaults compiles and runs.  Only the `Uplo` triangle is read by DenseSymMatProd (selfadjointView<Uplo>).
#ifndef SPECTRA_B200_DENSE_SYM_MAT_PROD_H
#define SPECTRA_B200_DENSE_SYM_MAT_PROD_H

#include <stdexcept>
#include <vector>

#include "SparseSymMatProd.h"

namespace Spectra {

namespace b200 {
// dense column-/row-major n x n matrix -> compressed arrays with every entry present
template <typename Scalar>
struct DenseAsCompressed
{
    std::vector<int64_t> outer;  // 64-bit offsets: n * n passes 2^31 at n = 46341
    std::vector<int32_t> inner;
    std::vector<Scalar> values;
    // rows != cols is rejected before a single element is read (the Eigen constructors delegate here)
    static Index checked_square(Index rows, Index cols, const char* what)
    {
        if (rows != cols)
            throw std::invalid_argument(what);
        return rows;
    }
    DenseAsCompressed(Index n, const Scalar* data)
    {
        if (n < 0 || data == nullptr)
            throw std::invalid_argument("dense operator: bad matrix");
        outer.resize(static_cast<size_t>(n) + 1);
        inner.resize(static_cast<size_t>(n * n));
        values.assign(data, data + n * n);
        for (Index o = 0; o <= n; o++)
            outer[static_cast<size_t>(o)] = static_cast<int64_t>(o) * n;
        for (Index o = 0; o < n; o++)
            for (Index k = 0; k < n; k++)
                inner[static_cast<size_t>(o * n + k)] = static_cast<int32_t>(k);
    }
};
}  // namespace b200

template <typename Scalar_, int Uplo = SPECTRA_B200_LOWER, int Flags = SPECTRA_B200_COLMAJOR>
class DenseSymMatProd : public b200::SparseOpBase
{
    static_assert(b200::IsSupportedScalar<Scalar_>::value, "the B200 path implements Scalar = double, and float with fp64 device arithmetic");
    b200::DenseAsCompressed<Scalar_> m_c;

public:
    using Scalar = Scalar_;

    // n x n matrix in `Flags` storage order (ColMajor: data[i + j * n])
    DenseSymMatProd(Index n, const Scalar* data) : m_c(n, data)
    {
        create_any(n, m_c.outer.data(), m_c.inner.data(), m_c.values.data(), Flags == SPECTRA_B200_ROWMAJOR,
                   Uplo == SPECTRA_B200_LOWER ? SB200_SYM_LOWER : SB200_SYM_UPPER);
    }
#ifdef SPECTRA_B200_HAS_EIGEN
    // Same constructor as the reference (DenseSymMatProd.h:45-53)
    explicit DenseSymMatProd(const Eigen::Matrix<Scalar, Eigen::Dynamic, Eigen::Dynamic, Flags>& mat) :
        DenseSymMatProd(b200::DenseAsCompressed<Scalar>::checked_square(mat.rows(), mat.cols(), "DenseSymMatProd: matrix must be square"), mat.data())
    {
    }
#endif
};

template <typename Scalar_, int Flags = SPECTRA_B200_COLMAJOR>
class DenseGenMatProd : public b200::SparseOpBase
{
    static_assert(b200::IsSupportedScalar<Scalar_>::value, "the B200 path implements Scalar = double, and float with fp64 device arithmetic");
    b200::DenseAsCompressed<Scalar_> m_c;

public:
    using Scalar = Scalar_;

    DenseGenMatProd(Index n, const Scalar* data) : m_c(n, data)
    {
        create_any(n, m_c.outer.data(), m_c.inner.data(), m_c.values.data(), Flags == SPECTRA_B200_ROWMAJOR, SB200_GENERAL);
    }
#ifdef SPECTRA_B200_HAS_EIGEN
    explicit DenseGenMatProd(const Eigen::Matrix<Scalar, Eigen::Dynamic, Eigen::Dynamic, Flags>& mat) :
        DenseGenMatProd(b200::DenseAsCompressed<Scalar>::checked_square(mat.rows(), mat.cols(), "DenseGenMatProd: matrix must be square"), mat.data())
    {
    }
#endif
};

}  // namespace Spectra

// Scalar = std::complex<double> (test/ComplexEigs.cpp:74-76 uses DenseGenMatProd<std::complex<double>>): a dense general complex matrix
// behind the complex CSR SpMV kernel of SparseGenMatProd<std::complex<double>>
#include "DenseHermMatProd.h"
#include "SparseGenMatProd.h"

namespace Spectra {

template <int Flags>
class DenseGenMatProd<std::complex<double>, Flags> : private b200::DenseAsCompressed32<std::complex<double>>, public SparseGenMatProd<std::complex<double>, Flags, int>
{
    using Pattern = b200::DenseAsCompressed32<std::complex<double>>;
    using Base = SparseGenMatProd<std::complex<double>, Flags, int>;

public:
    using Scalar = std::complex<double>;

    DenseGenMatProd(Index n, const Scalar* data) : Pattern(Pattern::checked(n, n, "")), Base(n, Pattern::outer.data(), Pattern::inner.data(), data) {}
#ifdef SPECTRA_B200_HAS_EIGEN
    explicit DenseGenMatProd(const Eigen::Matrix<Scalar, Eigen::Dynamic, Eigen::Dynamic, Flags>& mat) :
        Pattern(Pattern::checked(mat.rows(), mat.cols(), "DenseGenMatProd: matrix must be square")), Base(mat.rows(), Pattern::outer.data(), Pattern::inner.data(), mat.data())
    {
    }
#endif
};

}  // namespace Spectra
#endif

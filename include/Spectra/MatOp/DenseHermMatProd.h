// B200 shim of Spectra/MatOp/DenseHermMatProd.h:21-100 (the reference's default OpType of HermEigsSolver): y = A x for a dense complex
// Hermitian matrix of which only the `Uplo` triangle is read.  The matrix is uploaded as a full compressed matrix behind the complex CSR
// SpMV kernel of SparseHermMatProd.  Dense operators are outside the hot-path scope of this build; the wrapper exists so
// that code written against the reference's defaults (test/HermEigs.cpp) compiles and runs.
#ifndef SPECTRA_B200_DENSE_HERM_MAT_PROD_H
#define SPECTRA_B200_DENSE_HERM_MAT_PROD_H

#include <complex>
#include <stdexcept>
#include <vector>

#include "SparseHermMatProd.h"

namespace Spectra {

namespace b200 {
// dense n x n matrix -> compressed arrays with every entry present, 32-bit offsets (n * n < 2^31)
template <typename Scalar>
struct DenseAsCompressed32
{
    std::vector<int32_t> outer, inner;
    static Index checked(Index rows, Index cols, const char* what)
    {
        if (rows != cols)
            throw std::invalid_argument(what);
        if (rows >= 46341)
            throw std::invalid_argument("dense complex operator: n must stay below 46341");
        return rows;
    }
    explicit DenseAsCompressed32(Index n)
    {
        outer.resize(static_cast<size_t>(n) + 1);
        inner.resize(static_cast<size_t>(n * n));
        for (Index o = 0; o <= n; o++)
            outer[static_cast<size_t>(o)] = static_cast<int32_t>(o * n);
        for (Index o = 0; o < n; o++)
            for (Index k = 0; k < n; k++)
                inner[static_cast<size_t>(o * n + k)] = static_cast<int32_t>(k);
    }
};
}  // namespace b200

template <typename Scalar_, int Uplo = SPECTRA_B200_LOWER, int Flags = SPECTRA_B200_COLMAJOR>
class DenseHermMatProd : private b200::DenseAsCompressed32<Scalar_>, public SparseHermMatProd<Scalar_, Uplo, Flags, int>
{
    using Pattern = b200::DenseAsCompressed32<Scalar_>;
    using Base = SparseHermMatProd<Scalar_, Uplo, Flags, int>;

public:
    using Scalar = Scalar_;

    // n x n matrix in `Flags` storage order (ColMajor: data[i + j * n]); the values are uploaded at construction
    DenseHermMatProd(Index n, const Scalar* data) : Pattern(Pattern::checked(n, n, "")), Base(n, Pattern::outer.data(), Pattern::inner.data(), data) {}
#ifdef SPECTRA_B200_HAS_EIGEN
    // Same constructor as the reference (DenseHermMatProd.h:46-54)
    explicit DenseHermMatProd(const Eigen::Matrix<Scalar, Eigen::Dynamic, Eigen::Dynamic, Flags>& mat) :
        Pattern(Pattern::checked(mat.rows(), mat.cols(), "DenseHermMatProd: matrix must be square")), Base(mat.rows(), Pattern::outer.data(), Pattern::inner.data(), mat.data())
    {
    }
#endif
};

}  // namespace Spectra
#endif

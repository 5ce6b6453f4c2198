// B200 shim of Spectra/MatOp/SparseSymMatProd.h:30-105: y = A x for a real symmetric sparse matrix of
// which only the `Uplo` triangle is read (selfadjointView<Uplo>), executed by the sm_100a CSR SpMV kernel.
#ifndef SPECTRA_B200_SPARSE_SYM_MAT_PROD_H
#define SPECTRA_B200_SPARSE_SYM_MAT_PROD_H

#include <type_traits>

#include "../b200/Common.h"

namespace Spectra {

#ifndef SPECTRA_B200_HAS_EIGEN
// stand-ins for the Eigen flags used as template arguments by the reference
namespace Eigen_flags {
enum { Lower = 1, Upper = 2, ColMajor = 0, RowMajor = 1 };
}
#define SPECTRA_B200_LOWER Eigen_flags::Lower
#define SPECTRA_B200_UPPER Eigen_flags::Upper
#define SPECTRA_B200_COLMAJOR Eigen_flags::ColMajor
#define SPECTRA_B200_ROWMAJOR Eigen_flags::RowMajor
#else
#define SPECTRA_B200_LOWER Eigen::Lower
#define SPECTRA_B200_UPPER Eigen::Upper
#define SPECTRA_B200_COLMAJOR Eigen::ColMajor
#define SPECTRA_B200_ROWMAJOR Eigen::RowMajor
#endif

template <typename Scalar_, int Uplo = SPECTRA_B200_LOWER, int Flags = SPECTRA_B200_COLMAJOR, typename StorageIndex = int>
class SparseSymMatProd : public b200::SparseOpBase
{
    static_assert(b200::IsSupportedScalar<Scalar_>::value, "the B200 path implements Scalar = double, and float with fp64 device arithmetic");

public:
    using Scalar = Scalar_;

    // Raw compressed arrays with Eigen's layout: outer[n + 1], inner[nnz], values[nnz].
    SparseSymMatProd(Index n, const StorageIndex* outer, const StorageIndex* inner, const Scalar* values)
    {
        create_any(n, outer, inner, values, Flags == SPECTRA_B200_ROWMAJOR,
               Uplo == SPECTRA_B200_LOWER ? SB200_SYM_LOWER : SB200_SYM_UPPER);
    }
#ifdef SPECTRA_B200_HAS_EIGEN
    // Same constructor as the reference (SparseSymMatProd.h:57-65); the matrix must outlive the operator.
    explicit SparseSymMatProd(const Eigen::SparseMatrix<Scalar, Flags, StorageIndex>& mat)
    {
        if (mat.rows() != mat.cols())
            throw std::invalid_argument("SparseSymMatProd: matrix must be square");
        if (mat.isCompressed())
            create_any(mat.rows(), mat.outerIndexPtr(), mat.innerIndexPtr(), mat.valuePtr(), Flags == Eigen::RowMajor,
                       Uplo == Eigen::Lower ? SB200_SYM_LOWER : SB200_SYM_UPPER);
        else
        {
            m_packed.pack(mat);  // uncompressed mode (reserve + insert): accepted by the reference's Eigen::Ref, packed here
            create_any(mat.rows(), m_packed.outer.data(), m_packed.inner.data(), m_packed.values.data(), Flags == Eigen::RowMajor,
                       Uplo == Eigen::Lower ? SB200_SYM_LOWER : SB200_SYM_UPPER);
        }
    }

private:
    b200::PackedCopy<StorageIndex, Scalar> m_packed;
#endif
};

}  // namespace Spectra
#endif

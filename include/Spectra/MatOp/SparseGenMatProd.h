// B200 shim of Spectra/MatOp/SparseGenMatProd.h:29-104: y = A x for a general real sparse matrix.
#ifndef SPECTRA_B200_SPARSE_GEN_MAT_PROD_H
#define SPECTRA_B200_SPARSE_GEN_MAT_PROD_H

#include <type_traits>

#include "SparseSymMatProd.h"

namespace Spectra {

template <typename Scalar_, int Flags = SPECTRA_B200_COLMAJOR, typename StorageIndex = int>
class SparseGenMatProd : public b200::SparseOpBase
{
    static_assert(b200::IsSupportedScalar<Scalar_>::value, "the B200 path implements Scalar = double, and float with fp64 device arithmetic");

public:
    using Scalar = Scalar_;

    SparseGenMatProd(Index n, const StorageIndex* outer, const StorageIndex* inner, const Scalar* values)
    {
        create_any(n, outer, inner, values, Flags == SPECTRA_B200_ROWMAJOR, SB200_GENERAL);
    }
#ifdef SPECTRA_B200_HAS_EIGEN
    explicit SparseGenMatProd(const Eigen::SparseMatrix<Scalar, Flags, StorageIndex>& mat)
    {
        if (!mat.isCompressed())
            throw std::invalid_argument("SparseGenMatProd: matrix must be in compressed mode (call makeCompressed())");
        if (mat.rows() != mat.cols())
            throw std::invalid_argument("SparseGenMatProd: matrix must be square");
        create_any(mat.rows(), mat.outerIndexPtr(), mat.innerIndexPtr(), mat.valuePtr(), Flags == Eigen::RowMajor,
               SB200_GENERAL);
    }
#endif
};

}  // namespace Spectra
#endif

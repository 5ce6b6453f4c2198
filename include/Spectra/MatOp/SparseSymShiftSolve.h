// B200 shim of Spectra/MatOp/SparseSymShiftSolve.h:30-110: y = (A - sigma I)^{-1} x for a real symmetric sparse matrix
// (only the `Uplo` triangle is read).  set_shift() factorises on the device by block cyclic reduction, which requires a
// banded matrix (half-bandwidth <= 32); wider patterns throw std::invalid_argument at construction.
#ifndef SPECTRA_B200_SPARSE_SYM_SHIFT_SOLVE_H
#define SPECTRA_B200_SPARSE_SYM_SHIFT_SOLVE_H

#include "SparseSymMatProd.h"

namespace Spectra {

template <typename Scalar_, int Uplo = SPECTRA_B200_LOWER, int Flags = SPECTRA_B200_COLMAJOR, typename StorageIndex = int>
class SparseSymShiftSolve : public b200::SparseOpBase
{
    static_assert(b200::IsSupportedScalar<Scalar_>::value, "the B200 path implements Scalar = double, and float with fp64 device arithmetic");

public:
    using Scalar = Scalar_;

    // Raw compressed arrays with Eigen's layout: outer[n + 1], inner[nnz], values[nnz].
    SparseSymShiftSolve(Index n, const StorageIndex* outer, const StorageIndex* inner, const Scalar* values)
    {
        create_any(n, outer, inner, values, Flags == SPECTRA_B200_ROWMAJOR,
               Uplo == SPECTRA_B200_LOWER ? SB200_SYM_LOWER : SB200_SYM_UPPER, true);
    }
#ifdef SPECTRA_B200_HAS_EIGEN
    // Same constructor as the reference (SparseSymShiftSolve.h:57-68)
    explicit SparseSymShiftSolve(const Eigen::SparseMatrix<Scalar, Flags, StorageIndex>& mat)
    {
        if (mat.rows() != mat.cols())
            throw std::invalid_argument("SparseSymShiftSolve: matrix must be square");
        if (mat.isCompressed())
            create_any(mat.rows(), mat.outerIndexPtr(), mat.innerIndexPtr(), mat.valuePtr(), Flags == Eigen::RowMajor,
                       Uplo == Eigen::Lower ? SB200_SYM_LOWER : SB200_SYM_UPPER, true);
        else
        {
            m_packed.pack(mat);
            create_any(mat.rows(), m_packed.outer.data(), m_packed.inner.data(), m_packed.values.data(), Flags == Eigen::RowMajor,
                       Uplo == Eigen::Lower ? SB200_SYM_LOWER : SB200_SYM_UPPER, true);
        }
    }

private:
    b200::PackedCopy<StorageIndex, Scalar> m_packed;

public:
#endif

    // set_shift(sigma) (SparseSymShiftSolve.h:85-95); throws std::invalid_argument when the factorisation fails
    void set_shift(const Scalar& sigma) { b200::check(sb200_op_set_shift(m_op, static_cast<double>(sigma))); }
    // perform_op(x_in, y_out) = solve (:104-109) is SparseOpBase::perform_op
};

}  // namespace Spectra
#endif

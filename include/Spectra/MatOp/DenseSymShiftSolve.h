// B200 shim of Spectra/MatOp/DenseSymShiftSolve.h:20-105 (the reference's default OpType of SymEigsShiftSolver): y = (A - sigma I)^{-1} x
// for a dense real symmetric matrix of which only the `Uplo` triangle is read.  The matrix is uploaded as a full compressed matrix and
// handed to the same device shift-solve operator as SparseSymShiftSolve (band_solve.cu: explicit inverse for n <= 2048, block elimination
// with one dense block beyond).  Dense operators are outside the hot-path scope of this build; the wrapper exists so
// that code written against the reference's defaults -- its own test/SymEigsShift.cpp and test/Example1.cpp among it -- compiles and runs.
#ifndef SPECTRA_B200_DENSE_SYM_SHIFT_SOLVE_H
#define SPECTRA_B200_DENSE_SYM_SHIFT_SOLVE_H

#include "DenseSymMatProd.h"

namespace Spectra {

template <typename Scalar_, int Uplo = SPECTRA_B200_LOWER, int Flags = SPECTRA_B200_COLMAJOR>
class DenseSymShiftSolve : public b200::SparseOpBase
{
    static_assert(b200::IsSupportedScalar<Scalar_>::value, "the B200 path implements Scalar = double, and float with fp64 device arithmetic");
    b200::DenseAsCompressed<Scalar_> m_c;

public:
    using Scalar = Scalar_;

    // n x n matrix in `Flags` storage order (ColMajor: data[i + j * n])
    DenseSymShiftSolve(Index n, const Scalar* data) : m_c(n, data)
    {
        create_any(n, m_c.outer.data(), m_c.inner.data(), m_c.values.data(), Flags == SPECTRA_B200_ROWMAJOR,
                   Uplo == SPECTRA_B200_LOWER ? SB200_SYM_LOWER : SB200_SYM_UPPER, true);
    }
#ifdef SPECTRA_B200_HAS_EIGEN
    // Same constructor as the reference (DenseSymShiftSolve.h:50-61)
    explicit DenseSymShiftSolve(const Eigen::Matrix<Scalar, Eigen::Dynamic, Eigen::Dynamic, Flags>& mat) :
        DenseSymShiftSolve(b200::DenseAsCompressed<Scalar>::checked_square(mat.rows(), mat.cols(), "DenseSymShiftSolve: matrix must be square"), mat.data())
    {
    }
#endif
    // set_shift(sigma) (DenseSymShiftSolve.h:79-85); throws std::invalid_argument when the factorisation fails
    void set_shift(const Scalar& sigma) { b200::check(sb200_op_set_shift(m_op, static_cast<double>(sigma))); }
};

}  // namespace Spectra
#endif

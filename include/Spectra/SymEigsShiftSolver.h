// B200 shim of Spectra/SymEigsShiftSolver.h:148-196: eigenvalues closest to sigma through implicitly restarted Lanczos on
// (A - sigma I)^{-1}; Ritz values nu are mapped back by lambda = 1/nu + sigma before sorting (:163-169).
#ifndef SPECTRA_B200_SYM_EIGS_SHIFT_SOLVER_H
#define SPECTRA_B200_SYM_EIGS_SHIFT_SOLVER_H

#include "MatOp/SparseSymShiftSolve.h"
#include "SymEigsSolver.h"

namespace Spectra {

template <typename OpType = SparseSymShiftSolve<double>>
class SymEigsShiftSolver : public SymEigsSolver<OpType>
{
    using Base = SymEigsSolver<OpType>;
    // device operators are shifted inside sb200_sym_create_shift; user-defined host operators through their own set_shift
    static void shift_host_op(OpType&, double, std::true_type) {}
    static void shift_host_op(OpType& op, double sigma, std::false_type) { op.set_shift(static_cast<typename OpType::Scalar>(sigma)); }

public:
    using Scalar = typename OpType::Scalar;

    SymEigsShiftSolver(OpType& op, Index nev, Index ncv, const Scalar& sigma) : Base(op, nev, ncv, static_cast<double>(sigma), typename Base::ShiftInvert())
    {
        shift_host_op(op, static_cast<double>(sigma), std::is_base_of<b200::SparseOpBase, OpType>());
    }
};

}  // namespace Spectra
#endif

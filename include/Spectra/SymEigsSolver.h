// B200 shim of Spectra/SymEigsSolver.h:133-160 (+ the public surface of HermEigsBase.h:257-478):
// implicitly restarted Lanczos on the GPU behind the same constructor / init / compute / getters.
#ifndef SPECTRA_B200_SYM_EIGS_SOLVER_H
#define SPECTRA_B200_SYM_EIGS_SOLVER_H

#include <algorithm>

#include "MatOp/DenseSymMatProd.h"
#include "MatOp/SparseSymMatProd.h"
#include "Util/CompInfo.h"
#include "Util/SelectionRule.h"
#include "b200/Common.h"

namespace Spectra {

template <typename OpType = DenseSymMatProd<double>>  // the reference's default (SymEigsSolver.h:132)
class SymEigsSolver
{
    b200::OpBinding<OpType> m_bind;  // device-resident sparse operator, or a host-callback adapter for any other OpType
    sb200_sym_solver* m_s = nullptr;
    const OpType& m_op;  // the operator must outlive the solver (HermEigsBase.h:257-258)
    Index m_nev;

protected:
    // shift-and-invert construction used by SymEigsShiftSolver (SymEigsShiftSolver.h:190-195)
    struct ShiftInvert
    {
    };
    SymEigsSolver(OpType& op, Index nev, Index ncv, double sigma, ShiftInvert) : m_bind(op), m_op(op), m_nev(nev)
    {
        b200::check(sb200_sym_create_shift(m_bind.handle(), nev, ncv, sigma, &m_s));
    }

public:
    using Scalar = typename OpType::Scalar;  // double, or float (float storage at the boundary, fp64 arithmetic on the device)
    using Vector = b200::VectorOf<Scalar>;
    using Matrix = b200::MatrixOf<Scalar>;

    SymEigsSolver(OpType& op, Index nev, Index ncv) : m_bind(op), m_op(op), m_nev(nev) { b200::check(sb200_sym_create(m_bind.handle(), nev, ncv, &m_s)); }
    SymEigsSolver(const SymEigsSolver&) = delete;
    SymEigsSolver& operator=(const SymEigsSolver&) = delete;
    virtual ~SymEigsSolver()
    {
        if (m_s)
            sb200_sym_destroy(m_s);
    }

    void init(const Scalar* init_resid)
    {
        std::vector<double> buf;
        b200::check(sb200_sym_init(m_s, b200::widen(init_resid, m_op.rows(), buf)));
    }
    void init() { b200::check(sb200_sym_init(m_s, nullptr)); }

    Index compute(SortRule selection = SortRule::LargestMagn, Index maxit = 1000, Scalar tol = 1e-10, SortRule sorting = SortRule::LargestAlge)
    {
        int64_t nconv = 0;
        b200::check(sb200_sym_compute(m_s, static_cast<int>(selection), maxit, static_cast<double>(tol), static_cast<int>(sorting), &nconv));
        return static_cast<Index>(nconv);
    }

    CompInfo info() const
    {
        int v = 0;
        b200::check(sb200_sym_info(m_s, &v));
        return static_cast<CompInfo>(v);
    }
    Index num_iterations() const
    {
        int64_t v = 0;
        b200::check(sb200_sym_num_iterations(m_s, &v));
        return static_cast<Index>(v);
    }
    Index num_operations() const
    {
        int64_t v = 0;
        b200::check(sb200_sym_num_operations(m_s, &v));
        return static_cast<Index>(v);
    }

    Vector eigenvalues() const
    {
        std::vector<double> buf(static_cast<size_t>(m_nev));
        int64_t cnt = 0;
        b200::check(sb200_sym_eigenvalues(m_s, buf.data(), &cnt));
        Vector res(static_cast<Index>(cnt));
        for (int64_t i = 0; i < cnt; i++)
            res[i] = static_cast<Scalar>(buf[static_cast<size_t>(i)]);
        return res;
    }

    Matrix eigenvectors(Index nvec) const
    {
        nvec = (std::min)(nvec, m_nev);
        b200::Matrix buf(m_op.rows(), (std::max)(nvec, Index(1)));
        int64_t cnt = 0;
        b200::check(sb200_sym_eigenvectors(m_s, nvec, buf.data(), &cnt));
        b200::shrink_cols(buf, static_cast<Index>(cnt));
        return narrow(std::move(buf), std::is_same<Scalar, double>());
    }
    Matrix eigenvectors() const { return eigenvectors(m_nev); }

private:
    static Matrix narrow(b200::Matrix&& M, std::true_type) { return std::move(M); }
    static Matrix narrow(b200::Matrix&& M, std::false_type)
    {
        Matrix res(M.rows(), M.cols());
        for (Index q = 0; q < M.rows() * M.cols(); q++)
            res.data()[q] = static_cast<Scalar>(M.data()[q]);
        return res;
    }
};

}  // namespace Spectra
#endif

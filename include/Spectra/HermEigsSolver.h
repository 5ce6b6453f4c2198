// B200 shim of Spectra/HermEigsSolver.h:121-122 (+ the public surface of HermEigsBase.h:257-478 with a complex Scalar):
// implicitly restarted Lanczos for complex Hermitian operators on the GPU.  Eigenvalues are real, eigenvectors complex.
// (device-verified in round 2.)
#ifndef SPECTRA_B200_HERM_EIGS_SOLVER_H
#define SPECTRA_B200_HERM_EIGS_SOLVER_H

#include <algorithm>
#include <complex>

#include "MatOp/SparseHermMatProd.h"
#include "Util/CompInfo.h"
#include "Util/SelectionRule.h"
#include "b200/Common.h"

namespace Spectra {

template <typename OpType = SparseHermMatProd<std::complex<double>>>
class HermEigsSolver
{
    b200::OpBindingZ<OpType> m_bind;  // device-resident SparseHermMatProd, or a host-callback adapter for any other complex OpType
    sb200_sym_solver* m_s = nullptr;
    const OpType& m_op;  // the operator must outlive the solver (HermEigsBase.h:257-258)
    Index m_nev;

public:
    using Scalar = typename OpType::Scalar;   // std::complex<double>
    using RealScalar = double;
    using RealVector = b200::Vector;
    using Matrix = b200::ComplexMatrix;

    HermEigsSolver(OpType& op, Index nev, Index ncv) : m_bind(op), m_op(op), m_nev(nev) { b200::check(sb200_herm_create(m_bind.handle(), nev, ncv, &m_s)); }
    HermEigsSolver(const HermEigsSolver&) = delete;
    HermEigsSolver& operator=(const HermEigsSolver&) = delete;
    ~HermEigsSolver()
    {
        if (m_s)
            sb200_sym_destroy(m_s);
    }

    void init(const Scalar* init_resid) { b200::check(sb200_sym_init(m_s, reinterpret_cast<const double*>(init_resid))); }
    void init() { b200::check(sb200_sym_init(m_s, nullptr)); }

    Index compute(SortRule selection = SortRule::LargestMagn, Index maxit = 1000, RealScalar tol = 1e-10, SortRule sorting = SortRule::LargestAlge)
    {
        int64_t nconv = 0;
        b200::check(sb200_sym_compute(m_s, static_cast<int>(selection), maxit, tol, static_cast<int>(sorting), &nconv));
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

    RealVector eigenvalues() const
    {
        std::vector<double> buf(static_cast<size_t>(m_nev));
        int64_t cnt = 0;
        b200::check(sb200_sym_eigenvalues(m_s, buf.data(), &cnt));
        RealVector res(static_cast<Index>(cnt));
        for (int64_t i = 0; i < cnt; i++)
            res[i] = buf[static_cast<size_t>(i)];
        return res;
    }

    Matrix eigenvectors(Index nvec) const
    {
        nvec = (std::min)(nvec, m_nev);
        Matrix res(m_op.rows(), (std::max)(nvec, Index(1)));
        int64_t cnt = 0;
        b200::check(sb200_sym_eigenvectors(m_s, nvec, reinterpret_cast<double*>(res.data()), &cnt));
        b200::shrink_cols(res, static_cast<Index>(cnt));
        return res;
    }
    Matrix eigenvectors() const { return eigenvectors(m_nev); }
};

}  // namespace Spectra
#endif

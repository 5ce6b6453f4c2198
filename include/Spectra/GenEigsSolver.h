// B200 shim of Spectra/GenEigsSolver.h:158-186 (+ the public surface of GenEigsBase.h:409-611):
// implicitly restarted Arnoldi for real nonsymmetric matrices; complex Ritz pairs.
#ifndef SPECTRA_B200_GEN_EIGS_SOLVER_H
#define SPECTRA_B200_GEN_EIGS_SOLVER_H

#include <algorithm>

#include "MatOp/DenseGenMatProd.h"
#include "MatOp/SparseGenMatProd.h"
#include "Util/CompInfo.h"
#include "Util/SelectionRule.h"
#include "b200/Common.h"

namespace Spectra {

namespace b200 {
template <typename T>
struct IsComplexScalar : std::false_type
{
};
template <typename T>
struct IsComplexScalar<std::complex<T>> : std::true_type
{
};
// real operators bind through OpBinding (device wrapper or host-callback adapter), complex ones through OpBindingZ
template <typename OpType, bool Complex = IsComplexScalar<typename OpType::Scalar>::value>
struct GenBinding : OpBinding<OpType>
{
    using OpBinding<OpType>::OpBinding;
};
template <typename OpType>
struct GenBinding<OpType, true> : OpBindingZ<OpType>
{
    using OpBindingZ<OpType>::OpBindingZ;
};
}  // namespace b200

// Scalar = double / float: real nonsymmetric problems (complex Ritz pairs).  Scalar = std::complex<double>: GenEigsBase with a complex
// Scalar (GenEigsBase.h:111-140, test/ComplexEigs.cpp); device-verified in round 2, see DESIGN.md §4b.
template <typename OpType = DenseGenMatProd<double>>  // the reference's default (GenEigsSolver.h:157)
class GenEigsSolver
{
    b200::GenBinding<OpType> m_bind;
    sb200_gen_solver* m_s = nullptr;
    const OpType& m_op;
    Index m_nev;

public:
    using Scalar = typename OpType::Scalar;  // double, float (float storage at the boundary, fp64 arithmetic on the device), or std::complex<double>
    static constexpr bool kComplex = b200::IsComplexScalar<Scalar>::value;
    using RealScalar = typename std::conditional<kComplex, double, Scalar>::type;
    using Complex = typename std::conditional<kComplex, Scalar, std::complex<Scalar>>::type;
    using ComplexVector = b200::VectorOf<Complex>;
    using ComplexMatrix = b200::MatrixOf<Complex>;

    GenEigsSolver(OpType& op, Index nev, Index ncv) : m_bind(op), m_op(op), m_nev(nev) { b200::check(sb200_gen_create(m_bind.handle(), nev, ncv, &m_s)); }
    GenEigsSolver(const GenEigsSolver&) = delete;
    GenEigsSolver& operator=(const GenEigsSolver&) = delete;
    virtual ~GenEigsSolver()
    {
        if (m_s)
            sb200_gen_destroy(m_s);
    }

    void init(const Scalar* init_resid) { init_impl(init_resid, std::integral_constant<bool, kComplex>()); }
    void init() { b200::check(sb200_gen_init(m_s, nullptr)); }

    Index compute(SortRule selection = SortRule::LargestMagn, Index maxit = 1000, RealScalar tol = 1e-10, SortRule sorting = SortRule::LargestMagn)
    {
        int64_t nconv = 0;
        b200::check(sb200_gen_compute(m_s, static_cast<int>(selection), maxit, static_cast<double>(tol), static_cast<int>(sorting), &nconv));
        return static_cast<Index>(nconv);
    }
    CompInfo info() const
    {
        int v = 0;
        b200::check(sb200_gen_info(m_s, &v));
        return static_cast<CompInfo>(v);
    }
    Index num_iterations() const
    {
        int64_t v = 0;
        b200::check(sb200_gen_num_iterations(m_s, &v));
        return static_cast<Index>(v);
    }
    Index num_operations() const
    {
        int64_t v = 0;
        b200::check(sb200_gen_num_operations(m_s, &v));
        return static_cast<Index>(v);
    }

    ComplexVector eigenvalues() const
    {
        std::vector<double> buf(static_cast<size_t>(2 * m_nev));
        int64_t cnt = 0;
        b200::check(sb200_gen_eigenvalues(m_s, buf.data(), &cnt));
        ComplexVector res(static_cast<Index>(cnt));
        for (int64_t i = 0; i < cnt; i++)
            res[i] = Complex(static_cast<RealScalar>(buf[static_cast<size_t>(2 * i)]), static_cast<RealScalar>(buf[static_cast<size_t>(2 * i + 1)]));
        return res;
    }

    ComplexMatrix eigenvectors(Index nvec) const
    {
        nvec = (std::min)(nvec, m_nev);
        b200::ComplexMatrix buf(m_op.rows(), (std::max)(nvec, Index(1)));
        int64_t cnt = 0;
        // std::complex<double> is layout-compatible with interleaved (re, im) pairs
        b200::check(sb200_gen_eigenvectors(m_s, nvec, reinterpret_cast<double*>(buf.data()), &cnt));
        b200::shrink_cols(buf, static_cast<Index>(cnt));
        return narrow(std::move(buf), std::integral_constant<bool, std::is_same<RealScalar, double>::value>());
    }
    ComplexMatrix eigenvectors() const { return eigenvectors(m_nev); }

private:
    void init_impl(const Scalar* init_resid, std::false_type)
    {
        std::vector<double> buf;
        b200::check(sb200_gen_init(m_s, b200::widen(init_resid, m_op.rows(), buf)));
    }
    void init_impl(const Scalar* init_resid, std::true_type) { b200::check(sb200_gen_init(m_s, reinterpret_cast<const double*>(init_resid))); }
    static ComplexMatrix narrow(b200::ComplexMatrix&& M, std::true_type) { return std::move(M); }
    static ComplexMatrix narrow(b200::ComplexMatrix&& M, std::false_type)
    {
        ComplexMatrix res(M.rows(), M.cols());
        for (Index q = 0; q < M.rows() * M.cols(); q++)
            res.data()[q] = Complex(static_cast<RealScalar>(M.data()[q].real()), static_cast<RealScalar>(M.data()[q].imag()));
        return res;
    }
};

}  // namespace Spectra
#endif

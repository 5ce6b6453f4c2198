// B200 shim of Spectra/MatOp/SparseGenMatProd.h:29-104: y = A x for a general real sparse matrix.
#ifndef SPECTRA_B200_SPARSE_GEN_MAT_PROD_H
#define SPECTRA_B200_SPARSE_GEN_MAT_PROD_H

#include <complex>
#include <type_traits>
#include <vector>

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
        if (mat.rows() != mat.cols())
            throw std::invalid_argument("SparseGenMatProd: matrix must be square");
        if (mat.isCompressed())
            create_any(mat.rows(), mat.outerIndexPtr(), mat.innerIndexPtr(), mat.valuePtr(), Flags == Eigen::RowMajor, SB200_GENERAL);
        else
        {
            m_packed.pack(mat);  // uncompressed mode, as in the reference's README example (README.md:150-160)
            create_any(mat.rows(), m_packed.outer.data(), m_packed.inner.data(), m_packed.values.data(), Flags == Eigen::RowMajor, SB200_GENERAL);
        }
    }

private:
    b200::PackedCopy<StorageIndex, Scalar> m_packed;
#endif
};

// Scalar = std::complex<double> (the reference's wrapper is templated on Scalar; test/ComplexEigs.cpp:77-79 uses this instantiation):
// a general complex sparse matrix behind the complex CSR SpMV kernel, for GenEigsSolver.  (device-verified in round 2.)
template <int Flags, typename StorageIndex>
class SparseGenMatProd<std::complex<double>, Flags, StorageIndex> : public b200::DeviceOpTag
{
    static_assert(std::is_integral<StorageIndex>::value && (sizeof(StorageIndex) == 4 || sizeof(StorageIndex) == 8), "StorageIndex must be a 32- or 64-bit integer");
    sb200_op* m_op = nullptr;
    Index m_n = 0;
    std::vector<int32_t> m_inner32;

    void create(Index n, const StorageIndex* outer, const StorageIndex* inner, const std::complex<double>* values)
    {
        m_n = n;
        const int32_t* in32 = reinterpret_cast<const int32_t*>(inner);
        if (sizeof(StorageIndex) == 8)
        {
            const int64_t nnz = static_cast<int64_t>(outer[n]);
            m_inner32.resize(static_cast<size_t>(nnz));
            for (int64_t p = 0; p < nnz; p++)
                m_inner32[static_cast<size_t>(p)] = static_cast<int32_t>(inner[p]);
            in32 = m_inner32.data();
        }
        b200::check(sb200_op_create_sparse_herm(n, outer, sizeof(StorageIndex) == 8 ? 1 : 0, in32, reinterpret_cast<const double*>(values),
                                                Flags == SPECTRA_B200_ROWMAJOR ? SB200_ROW_MAJOR : SB200_COL_MAJOR, SB200_GENERAL, &m_op));
    }

public:
    using Scalar = std::complex<double>;

    SparseGenMatProd(Index n, const StorageIndex* outer, const StorageIndex* inner, const Scalar* values) { create(n, outer, inner, values); }
#ifdef SPECTRA_B200_HAS_EIGEN
    explicit SparseGenMatProd(const Eigen::SparseMatrix<Scalar, Flags, StorageIndex>& mat)
    {
        if (mat.rows() != mat.cols())
            throw std::invalid_argument("SparseGenMatProd: matrix must be square");
        if (mat.isCompressed())
            create(mat.rows(), mat.outerIndexPtr(), mat.innerIndexPtr(), mat.valuePtr());
        else
        {
            b200::PackedCopy<StorageIndex, Scalar> packed;  // the arrays are uploaded at construction and not referenced afterwards
            packed.pack(mat);
            create(mat.rows(), packed.outer.data(), packed.inner.data(), packed.values.data());
        }
    }
#endif
    SparseGenMatProd(const SparseGenMatProd&) = delete;
    SparseGenMatProd& operator=(const SparseGenMatProd&) = delete;
    ~SparseGenMatProd()
    {
        if (m_op)
            sb200_op_destroy(m_op);
    }
    Index rows() const { return m_n; }
    Index cols() const { return m_n; }
    sb200_op* handle() const { return m_op; }
    void perform_op(const Scalar* x_in, Scalar* y_out) const
    {
        b200::check(sb200_op_perform_op(m_op, reinterpret_cast<const double*>(x_in), reinterpret_cast<double*>(y_out)));
    }
};

}  // namespace Spectra
#endif

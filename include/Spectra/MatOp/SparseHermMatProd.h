// B200 shim of Spectra/MatOp/SparseHermMatProd.h:21-89: y = A x for a complex Hermitian sparse matrix of which only the `Uplo`
// triangle is read (selfadjointView<Uplo>: mirrored conjugated, diagonal taken as real), executed by the sm_100a complex CSR
// SpMV kernel.  Scalar = std::complex<double>.  (device-verified in round 2.)
#ifndef SPECTRA_B200_SPARSE_HERM_MAT_PROD_H
#define SPECTRA_B200_SPARSE_HERM_MAT_PROD_H

#include <complex>
#include <type_traits>
#include <vector>

#include "SparseSymMatProd.h"

namespace Spectra {

template <typename Scalar_, int Uplo = SPECTRA_B200_LOWER, int Flags = SPECTRA_B200_COLMAJOR, typename StorageIndex = int>
class SparseHermMatProd : public b200::DeviceOpTag
{
    static_assert(std::is_same<Scalar_, std::complex<double>>::value, "the B200 path implements Scalar = std::complex<double>");
    static_assert(std::is_integral<StorageIndex>::value && (sizeof(StorageIndex) == 4 || sizeof(StorageIndex) == 8), "StorageIndex must be a 32- or 64-bit integer");

    sb200_op* m_op = nullptr;
    Index m_n = 0;
    std::vector<int32_t> m_inner32;

    void create(Index n, const StorageIndex* outer, const StorageIndex* inner, const Scalar_* values)
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
        // std::complex<double> is layout-compatible with double[2] (re, im)
        b200::check(sb200_op_create_sparse_herm(n, outer, sizeof(StorageIndex) == 8 ? 1 : 0, in32, reinterpret_cast<const double*>(values),
                                                Flags == SPECTRA_B200_ROWMAJOR ? SB200_ROW_MAJOR : SB200_COL_MAJOR,
                                                Uplo == SPECTRA_B200_LOWER ? SB200_HERM_LOWER : SB200_HERM_UPPER, &m_op));
    }

public:
    using Scalar = Scalar_;

    // Raw compressed arrays with Eigen's layout: outer[n + 1], inner[nnz], values[nnz] (uploaded once).
    SparseHermMatProd(Index n, const StorageIndex* outer, const StorageIndex* inner, const Scalar* values) { create(n, outer, inner, values); }
#ifdef SPECTRA_B200_HAS_EIGEN
    // Same constructor as the reference (SparseHermMatProd.h:46-54).
    explicit SparseHermMatProd(const Eigen::SparseMatrix<Scalar, Flags, StorageIndex>& mat)
    {
        if (mat.rows() != mat.cols())
            throw std::invalid_argument("SparseHermMatProd: matrix must be square");
        if (mat.isCompressed())
            create(mat.rows(), mat.outerIndexPtr(), mat.innerIndexPtr(), mat.valuePtr());
        else
        {
            b200::PackedCopy<StorageIndex, Scalar> packed;  // uncompressed mode: packed once; the arrays are uploaded at construction
            packed.pack(mat);
            create(mat.rows(), packed.outer.data(), packed.inner.data(), packed.values.data());
        }
    }
#endif
    SparseHermMatProd(const SparseHermMatProd&) = delete;
    SparseHermMatProd& operator=(const SparseHermMatProd&) = delete;
    ~SparseHermMatProd()
    {
        if (m_op)
            sb200_op_destroy(m_op);
    }

    Index rows() const { return m_n; }
    Index cols() const { return m_n; }
    sb200_op* handle() const { return m_op; }

    // y_out = A * x_in, host pointers (SparseHermMatProd.h:83-88)
    void perform_op(const Scalar* x_in, Scalar* y_out) const
    {
        b200::check(sb200_op_perform_op(m_op, reinterpret_cast<const double*>(x_in), reinterpret_cast<double*>(y_out)));
    }
};

}  // namespace Spectra
#endif

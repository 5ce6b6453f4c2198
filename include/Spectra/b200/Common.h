// Shared pieces of the header-only C++ shim over the C ABI (include/spectra_b200.h).
//
// The shim keeps the reference's class and method names so that user code written against
// yixuan/spectra compiles unchanged for the supported path:
//     SparseSymMatProd<double> op(...);  SymEigsSolver<SparseSymMatProd<double>> eigs(op, nev, ncv);
//     eigs.init();  eigs.compute(SortRule::LargestAlge);  eigs.info();  eigs.eigenvalues();  eigs.eigenvectors();
// When Eigen is available (<Eigen/Core> on the include path) results are Eigen vectors / matrices and
// the operator constructors accept Eigen::SparseMatrix; otherwise the light containers below are used
// (same element access: v[i], M(i, j), .data(), .size(), .rows(), .cols()).
#ifndef SPECTRA_B200_COMMON_H
#define SPECTRA_B200_COMMON_H

#include <complex>
#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include "../../spectra_b200.h"

#if defined(__has_include)
#if __has_include(<Eigen/Core>) && __has_include(<Eigen/SparseCore>) && !defined(SPECTRA_B200_NO_EIGEN)
#define SPECTRA_B200_HAS_EIGEN 1
#include <Eigen/Core>
#include <Eigen/SparseCore>
#endif
#endif

namespace Spectra {

using Index = std::ptrdiff_t;

namespace b200 {

// status -> the exception type the reference throws
inline void check(int status)
{
    if (status == SB200_OK)
        return;
    const std::string msg = sb200_last_error();
    switch (status)
    {
        case SB200_INVALID_ARGUMENT: throw std::invalid_argument(msg);
        case SB200_LOGIC: throw std::logic_error(msg);
        default: throw std::runtime_error(msg);
    }
}

#ifndef SPECTRA_B200_HAS_EIGEN
template <typename T>
class VectorT
{
    std::vector<T> m_v;

public:
    VectorT() {}
    explicit VectorT(Index n) : m_v(static_cast<size_t>(n)) {}
    Index size() const { return static_cast<Index>(m_v.size()); }
    Index rows() const { return size(); }
    Index cols() const { return 1; }
    T& operator[](Index i) { return m_v[static_cast<size_t>(i)]; }
    const T& operator[](Index i) const { return m_v[static_cast<size_t>(i)]; }
    T& operator()(Index i) { return m_v[static_cast<size_t>(i)]; }
    const T& operator()(Index i) const { return m_v[static_cast<size_t>(i)]; }
    T* data() { return m_v.data(); }
    const T* data() const { return m_v.data(); }
    void resize(Index n) { m_v.resize(static_cast<size_t>(n)); }
};

// column-major dense matrix, element (i, j) at data()[i + j * rows()]
template <typename T>
class MatrixT
{
    Index m_r = 0, m_c = 0;
    std::vector<T> m_v;

public:
    MatrixT() {}
    MatrixT(Index r, Index c) : m_r(r), m_c(c), m_v(static_cast<size_t>(r * c)) {}
    Index rows() const { return m_r; }
    Index cols() const { return m_c; }
    Index size() const { return m_r * m_c; }
    T& operator()(Index i, Index j) { return m_v[static_cast<size_t>(i + j * m_r)]; }
    const T& operator()(Index i, Index j) const { return m_v[static_cast<size_t>(i + j * m_r)]; }
    T* data() { return m_v.data(); }
    const T* data() const { return m_v.data(); }
    void resize(Index r, Index c)
    {
        m_r = r;
        m_c = c;
        m_v.resize(static_cast<size_t>(r * c));
    }
    // keep the first c columns (column-major storage makes this a truncation)
    void conservative_resize_cols(Index c)
    {
        m_c = c;
        m_v.resize(static_cast<size_t>(m_r * c));
    }
};
template <typename T>
using VectorOf = VectorT<T>;
template <typename T>
using MatrixOf = MatrixT<T>;
template <typename T>
inline void shrink_cols(MatrixT<T>& M, Index c)
{
    M.conservative_resize_cols(c);
}
#else
template <typename T>
using VectorOf = Eigen::Matrix<T, Eigen::Dynamic, 1>;
template <typename T>
using MatrixOf = Eigen::Matrix<T, Eigen::Dynamic, Eigen::Dynamic>;
template <typename T>
inline void shrink_cols(MatrixOf<T>& M, Index c)
{
    M.conservativeResize(M.rows(), c);
}
#endif
using Vector = VectorOf<double>;
using Matrix = MatrixOf<double>;
using ComplexVector = VectorOf<std::complex<double>>;
using ComplexMatrix = MatrixOf<std::complex<double>>;

// Scalar = float (the reference's wrappers are templated on Scalar, SparseSymMatProd.h:30): the shim keeps the user's float
// storage and float results, the device path computes in fp64 -- values are widened once at upload, vectors at the boundary.
// (The path is HBM-bound; fp32 device storage would halve its bytes and is not implemented.)
template <typename T>
struct IsSupportedScalar : std::integral_constant<bool, std::is_same<T, double>::value || std::is_same<T, float>::value>
{
};
inline const double* widen(const double* p, Index, std::vector<double>&) { return p; }
inline const double* widen(const float* p, Index n, std::vector<double>& buf)
{
    buf.resize(static_cast<size_t>(n));
    for (Index i = 0; i < n; i++)
        buf[static_cast<size_t>(i)] = static_cast<double>(p[i]);
    return buf.data();
}

#ifdef SPECTRA_B200_HAS_EIGEN
// An Eigen::SparseMatrix in UNCOMPRESSED mode (after reserve() / insert(), e.g. the reference's README example, README.md:146-178): the
// reference binds its matrix through Eigen::Ref<const SparseMatrix>, which accepts either mode, so the wrappers must too.  The inner
// vectors (outer[k] .. outer[k] + innerNonZeros[k]) are packed once into compressed arrays owned by the wrapper.
template <typename StorageIndex, typename Scalar>
struct PackedCopy
{
    std::vector<StorageIndex> outer, inner;
    std::vector<Scalar> values;
    template <typename SpMat>
    void pack(const SpMat& mat)
    {
        const Index no = mat.outerSize();
        const StorageIndex* o = mat.outerIndexPtr();
        const StorageIndex* c = mat.innerNonZeroPtr();
        const StorageIndex* in = mat.innerIndexPtr();
        const Scalar* v = mat.valuePtr();
        outer.assign(static_cast<size_t>(no + 1), StorageIndex(0));
        inner.clear();
        values.clear();
        for (Index k = 0; k < no; k++)
        {
            for (StorageIndex p = o[k]; p < o[k] + c[k]; p++)
            {
                inner.push_back(in[p]);
                values.push_back(v[p]);
            }
            outer[static_cast<size_t>(k + 1)] = static_cast<StorageIndex>(inner.size());
        }
    }
};
#endif

// Marks operator wrappers that own a device-resident sb200_op (exposed through handle()); any other OpType is treated as a
// user-defined host operator and wrapped in a callback adapter.
struct DeviceOpTag
{
};

// Device-resident sparse operator shared by SparseSymMatProd / SparseGenMatProd.
// The user's compressed arrays must outlive the operator (the reference holds an Eigen::Ref to the
// user's matrix, SparseSymMatProd.h:46-48); they are uploaded once at construction.
class SparseOpBase : public DeviceOpTag
{
protected:
    sb200_op* m_op = nullptr;
    Index m_n = 0;
    const void* m_outer = nullptr;
    bool m_outer64 = false;
    const int32_t* m_inner = nullptr;
    const double* m_values = nullptr;
    bool m_row_major = false;

    void create(Index n, const void* outer, bool outer64, const int32_t* inner, const double* values, bool row_major, int mode, bool shift_solve = false)
    {
        m_n = n;
        m_outer = outer;
        m_outer64 = outer64;
        m_inner = inner;
        m_values = values;
        m_row_major = row_major;
        if (shift_solve)
            check(sb200_op_create_shift_solve(n, outer, outer64 ? 1 : 0, inner, values, row_major ? SB200_ROW_MAJOR : SB200_COL_MAJOR, mode, &m_op));
        else
            check(sb200_op_create_sparse(n, outer, outer64 ? 1 : 0, inner, values, row_major ? SB200_ROW_MAJOR : SB200_COL_MAJOR, mode, nullptr, &m_op));
    }
    // StorageIndex = int (Eigen's default) is passed through; 64-bit StorageIndex keeps its 64-bit outer offsets and has its
    // inner indices narrowed once (the device CSR addresses n < 2^31 columns with 32-bit ids)
    std::vector<int32_t> m_inner32;
    std::vector<double> m_values64;  // Scalar = float: the widened copy of the user's values
    template <typename StorageIndex>
    void create_any(Index n, const StorageIndex* outer, const StorageIndex* inner, const float* values, bool row_major, int mode, bool shift_solve = false)
    {
        create_any(n, outer, inner, widen(values, static_cast<Index>(outer[n]), m_values64), row_major, mode, shift_solve);
    }
    template <typename StorageIndex>
    void create_any(Index n, const StorageIndex* outer, const StorageIndex* inner, const double* values, bool row_major, int mode, bool shift_solve = false)
    {
        static_assert(std::is_integral<StorageIndex>::value && (sizeof(StorageIndex) == 4 || sizeof(StorageIndex) == 8), "StorageIndex must be a 32- or 64-bit integer");
        if (sizeof(StorageIndex) == 4)
        {
            create(n, outer, false, reinterpret_cast<const int32_t*>(inner), values, row_major, mode, shift_solve);
            return;
        }
        if (n >= (Index(1) << 31))
            throw std::invalid_argument("matrix order must be below 2^31");
        const int64_t nnz = static_cast<int64_t>(outer[n]);
        m_inner32.resize(static_cast<size_t>(nnz));
        for (int64_t p = 0; p < nnz; p++)
            m_inner32[static_cast<size_t>(p)] = static_cast<int32_t>(inner[p]);
        create(n, outer, true, m_inner32.data(), values, row_major, mode, shift_solve);
    }
    // 64-bit offsets with 32-bit inner indices (the dense wrappers: n * n entries can pass 2^31 while every index stays below n)
    void create_any(Index n, const int64_t* outer, const int32_t* inner, const double* values, bool row_major, int mode, bool shift_solve = false)
    {
        if (n >= (Index(1) << 31))
            throw std::invalid_argument("matrix order must be below 2^31");
        create(n, outer, true, inner, values, row_major, mode, shift_solve);
    }
    void create_any(Index n, const int64_t* outer, const int32_t* inner, const float* values, bool row_major, int mode, bool shift_solve = false)
    {
        create_any(n, outer, inner, widen(values, static_cast<Index>(outer[n]), m_values64), row_major, mode, shift_solve);
    }
    int64_t outer_at(Index i) const
    {
        return m_outer64 ? static_cast<const int64_t*>(m_outer)[i] : static_cast<int64_t>(static_cast<const int32_t*>(m_outer)[i]);
    }

public:
    SparseOpBase() {}
    SparseOpBase(const SparseOpBase&) = delete;
    SparseOpBase& operator=(const SparseOpBase&) = delete;
    SparseOpBase(SparseOpBase&& o) noexcept { *this = std::move(o); }
    SparseOpBase& operator=(SparseOpBase&& o) noexcept
    {
        if (this != &o)
        {
            if (m_op)
                sb200_op_destroy(m_op);
            m_op = o.m_op;
            m_n = o.m_n;
            m_outer = o.m_outer;
            m_outer64 = o.m_outer64;
            m_inner = o.m_inner;
            m_values = o.m_values;
            m_row_major = o.m_row_major;
            const bool own_inner = !o.m_inner32.empty();
            m_inner32 = std::move(o.m_inner32);
            if (own_inner)
                m_inner = m_inner32.data();
            const bool own_values = !o.m_values64.empty();
            m_values64 = std::move(o.m_values64);
            if (own_values)
                m_values = m_values64.data();
            o.m_op = nullptr;
        }
        return *this;
    }
    ~SparseOpBase()
    {
        if (m_op)
            sb200_op_destroy(m_op);
    }

    Index rows() const { return m_n; }
    Index cols() const { return m_n; }
    sb200_op* handle() const { return m_op; }

    // y_out = A * x_in, host pointers (SparseSymMatProd.h:83-88 / SparseGenMatProd.h:82-87)
    void perform_op(const double* x_in, double* y_out) const { check(sb200_op_perform_op(m_op, x_in, y_out)); }
    void perform_op(const float* x_in, float* y_out) const
    {
        std::vector<double> x(static_cast<size_t>(m_n)), y(static_cast<size_t>(m_n));
        for (Index i = 0; i < m_n; i++)
            x[static_cast<size_t>(i)] = x_in[i];
        check(sb200_op_perform_op(m_op, x.data(), y.data()));
        for (Index i = 0; i < m_n; i++)
            y_out[i] = static_cast<float>(y[static_cast<size_t>(i)]);
    }

    // operator*(Matrix) (SparseSymMatProd.h:93-96)
    Matrix operator*(const Matrix& mat_in) const
    {
        Matrix res(m_n, mat_in.cols());
        check(sb200_op_apply_matrix(m_op, mat_in.data(), mat_in.cols(), res.data()));
        return res;
    }
    MatrixOf<float> operator*(const MatrixOf<float>& mat_in) const
    {
        MatrixOf<float> res(m_n, mat_in.cols());
        for (Index c = 0; c < mat_in.cols(); c++)
            perform_op(mat_in.data() + c * m_n, res.data() + c * m_n);
        return res;
    }

    // operator()(i, j): the stored coefficient of the user's matrix (SparseSymMatProd.h:101-104)
    double operator()(Index i, Index j) const
    {
        const Index o = m_row_major ? i : j, k = m_row_major ? j : i;
        for (int64_t p = outer_at(o); p < outer_at(o + 1); p++)
            if (m_inner[p] == k)
                return m_values[p];
        return 0.0;
    }
};

// Adapter for user-defined operators (the OpType concept of the reference, SymEigsSolver.h:99-114): any class with
// rows() and perform_op(const double* x_in, double* y_out) const.  The solver keeps the Krylov iteration on the GPU and
// calls back into the user's host code once per matrix operation.
template <typename OpType>
class HostOpAdapter
{
    sb200_op* m_op = nullptr;
    const OpType* m_user;
    Index m_rows = 0;
    mutable std::vector<typename OpType::Scalar> m_x, m_y;  // Scalar = float: narrowed operand / result of the user's perform_op
    static void call(const OpType* op, const double* x, double* y, const HostOpAdapter*, std::true_type) { op->perform_op(x, y); }
    static void call(const OpType* op, const double* x, double* y, const HostOpAdapter* self, std::false_type)
    {
        using S = typename OpType::Scalar;
        const size_t n = static_cast<size_t>(self->m_rows);
        self->m_x.resize(n);
        self->m_y.resize(n);
        for (size_t i = 0; i < n; i++)
            self->m_x[i] = static_cast<S>(x[i]);
        op->perform_op(self->m_x.data(), self->m_y.data());
        for (size_t i = 0; i < n; i++)
            y[i] = static_cast<double>(self->m_y[i]);
    }
    static void trampoline(const double* x, double* y, void* self)
    {
        const HostOpAdapter* a = static_cast<const HostOpAdapter*>(self);
        call(a->m_user, x, y, a, std::is_same<typename OpType::Scalar, double>());
    }

public:
    explicit HostOpAdapter(const OpType& op) : m_user(&op), m_rows(static_cast<Index>(op.rows()))
    {
        static_assert(IsSupportedScalar<typename OpType::Scalar>::value, "user-defined operators must use Scalar = double or float");
        check(sb200_op_create_callback(static_cast<int64_t>(op.rows()), &HostOpAdapter::trampoline, this, &m_op));
    }
    HostOpAdapter(const HostOpAdapter&) = delete;
    HostOpAdapter& operator=(const HostOpAdapter&) = delete;
    ~HostOpAdapter()
    {
        if (m_op)
            sb200_op_destroy(m_op);
    }
    sb200_op* handle() const { return m_op; }
};

// The same for user-defined COMPLEX operators (OpType::Scalar = std::complex<double>, HermEigsSolver): vectors cross the C ABI as
// interleaved (re, im) doubles, which is the memory layout of std::complex<double>.
template <typename OpType>
class HostOpAdapterZ
{
    sb200_op* m_op = nullptr;
    const OpType* m_user;
    static void trampoline(const double* x, double* y, void* self)
    {
        static_cast<const OpType*>(self)->perform_op(reinterpret_cast<const std::complex<double>*>(x), reinterpret_cast<std::complex<double>*>(y));
    }

public:
    explicit HostOpAdapterZ(const OpType& op) : m_user(&op)
    {
        static_assert(std::is_same<typename OpType::Scalar, std::complex<double>>::value, "complex user-defined operators must use Scalar = std::complex<double>");
        check(sb200_op_create_callback_z(static_cast<int64_t>(op.rows()), &HostOpAdapterZ::trampoline, const_cast<OpType*>(m_user), &m_op));
    }
    HostOpAdapterZ(const HostOpAdapterZ&) = delete;
    HostOpAdapterZ& operator=(const HostOpAdapterZ&) = delete;
    ~HostOpAdapterZ()
    {
        if (m_op)
            sb200_op_destroy(m_op);
    }
    sb200_op* handle() const { return m_op; }
};
template <typename OpType, bool IsDevice = std::is_base_of<DeviceOpTag, OpType>::value>
struct OpBindingZ
{
    explicit OpBindingZ(OpType& op) : m_h(op.handle()) {}
    sb200_op* handle() const { return m_h; }
    sb200_op* m_h;
};
template <typename OpType>
struct OpBindingZ<OpType, false>
{
    explicit OpBindingZ(OpType& op) : m_adapter(op) {}
    sb200_op* handle() const { return m_adapter.handle(); }
    HostOpAdapterZ<OpType> m_adapter;
};

// Picks the device handle of an operator: device-resident sparse wrappers expose it directly, anything else is wrapped.
template <typename OpType, bool IsDevice = std::is_base_of<DeviceOpTag, OpType>::value>
struct OpBinding
{
    explicit OpBinding(OpType& op) : m_h(op.handle()) {}
    sb200_op* handle() const { return m_h; }
    sb200_op* m_h;
};
template <typename OpType>
struct OpBinding<OpType, false>
{
    explicit OpBinding(OpType& op) : m_adapter(op) {}
    sb200_op* handle() const { return m_adapter.handle(); }
    HostOpAdapter<OpType> m_adapter;
};

}  // namespace b200
}  // namespace Spectra
#endif

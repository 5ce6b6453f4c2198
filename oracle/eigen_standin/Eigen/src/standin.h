// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product path.
//
// A stand-in for the subset of the Eigen 3.4 API that yixuan/spectra's Lanczos / Arnoldi path uses,
// so that the reference's OWN headers (/root/reference/include/Spectra, compiled where they lie,
// never copied) build into oracle/_ref/libspectra_ref.so without Eigen, which this image does not
// have (the reference fetches Eigen 3.4 through CPM at configure time, CMakeLists.txt:25-38).
//
// What this is: eager (no expression templates) column-major dense matrices, strided views for
// Map / Ref / Block, the Jacobi / Householder primitives restated from Eigen 3.4.0's published
// definitions (Jacobi.h makeGivens, Householder.h makeHouseholder, MathFunctions.h hypot, ComplexSchur.h), and a
// compressed sparse matrix with the two products the reference calls.  Everything the reference
// decides -- every branch, shift, deflation test, restart rule, iteration count -- is the
// reference's own compiled code; only the BLAS-level loops (dot, axpy, gemv, small gemm) and their
// summation order are this file's.  Eigen itself does not specify that order either (it depends
// on the SIMD width of the build), so the reference's results are defined up to it.
//
// Not a general Eigen replacement: unsupported calls fail to compile.
#pragma once

#include <algorithm>
#include <cassert>
#include <cmath>
#include <complex>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <limits>
#include <ostream>
#include <stdexcept>
#include <type_traits>
#include <utility>
#include <vector>

namespace Eigen {

using Index = std::ptrdiff_t;
enum : int { Dynamic = -1 };
enum : int { ColMajor = 0, RowMajor = 1 };
enum : int { Lower = 1, Upper = 2, StrictlyLower = 9, StrictlyUpper = 10, UnitLower = 5, UnitUpper = 6 };
enum ComputationInfo { Success = 0, NumericalIssue = 1, NoConvergence = 2, InvalidInput = 3 };

// ------------------------------------------------------------------------------------------------
// NumTraits / numext
// ------------------------------------------------------------------------------------------------
template <typename T>
struct NumTraits
{
    using Real = T;
    enum { IsComplex = 0 };
    static constexpr T epsilon() { return std::numeric_limits<T>::epsilon(); }
    static constexpr T lowest() { return std::numeric_limits<T>::lowest(); }
    static constexpr T highest() { return (std::numeric_limits<T>::max)(); }
};
template <typename T>
struct NumTraits<std::complex<T>>
{
    using Real = T;
    enum { IsComplex = 1 };
    static constexpr T epsilon() { return std::numeric_limits<T>::epsilon(); }
};

namespace numext {
template <typename T>
using numeric_limits = std::numeric_limits<T>;

template <typename T>
inline T real(const T& x) { return x; }
template <typename T>
inline T real(const std::complex<T>& x) { return x.real(); }
template <typename T>
inline T imag(const T&) { return T(0); }
template <typename T>
inline T imag(const std::complex<T>& x) { return x.imag(); }
template <typename T>
inline T conj(const T& x) { return x; }
template <typename T>
inline std::complex<T> conj(const std::complex<T>& x) { return std::conj(x); }
template <typename T>
inline T abs2(const T& x) { return x * x; }
template <typename T>
inline T abs2(const std::complex<T>& x) { return x.real() * x.real() + x.imag() * x.imag(); }
template <typename T>
inline T norm1(const T& x) { return std::abs(x); }
template <typename T>
inline T norm1(const std::complex<T>& x) { return std::abs(x.real()) + std::abs(x.imag()); }
template <typename T>
inline T maxi(const T& a, const T& b) { return a < b ? b : a; }
template <typename T>
inline T mini(const T& a, const T& b) { return b < a ? b : a; }
template <typename T>
inline T& real_ref(std::complex<T>& x) { return reinterpret_cast<T*>(&x)[0]; }
template <typename T>
inline T& real_ref(T& x) { return x; }

// Eigen 3.4.0 MathFunctions.h, positive_real_hypot
template <typename T>
inline T hypot(const T& xx, const T& yy)
{
    T x = std::abs(xx), y = std::abs(yy);
    if (std::isinf(x) || std::isinf(y))
        return std::numeric_limits<T>::infinity();
    if (std::isnan(x) || std::isnan(y))
        return std::numeric_limits<T>::quiet_NaN();
    T p = (std::max)(x, y);
    if (p == T(0))
        return T(0);
    T qp = (std::min)(y, x) / p;
    return p * std::sqrt(T(1) + qp * qp);
}
}  // namespace numext

namespace internal {
template <typename T>
struct is_scalar : std::is_arithmetic<T>
{};
template <typename T>
struct is_scalar<std::complex<T>> : std::true_type
{};

template <typename A, typename B>
using prod_t = decltype(std::declval<A>() * std::declval<B>());

// Random(): uniform in [-1, 1] from std::rand(), as Eigen's default does (the reference's tests seed it with std::srand(123))
template <typename T>
struct random_scalar
{
    static T run() { return T(2.0 * double(std::rand()) / double(RAND_MAX) - 1.0); }
};
template <typename T>
struct random_scalar<std::complex<T>>
{
    static std::complex<T> run()
    {
        const T re = random_scalar<T>::run();
        const T im = random_scalar<T>::run();
        return std::complex<T>(re, im);
    }
};

// "Packets" for UpperHessenbergSchur.h's hand-vectorised Householder update (:204-270): two lanes,
// element-wise arithmetic, so the results equal the scalar loop's.
template <typename T>
struct Packet2
{
    T v[2];
};
template <typename T>
struct packet_traits
{
    using type = Packet2<T>;
    enum { size = 2 };
};
template <typename P, typename T>
inline P ploadu(const T* p) { return P{{p[0], p[1]}}; }
template <typename T>
inline void pstoreu(T* p, const Packet2<T>& a) { p[0] = a.v[0]; p[1] = a.v[1]; }
template <typename P, typename T>
inline P pset1(const T& a) { return P{{a, a}}; }
template <typename T>
inline Packet2<T> padd(const Packet2<T>& a, const Packet2<T>& b) { return {{a.v[0] + b.v[0], a.v[1] + b.v[1]}}; }
template <typename T>
inline Packet2<T> psub(const Packet2<T>& a, const Packet2<T>& b) { return {{a.v[0] - b.v[0], a.v[1] - b.v[1]}}; }
template <typename T>
inline Packet2<T> pmul(const Packet2<T>& a, const Packet2<T>& b) { return {{a.v[0] * b.v[0], a.v[1] * b.v[1]}}; }
}  // namespace internal

// ------------------------------------------------------------------------------------------------
// Forward declarations
// ------------------------------------------------------------------------------------------------
template <typename Derived>
class MatrixBase;
template <typename S, int R, int C, int Opt = ColMajor>
class Matrix;
template <typename S, int R = Dynamic, int C = Dynamic>
class View;
template <typename S, int R, int C>
class Array;
template <typename S>
class ArrayRef;
template <typename S>
class JacobiRotation;
template <typename S>
class AdjointView;
template <typename S, int Uplo>
class DenseSelfAdjointView;
template <typename S>
class DiagonalWrapper;

namespace internal {
template <typename T>
struct traits;
template <typename S, int R, int C, int Opt>
struct traits<Matrix<S, R, C, Opt>>
{
    using Scalar = S;
    enum { Rows = R, Cols = C };
};
template <typename S, int R, int C>
struct traits<View<S, R, C>>
{
    using Scalar = S;
    enum { Rows = R, Cols = C };
};
}  // namespace internal

template <typename S>
class JacobiRotation
{
    S m_c, m_s;

public:
    JacobiRotation() : m_c(1), m_s(0) {}
    JacobiRotation(const S& c, const S& s) : m_c(c), m_s(s) {}
    S& c() { return m_c; }
    S c() const { return m_c; }
    S& s() { return m_s; }
    S s() const { return m_s; }
    JacobiRotation adjoint() const { return JacobiRotation(numext::conj(m_c), -m_s); }
    JacobiRotation transpose() const { return JacobiRotation(m_c, -numext::conj(m_s)); }

    void makeGivens(const S& p, const S& q, S* r = nullptr) { make_givens_(p, q, r, std::integral_constant<bool, NumTraits<S>::IsComplex != 0>()); }

private:
    // Eigen 3.4.0 Jacobi.h, makeGivens(p, q, r, true_type): complex scalars (used by ComplexSchur)
    void make_givens_(const S& p, const S& q, S* r, std::true_type)
    {
        using std::abs;
        using std::sqrt;
        using RealScalar = typename NumTraits<S>::Real;
        if (q == S(0))
        {
            m_c = numext::real(p) < 0 ? S(-1) : S(1);
            m_s = 0;
            if (r)
                *r = m_c * p;
        }
        else if (p == S(0))
        {
            m_c = 0;
            m_s = -q / abs(q);
            if (r)
                *r = abs(q);
        }
        else
        {
            RealScalar p1 = numext::norm1(p);
            RealScalar q1 = numext::norm1(q);
            if (p1 >= q1)
            {
                S ps = p / p1;
                RealScalar p2 = numext::abs2(ps);
                S qs = q / p1;
                RealScalar q2 = numext::abs2(qs);
                RealScalar u = sqrt(RealScalar(1) + q2 / p2);
                if (numext::real(p) < RealScalar(0))
                    u = -u;
                m_c = S(1) / u;
                m_s = -qs * numext::conj(ps) * (m_c / p2);
                if (r)
                    *r = p * u;
            }
            else
            {
                S ps = p / q1;
                RealScalar p2 = numext::abs2(ps);
                S qs = q / q1;
                RealScalar q2 = numext::abs2(qs);
                RealScalar u = q1 * sqrt(p2 + q2);
                if (numext::real(p) < RealScalar(0))
                    u = -u;
                p1 = abs(p);
                ps = p / p1;
                m_c = p1 / u;
                m_s = -numext::conj(ps) * (q / u);
                if (r)
                    *r = ps * u;
            }
        }
    }

    // Eigen 3.4.0 Jacobi.h, makeGivens(p, q, r, false_type): real scalars
    void make_givens_(const S& p, const S& q, S* r, std::false_type)
    {
        using std::abs;
        using std::sqrt;
        if (q == S(0))
        {
            m_c = p < S(0) ? S(-1) : S(1);
            m_s = S(0);
            if (r)
                *r = abs(p);
        }
        else if (p == S(0))
        {
            m_c = S(0);
            m_s = q < S(0) ? S(1) : S(-1);
            if (r)
                *r = abs(q);
        }
        else if (abs(p) > abs(q))
        {
            S t = q / p;
            S u = sqrt(S(1) + numext::abs2(t));
            if (p < S(0))
                u = -u;
            m_c = S(1) / u;
            m_s = -t * m_c;
            if (r)
                *r = p * u;
        }
        else
        {
            S t = p / q;
            S u = sqrt(S(1) + numext::abs2(t));
            if (q < S(0))
                u = -u;
            m_s = -S(1) / u;
            m_c = -t * m_s;
            if (r)
                *r = q * u;
        }
    }
};

// ------------------------------------------------------------------------------------------------
// MatrixBase: every dense object.  Derived provides rows(), cols(), ptr(), rstride(), cstride().
// ------------------------------------------------------------------------------------------------
template <typename Derived>
class MatrixBase
{
public:
    using Scalar = typename internal::traits<Derived>::Scalar;
    using RealScalar = typename NumTraits<Scalar>::Real;
    enum { RowsAtCompileTime = internal::traits<Derived>::Rows, ColsAtCompileTime = internal::traits<Derived>::Cols };
    using PlainObject = Matrix<Scalar, Dynamic, Dynamic, ColMajor>;
    using DynView = View<Scalar, Dynamic, Dynamic>;
    using ColView = View<Scalar, Dynamic, 1>;
    using RowView = View<Scalar, 1, Dynamic>;

    Derived& derived() { return *static_cast<Derived*>(this); }
    const Derived& derived() const { return *static_cast<const Derived*>(this); }

    Index rows() const { return derived().rows_(); }
    Index cols() const { return derived().cols_(); }
    Index size() const { return rows() * cols(); }
    Scalar* ptr_() const { return derived().ptr_impl(); }
    Index rs() const { return derived().rs_impl(); }
    Index cs() const { return derived().cs_impl(); }

    Scalar& coeffRef(Index i, Index j) { return ptr_()[i * rs() + j * cs()]; }
    const Scalar& coeffRef(Index i, Index j) const { return ptr_()[i * rs() + j * cs()]; }
    const Scalar& coeff(Index i, Index j) const { return ptr_()[i * rs() + j * cs()]; }
    Scalar& operator()(Index i, Index j) { return coeffRef(i, j); }
    const Scalar& operator()(Index i, Index j) const { return coeff(i, j); }

    // linear (vector) access
    Index lstride() const { return cols() == 1 ? rs() : cs(); }
    Scalar& coeffRef(Index i) { return ptr_()[i * lstride()]; }
    const Scalar& coeff(Index i) const { return ptr_()[i * lstride()]; }
    Scalar& operator()(Index i) { return coeffRef(i); }
    const Scalar& operator()(Index i) const { return coeff(i); }
    Scalar& operator[](Index i) { return coeffRef(i); }
    const Scalar& operator[](Index i) const { return coeff(i); }

    Scalar* data() { return ptr_(); }
    const Scalar* data() const { return ptr_(); }

    Derived& noalias() { return derived(); }

    // ---- views ----
    DynView block(Index i, Index j, Index r, Index c) const { return DynView(ptr_() + i * rs() + j * cs(), r, c, rs(), cs()); }
    template <int R, int C>
    View<Scalar, R, C> block(Index i, Index j) const
    {
        return View<Scalar, R, C>(ptr_() + i * rs() + j * cs(), R, C, rs(), cs());
    }
    ColView col(Index j) const { return ColView(ptr_() + j * cs(), rows(), 1, rs(), cs()); }
    RowView row(Index i) const { return RowView(ptr_() + i * rs(), 1, cols(), rs(), cs()); }
    DynView leftCols(Index n) const { return block(0, 0, rows(), n); }
    DynView rightCols(Index n) const { return block(0, cols() - n, rows(), n); }
    DynView topRows(Index n) const { return block(0, 0, n, cols()); }
    DynView bottomRows(Index n) const { return block(rows() - n, 0, n, cols()); }
    ColView diagonal(Index k = 0) const
    {
        const Index i0 = k < 0 ? -k : 0, j0 = k > 0 ? k : 0;
        const Index len = (std::max)(Index(0), (std::min)(rows() - i0, cols() - j0));
        return ColView(ptr_() + i0 * rs() + j0 * cs(), len, 1, rs() + cs(), 0);
    }
    // segment / head / tail keep the orientation of the vector they are taken from
    View<Scalar, RowsAtCompileTime == 1 ? 1 : Dynamic, RowsAtCompileTime == 1 ? Dynamic : 1> segment(Index i, Index n) const
    {
        using V = View<Scalar, RowsAtCompileTime == 1 ? 1 : Dynamic, RowsAtCompileTime == 1 ? Dynamic : 1>;
        if (cols() == 1 && RowsAtCompileTime != 1)
            return V(ptr_() + i * rs(), n, 1, rs(), cs());
        return V(ptr_() + i * cs(), 1, n, rs(), cs());
    }
    auto head(Index n) const { return segment(0, n); }
    auto tail(Index n) const { return segment(size() - n, n); }

    AdjointView<Scalar> adjoint() const { return AdjointView<Scalar>(DynView(ptr_(), rows(), cols(), rs(), cs())); }
    Matrix<Scalar, Dynamic, Dynamic> transpose() const  // evaluated on the spot (used by the reference's tests, not by its solvers)
    {
        Matrix<Scalar, Dynamic, Dynamic> res(cols(), rows());
        for (Index j = 0; j < cols(); j++)
            for (Index i = 0; i < rows(); i++)
                res(j, i) = coeff(i, j);
        return res;
    }

    // ---- what the reference's own unit tests (test/*.cpp) use on top of the solver headers ----
    template <int Uplo>
    DenseSelfAdjointView<Scalar, Uplo> selfadjointView() const
    {
        return DenseSelfAdjointView<Scalar, Uplo>(DynView(ptr_(), rows(), cols(), rs(), cs()));
    }
    // triangularView<Mode>(): converts to a plain matrix (zeros outside the triangle); setZero() clears the triangle in place
    template <int Mode>
    struct TriView
    {
        DynView v;
        static bool inside(Index i, Index j)
        {
            return Mode == Upper ? i <= j : (Mode == Lower ? i >= j : (Mode == StrictlyLower ? i > j : (Mode == StrictlyUpper ? i < j : false)));
        }
        operator Matrix<Scalar, Dynamic, Dynamic>() const
        {
            Matrix<Scalar, Dynamic, Dynamic> res(v.rows(), v.cols());
            for (Index j = 0; j < v.cols(); j++)
                for (Index i = 0; i < v.rows(); i++)
                    if (inside(i, j))
                        res(i, j) = v.coeff(i, j);
            return res;
        }
        void setZero()
        {
            for (Index j = 0; j < v.cols(); j++)
                for (Index i = 0; i < v.rows(); i++)
                    if (inside(i, j))
                        v.coeffRef(i, j) = Scalar(0);
        }
    };
    template <int Mode>
    TriView<Mode> triangularView() const { return TriView<Mode>{DynView(ptr_(), rows(), cols(), rs(), cs())}; }
    DiagonalWrapper<Scalar> asDiagonal() const { return DiagonalWrapper<Scalar>(ColView(ptr_(), size(), 1, lstride(), 0)); }
    DynView topLeftCorner(Index r, Index c) const { return block(0, 0, r, c); }
    DynView bottomRightCorner(Index r, Index c) const { return block(rows() - r, cols() - c, r, c); }
    template <int N>
    DynView rightCols() const { return block(0, cols() - N, rows(), N); }
    template <int N>
    DynView leftCols() const { return block(0, 0, rows(), N); }
    Matrix<Scalar, Dynamic, Dynamic> eval() const { return Matrix<Scalar, Dynamic, Dynamic>(derived()); }
    Matrix<Scalar, Dynamic, Dynamic> inverse() const;  // Gauss-Jordan with partial pivoting (tests only)
    template <typename Other>
    bool isApprox(const MatrixBase<Other>& o, const RealScalar& prec = RealScalar(sizeof(RealScalar) == 4 ? 1e-5 : 1e-12)) const  // NumTraits::dummy_precision()
    {
        // Eigen: ||a - b||^2 <= prec^2 min(||a||^2, ||b||^2)
        if (o.rows() != rows() || o.cols() != cols())
            return false;
        RealScalar d(0);
        for (Index j = 0; j < cols(); j++)
            for (Index i = 0; i < rows(); i++)
                d += numext::abs2(coeff(i, j) - o.coeff(i, j));
        return d <= prec * prec * (std::min)(squaredNorm(), o.squaredNorm());
    }
    Scalar maxCoeff(Index* idx) const
    {
        Scalar m = coeff(0);
        *idx = 0;
        for (Index i = 1; i < size(); i++)
            if (coeff(i) > m)
            {
                m = coeff(i);
                *idx = i;
            }
        return m;
    }
    Scalar minCoeff() const
    {
        Scalar m = coeff(0, 0);
        for (Index j = 0; j < cols(); j++)
            for (Index i = 0; i < rows(); i++)
                if (coeff(i, j) < m)
                    m = coeff(i, j);
        return m;
    }
    Scalar trace() const
    {
        Scalar t(0);
        for (Index i = 0; i < (std::min)(rows(), cols()); i++)
            t += coeff(i, i);
        return t;
    }
    Derived& setRandom()
    {
        for (Index j = 0; j < cols(); j++)
            for (Index i = 0; i < rows(); i++)
                coeffRef(i, j) = internal::random_scalar<Scalar>::run();
        return derived();
    }

    // real(): the object itself for real scalars, a real copy for complex ones
    template <typename T = Scalar>
    typename std::enable_if<!NumTraits<T>::IsComplex, const Derived&>::type real() const { return derived(); }
    template <typename T = Scalar>
    typename std::enable_if<NumTraits<T>::IsComplex, Matrix<RealScalar, Dynamic, Dynamic>>::type real() const
    {
        Matrix<RealScalar, Dynamic, Dynamic> res(rows(), cols());
        for (Index j = 0; j < cols(); j++)
            for (Index i = 0; i < rows(); i++)
                res(i, j) = coeff(i, j).real();
        return res;
    }

    // ---- fills ----
    Derived& setZero() { return setConstant(Scalar(0)); }
    Derived& setConstant(const Scalar& v)
    {
        const Index r = rows(), c = cols();
        for (Index j = 0; j < c; j++)
            for (Index i = 0; i < r; i++)
                coeffRef(i, j) = v;
        return derived();
    }
    Derived& setIdentity()
    {
        setZero();
        const Index k = (std::min)(rows(), cols());
        for (Index i = 0; i < k; i++)
            coeffRef(i, i) = Scalar(1);
        return derived();
    }

    // ---- element-wise assignment helpers (shape: equal, or both vectors of equal length) ----
    template <typename Other, typename F>
    void zip_(const MatrixBase<Other>& o, F f)
    {
        const Index r = rows(), c = cols();
        if (o.rows() == r && o.cols() == c)
        {
            if (c == 1 && rs() == 1 && o.rs() == 1)
            {
                Scalar* d = ptr_();
                const auto* s = o.ptr_();
                for (Index i = 0; i < r; i++)
                    f(d[i], s[i]);
                return;
            }
            for (Index j = 0; j < c; j++)
                for (Index i = 0; i < r; i++)
                    f(coeffRef(i, j), o.coeff(i, j));
        }
        else if ((r == 1 || c == 1) && (o.rows() == 1 || o.cols() == 1) && size() == o.size())
        {
            const Index n = size();
            for (Index i = 0; i < n; i++)
                f(coeffRef(i), o.coeff(i));
        }
        else
            throw std::logic_error("Eigen stand-in: shape mismatch in element-wise operation");
    }
    template <typename Other>
    Derived& operator+=(const MatrixBase<Other>& o)
    {
        zip_(o, [](Scalar& a, const typename Other::Scalar& b) { a += b; });
        return derived();
    }
    template <typename Other>
    Derived& operator-=(const MatrixBase<Other>& o)
    {
        zip_(o, [](Scalar& a, const typename Other::Scalar& b) { a -= b; });
        return derived();
    }
    template <typename T, typename = typename std::enable_if<internal::is_scalar<T>::value>::type>
    Derived& operator*=(const T& s)
    {
        const Index r = rows(), c = cols();
        for (Index j = 0; j < c; j++)
            for (Index i = 0; i < r; i++)
                coeffRef(i, j) *= s;
        return derived();
    }
    template <typename T, typename = typename std::enable_if<internal::is_scalar<T>::value>::type>
    Derived& operator/=(const T& s)
    {
        const Index r = rows(), c = cols();
        for (Index j = 0; j < c; j++)
            for (Index i = 0; i < r; i++)
                coeffRef(i, j) /= s;
        return derived();
    }

    template <typename Other>
    void swap(MatrixBase<Other>& o)
    {
        zip_(o, [](Scalar& a, const Scalar& b) { std::swap(a, const_cast<Scalar&>(b)); });
    }
    template <typename Other>
    void swap(MatrixBase<Other>&& o) { swap(o); }

    // ---- reductions ----
    template <typename Other>
    Scalar dot(const MatrixBase<Other>& o) const
    {
        const Index n = size();
        if (o.size() != n)
            throw std::logic_error("Eigen stand-in: dot() of different lengths");
        const Index sa = lstride(), sb = o.lstride();
        const Scalar* a = ptr_();
        const auto* b = o.ptr_();
        // one accumulator, ascending index: the order of the restatement's single-thread dot (oracle/solver.hpp Blas::dot),
        // so that a strict (-ffp-contract=off) build of the restatement can be compared with the reference bit for bit
        Scalar res(0);
        for (Index i = 0; i < n; i++)
            res += numext::conj(a[i * sa]) * b[i * sb];
        return res;
    }
    RealScalar squaredNorm() const
    {
        const Index r = rows(), c = cols();
        RealScalar total(0);
        for (Index j = 0; j < c; j++)
        {
            const Scalar* a = ptr_() + j * cs();
            const Index s = rs();
            for (Index i = 0; i < r; i++)
                total += numext::abs2(a[i * s]);
        }
        return total;
    }
    RealScalar norm() const { return std::sqrt(squaredNorm()); }
    void normalize()
    {
        const RealScalar z = squaredNorm();
        if (z > RealScalar(0))
            *this /= std::sqrt(z);
    }
    Matrix<RealScalar, Dynamic, Dynamic> cwiseAbs() const
    {
        Matrix<RealScalar, Dynamic, Dynamic> res(rows(), cols());
        for (Index j = 0; j < cols(); j++)
            for (Index i = 0; i < rows(); i++)
                res(i, j) = std::abs(coeff(i, j));
        return res;
    }
    Scalar sum() const
    {
        Scalar acc(0);
        for (Index j = 0; j < cols(); j++)
            for (Index i = 0; i < rows(); i++)
                acc += coeff(i, j);
        return acc;
    }
    Scalar maxCoeff() const
    {
        if (size() == 0)
            throw std::logic_error("Eigen stand-in: maxCoeff() of an empty object");
        Scalar m = coeff(0, 0);
        for (Index j = 0; j < cols(); j++)
            for (Index i = 0; i < rows(); i++)
                if (coeff(i, j) > m)
                    m = coeff(i, j);
        return m;
    }
    // vector form with index (UpperHessenbergEigen.h:370)
    Scalar minCoeff(Index* idx) const
    {
        Scalar m = coeff(0);
        *idx = 0;
        for (Index i = 1; i < size(); i++)
            if (coeff(i) < m)
            {
                m = coeff(i);
                *idx = i;
            }
        return m;
    }
    Scalar value() const { return coeff(0, 0); }

    template <typename T>
    Matrix<T, Dynamic, Dynamic> cast() const
    {
        Matrix<T, Dynamic, Dynamic> res(rows(), cols());
        for (Index j = 0; j < cols(); j++)
            for (Index i = 0; i < rows(); i++)
                res(i, j) = T(coeff(i, j));
        return res;
    }

    // coefficient-wise view: vectors by their stride, contiguous matrices flattened in storage order
    ArrayRef<Scalar> array() const
    {
        if (cols() > 1 && rows() > 1)
        {
            if (!(rs() == 1 && cs() == rows()))
                throw std::logic_error("Eigen stand-in: array() of a non-contiguous matrix block");
            return ArrayRef<Scalar>(ptr_(), size(), 1);
        }
        return ArrayRef<Scalar>(ptr_(), size(), lstride());
    }

    // ---- plane rotations (Eigen 3.4.0 Jacobi.h: apply_rotation_in_the_plane(x, y, j) is
    // x_i <- c x_i + conj(s) y_i ; y_i <- -s x_i + conj(c) y_i; applyOnTheLeft passes rows p, q and j,
    // applyOnTheRight passes columns p, q and j.transpose()) ----
    template <typename T>
    void applyOnTheRight(Index p, Index q, const JacobiRotation<T>& j)
    {
        const JacobiRotation<T> jt = j.transpose();
        const Index r = rows();
        for (Index i = 0; i < r; i++)
        {
            const Scalar xi = coeff(i, p), yi = coeff(i, q);
            coeffRef(i, p) = jt.c() * xi + numext::conj(jt.s()) * yi;
            coeffRef(i, q) = -jt.s() * xi + numext::conj(jt.c()) * yi;
        }
    }
    template <typename T>
    void applyOnTheLeft(Index p, Index q, const JacobiRotation<T>& j)
    {
        const Index c = cols();
        for (Index k = 0; k < c; k++)
        {
            const Scalar xi = coeff(p, k), yi = coeff(q, k);
            coeffRef(p, k) = j.c() * xi + numext::conj(j.s()) * yi;
            coeffRef(q, k) = -j.s() * xi + numext::conj(j.c()) * yi;
        }
    }

    // Eigen 3.4.0 Householder.h, makeHouseholder(essential, tau, beta)
    template <typename Ess>
    void makeHouseholder(Ess& essential, Scalar& tau, RealScalar& beta) const
    {
        using std::sqrt;
        const Index n = size();
        RealScalar tailSqNorm(0);
        for (Index i = 1; i < n; i++)
            tailSqNorm += numext::abs2(coeff(i));
        const Scalar c0 = coeff(0);
        const RealScalar tol = (std::numeric_limits<RealScalar>::min)();
        if (tailSqNorm <= tol && numext::abs2(numext::imag(c0)) <= tol)
        {
            tau = Scalar(0);
            beta = numext::real(c0);
            essential.setZero();
        }
        else
        {
            beta = sqrt(numext::abs2(c0) + tailSqNorm);
            if (numext::real(c0) >= RealScalar(0))
                beta = -beta;
            for (Index i = 1; i < n; i++)
                essential.coeffRef(i - 1) = coeff(i) / (c0 - beta);
            tau = numext::conj((beta - c0) / beta);
        }
    }
};

// ------------------------------------------------------------------------------------------------
// View: (pointer, rows, cols, row stride, column stride).  Map, Ref and every block are views.
// Constness is not tracked (a view of a const object can be written through; the reference does not).
// ------------------------------------------------------------------------------------------------
template <typename S, int R, int C>
class View : public MatrixBase<View<S, R, C>>
{
protected:
    S* m_p;
    Index m_r, m_c, m_rs, m_cs;

public:
    using Base = MatrixBase<View<S, R, C>>;
    using Scalar = S;
    enum { IsRowMajor = 0 };
    View() : m_p(nullptr), m_r(0), m_c(0), m_rs(1), m_cs(0) {}
    View(const S* p, Index r, Index c, Index rs, Index cs) : m_p(const_cast<S*>(p)), m_r(r), m_c(c), m_rs(rs), m_cs(cs) {}
    View(const View&) = default;
    template <int R2, int C2>
    View(const View<S, R2, C2>& o) : m_p(o.ptr_impl()), m_r(o.rows_()), m_c(o.cols_()), m_rs(o.rs_impl()), m_cs(o.cs_impl())
    {}
    template <int R2, int C2, int Opt>
    View(const Matrix<S, R2, C2, Opt>& o) : m_p(o.ptr_impl()), m_r(o.rows_()), m_c(o.cols_()), m_rs(o.rs_impl()), m_cs(o.cs_impl())
    {}

    Index rows_() const { return m_r; }
    Index cols_() const { return m_c; }
    S* ptr_impl() const { return m_p; }
    Index rs_impl() const { return m_rs; }
    Index cs_impl() const { return m_cs; }

    // assignment copies elements (never rebinds)
    View& operator=(const View& o)
    {
        this->zip_(o, [](S& a, const S& b) { a = b; });
        return *this;
    }
    template <typename Other>
    View& operator=(const MatrixBase<Other>& o)
    {
        this->zip_(o, [](S& a, const typename Other::Scalar& b) { a = b; });
        return *this;
    }
};

template <typename S>
class AdjointView
{
public:
    using Scalar = S;
    View<S, Dynamic, Dynamic> m;  // the object whose adjoint this is
    explicit AdjointView(const View<S, Dynamic, Dynamic>& v) : m(v) {}
    Index rows() const { return m.cols(); }
    Index cols() const { return m.rows(); }
    Matrix<S, Dynamic, Dynamic> eval() const;
};

template <typename S, int Uplo>
class DenseSelfAdjointView
{
public:
    View<S, Dynamic, Dynamic> m;
    explicit DenseSelfAdjointView(const View<S, Dynamic, Dynamic>& v) : m(v) {}
    // entry (i, j) of the Hermitian matrix the stored triangle stands for
    S at(Index i, Index j) const
    {
        if (i == j)
            return S(numext::real(m.coeff(i, i)));
        const bool stored = (Uplo == Lower) ? i > j : i < j;
        return stored ? m.coeff(i, j) : numext::conj(m.coeff(j, i));
    }
};

template <typename S>
class DiagonalWrapper
{
public:
    View<S, Dynamic, 1> d;
    explicit DiagonalWrapper(const View<S, Dynamic, 1>& v) : d(v) {}
};

// ------------------------------------------------------------------------------------------------
// Matrix: owning, column-major (the RowMajor option exists only as a compile-time flag for the
// reference's static_asserts; RowMajor dense matrices are not instantiated on this path).
// ------------------------------------------------------------------------------------------------
template <typename S, int R, int C, int Opt>
class Matrix : public MatrixBase<Matrix<S, R, C, Opt>>
{
    std::vector<S> m_a;
    Index m_r, m_c;

    void alloc_(Index r, Index c)
    {
        m_r = r;
        m_c = c;
        m_a.assign(static_cast<size_t>(r * c), S());
    }

public:
    using Base = MatrixBase<Matrix<S, R, C, Opt>>;
    using Scalar = S;
    using PlainObject = Matrix;
    enum { IsRowMajor = (Opt == RowMajor) ? 1 : 0 };

    Matrix() { alloc_(R == Dynamic ? 0 : R, C == Dynamic ? 0 : C); }
    explicit Matrix(Index n)
    {
        static_assert(R == 1 || C == 1 || R == Dynamic || C == Dynamic, "");
        if (C == 1)
            alloc_(n, 1);
        else if (R == 1)
            alloc_(1, n);
        else
            alloc_(n, n);  // not used by the reference
    }
    Matrix(Index r, Index c) { alloc_(r, c); }
    Matrix(const Matrix&) = default;
    Matrix(Matrix&&) = default;
    template <typename Other>
    Matrix(const MatrixBase<Other>& o) : m_r(0), m_c(0)
    {
        assign_(o);
    }

    Index rows_() const { return m_r; }
    Index cols_() const { return m_c; }
    S* ptr_impl() const { return const_cast<S*>(m_a.data()); }
    Index rs_impl() const { return 1; }
    Index cs_impl() const { return m_r; }

    void resize(Index n)
    {
        if (C == 1)
            resize(n, 1);
        else
            resize(1, n);
    }
    void resize(Index r, Index c)
    {
        if (r != m_r || c != m_c)
            alloc_(r, c);
    }

    template <typename Other>
    void assign_(const MatrixBase<Other>& o)
    {
        Index r = o.rows(), c = o.cols();
        // a compile-time vector takes a vector of the other orientation as a sequence
        if (C == 1 && c != 1 && r == 1)
            std::swap(r, c);
        if (R == 1 && r != 1 && c == 1)
            std::swap(r, c);
        if (r != m_r || c != m_c)
        {
            m_r = r;
            m_c = c;
            m_a.resize(static_cast<size_t>(r * c));
        }
        this->zip_(o, [](S& a, const typename Other::Scalar& b) { a = S(b); });
    }
    Matrix& operator=(const Matrix& o)
    {
        m_a = o.m_a;
        m_r = o.m_r;
        m_c = o.m_c;
        return *this;
    }
    Matrix& operator=(Matrix&& o) = default;
    template <typename Other>
    Matrix& operator=(const MatrixBase<Other>& o)
    {
        assign_(o);
        return *this;
    }

    // M << a, b, c, ...;  row by row
    struct CommaInit
    {
        Matrix& m;
        Index k;
        CommaInit& operator,(const S& v)
        {
            m(k / m.cols(), k % m.cols()) = v;
            k++;
            return *this;
        }
    };
    CommaInit operator<<(const S& v)
    {
        (*this)(0, 0) = v;
        return CommaInit{*this, 1};
    }
    // Matrix <-> Matrix swap exchanges storage (Arnoldi.h:337, UpperHessenbergSchur.h swap_T/swap_U)
    void swap(Matrix& o)
    {
        m_a.swap(o.m_a);
        std::swap(m_r, o.m_r);
        std::swap(m_c, o.m_c);
    }
    using Base::swap;

    static Matrix Zero(Index r, Index c)
    {
        Matrix m(r, c);
        return m;
    }
    static Matrix Zero(Index n)
    {
        Matrix m(n);
        return m;
    }
    static Matrix Zero()
    {
        Matrix m;
        m.setZero();
        return m;
    }
    static Matrix Identity(Index r, Index c)
    {
        Matrix m(r, c);
        m.setIdentity();
        return m;
    }
    static Matrix Constant(Index r, Index c, const S& v)
    {
        Matrix m(r, c);
        m.setConstant(v);
        return m;
    }
    static Matrix Constant(Index n, const S& v)
    {
        Matrix m(n);
        m.setConstant(v);
        return m;
    }
    static Matrix Random(Index r, Index c)
    {
        Matrix m(r, c);
        m.setRandom();
        return m;
    }
    static Matrix Random(Index n)
    {
        Matrix m(n);
        m.setRandom();
        return m;
    }
    static Matrix Identity(Index n)
    {
        Matrix m(n, n);
        m.setIdentity();
        return m;
    }
    template <typename SA>
    Matrix(const AdjointView<SA>& a) : m_r(0), m_c(0)
    {
        assign_(a.eval());
    }
    // keeps the leading block
    void conservativeResize(Index r, Index c)
    {
        if (r == m_r && c == m_c)
            return;
        Matrix t(r, c);
        for (Index j = 0; j < (std::min)(c, m_c); j++)
            for (Index i = 0; i < (std::min)(r, m_r); i++)
                t(i, j) = (*this)(i, j);
        swap(t);
    }
    void conservativeResize(Index n)
    {
        if (C == 1)
            conservativeResize(n, 1);
        else
            conservativeResize(1, n);
    }
    // vector of n equally spaced values from low to high
    Matrix& setLinSpaced(Index n, const S& low, const S& high)
    {
        resize(n);
        for (Index i = 0; i < n; i++)
            (*this)[i] = n > 1 ? S(low + (high - low) * i / (n - 1)) : high;
        return *this;
    }
};

// ------------------------------------------------------------------------------------------------
// Map<PlainT>, Ref<PlainT>: views with Eigen's constructor forms
// ------------------------------------------------------------------------------------------------
namespace internal {
template <typename PlainT>
struct plain_info
{
    using P = typename std::remove_const<PlainT>::type;
    using Scalar = typename P::Scalar;
    enum { Rows = P::RowsAtCompileTime, Cols = P::ColsAtCompileTime };
};
}  // namespace internal

template <typename PlainT, typename Enable = void>
class Map : public View<typename internal::plain_info<PlainT>::Scalar, internal::plain_info<PlainT>::Rows, internal::plain_info<PlainT>::Cols>
{
    using Info = internal::plain_info<PlainT>;
    using S = typename Info::Scalar;
    using V = View<S, Info::Rows, Info::Cols>;

public:
    using Scalar = S;
    using PlainObject = typename Info::P;
    Map(const S* p, Index n) : V(p, Info::Rows == 1 ? 1 : n, Info::Rows == 1 ? n : 1, 1, Info::Rows == 1 ? 1 : n) {}
    Map(const S* p, Index r, Index c) : V(p, r, c, 1, r) {}
    Map(const Map&) = default;
    Map& operator=(const Map& o)
    {
        V::operator=(static_cast<const V&>(o));
        return *this;
    }
    template <typename Other>
    Map& operator=(const MatrixBase<Other>& o)
    {
        V::operator=(o);
        return *this;
    }
};

template <typename PlainT, typename Enable = void>
class Ref : public View<typename internal::plain_info<PlainT>::Scalar, internal::plain_info<PlainT>::Rows, internal::plain_info<PlainT>::Cols>
{
    using Info = internal::plain_info<PlainT>;
    using S = typename Info::Scalar;
    using V = View<S, Info::Rows, Info::Cols>;

public:
    using Scalar = S;
    using PlainObject = typename Info::P;
    template <typename Other, typename = typename std::enable_if<std::is_same<typename Other::Scalar, S>::value>::type>
    Ref(const MatrixBase<Other>& o) : V(o.ptr_(), o.rows(), o.cols(), o.rs(), o.cs())
    {}
    Ref(const Ref&) = default;
    Ref& operator=(const Ref& o)
    {
        V::operator=(static_cast<const V&>(o));
        return *this;
    }
    template <typename Other>
    Ref& operator=(const MatrixBase<Other>& o)
    {
        V::operator=(o);
        return *this;
    }
};

namespace internal {
template <typename PlainT, typename E>
struct traits<Map<PlainT, E>> : traits<View<typename plain_info<PlainT>::Scalar, plain_info<PlainT>::Rows, plain_info<PlainT>::Cols>>
{};
template <typename PlainT, typename E>
struct traits<Ref<PlainT, E>> : traits<View<typename plain_info<PlainT>::Scalar, plain_info<PlainT>::Rows, plain_info<PlainT>::Cols>>
{};
}  // namespace internal

// ------------------------------------------------------------------------------------------------
// Eager arithmetic
// ------------------------------------------------------------------------------------------------
template <typename A, typename B>
Matrix<internal::prod_t<typename A::Scalar, typename B::Scalar>, Dynamic, Dynamic> operator+(const MatrixBase<A>& a, const MatrixBase<B>& b)
{
    Matrix<internal::prod_t<typename A::Scalar, typename B::Scalar>, Dynamic, Dynamic> res(a);
    res += b;
    return res;
}
template <typename A, typename B>
Matrix<internal::prod_t<typename A::Scalar, typename B::Scalar>, Dynamic, Dynamic> operator-(const MatrixBase<A>& a, const MatrixBase<B>& b)
{
    Matrix<internal::prod_t<typename A::Scalar, typename B::Scalar>, Dynamic, Dynamic> res(a);
    res -= b;
    return res;
}
template <typename A>
Matrix<typename A::Scalar, Dynamic, Dynamic> operator-(const MatrixBase<A>& a)
{
    Matrix<typename A::Scalar, Dynamic, Dynamic> res(a);
    res *= typename A::Scalar(-1);
    return res;
}
template <typename A, typename T, typename = typename std::enable_if<internal::is_scalar<T>::value>::type>
Matrix<internal::prod_t<typename A::Scalar, T>, Dynamic, Dynamic> operator*(const MatrixBase<A>& a, const T& s)
{
    Matrix<internal::prod_t<typename A::Scalar, T>, Dynamic, Dynamic> res(a);
    res *= s;
    return res;
}
template <typename A, typename T, typename = typename std::enable_if<internal::is_scalar<T>::value>::type>
Matrix<internal::prod_t<typename A::Scalar, T>, Dynamic, Dynamic> operator*(const T& s, const MatrixBase<A>& a)
{
    Matrix<internal::prod_t<typename A::Scalar, T>, Dynamic, Dynamic> res(a);
    res *= s;
    return res;
}
template <typename A, typename T, typename = typename std::enable_if<internal::is_scalar<T>::value>::type>
Matrix<internal::prod_t<typename A::Scalar, T>, Dynamic, Dynamic> operator/(const MatrixBase<A>& a, const T& s)
{
    Matrix<internal::prod_t<typename A::Scalar, T>, Dynamic, Dynamic> res(a);
    res /= s;
    return res;
}

// C = A B, column-oriented (each column of C is a sum of scaled columns of A, ascending k)
template <typename A, typename B>
Matrix<internal::prod_t<typename A::Scalar, typename B::Scalar>, Dynamic, Dynamic> operator*(const MatrixBase<A>& a, const MatrixBase<B>& b)
{
    using RS = internal::prod_t<typename A::Scalar, typename B::Scalar>;
    if (a.cols() != b.rows())
        throw std::logic_error("Eigen stand-in: inner dimensions differ in a matrix product");
    const Index m = a.rows(), kk = a.cols(), n = b.cols();
    Matrix<RS, Dynamic, Dynamic> res(m, n);
    for (Index j = 0; j < n; j++)
    {
        RS* c = res.data() + j * m;
        for (Index k = 0; k < kk; k++)
        {
            const auto bk = b.coeff(k, j);
            const typename A::Scalar* ak = a.ptr_() + k * a.cs();
            const Index s = a.rs();
            if (s == 1)
                for (Index i = 0; i < m; i++)
                    c[i] += ak[i] * bk;
            else
                for (Index i = 0; i < m; i++)
                    c[i] += ak[i * s] * bk;
        }
    }
    return res;
}

// C = A^H B: one conjugated dot product per entry
template <typename SA, typename B>
Matrix<internal::prod_t<SA, typename B::Scalar>, Dynamic, Dynamic> operator*(const AdjointView<SA>& at, const MatrixBase<B>& b)
{
    using RS = internal::prod_t<SA, typename B::Scalar>;
    const auto& a = at.m;
    if (a.rows() != b.rows())
        throw std::logic_error("Eigen stand-in: inner dimensions differ in an adjoint product");
    Matrix<RS, Dynamic, Dynamic> res(a.cols(), b.cols());
    for (Index j = 0; j < b.cols(); j++)
        for (Index i = 0; i < a.cols(); i++)
            res(i, j) = a.col(i).dot(b.col(j));
    return res;
}

template <typename S>
Matrix<S, Dynamic, Dynamic> AdjointView<S>::eval() const
{
    Matrix<S, Dynamic, Dynamic> res(m.cols(), m.rows());
    for (Index j = 0; j < m.rows(); j++)
        for (Index i = 0; i < m.cols(); i++)
            res(i, j) = numext::conj(m.coeff(j, i));
    return res;
}
// an adjoint / transpose anywhere else in an expression is evaluated first (the reference's unit tests: Q.adjoint() * H * Q, mat + mat.transpose(), ...)
template <typename A, typename SB>
auto operator*(const MatrixBase<A>& a, const AdjointView<SB>& b) -> decltype(a * b.eval()) { return a * b.eval(); }
template <typename SA, typename SB>
auto operator*(const AdjointView<SA>& a, const AdjointView<SB>& b) -> decltype(a * b.eval()) { return a * b.eval(); }
template <typename A, typename SB>
auto operator+(const MatrixBase<A>& a, const AdjointView<SB>& b) -> decltype(a + b.eval()) { return a + b.eval(); }
template <typename A, typename SB>
auto operator-(const MatrixBase<A>& a, const AdjointView<SB>& b) -> decltype(a - b.eval()) { return a - b.eval(); }
template <typename SA, typename B>
auto operator+(const AdjointView<SA>& a, const MatrixBase<B>& b) -> decltype(a.eval() + b) { return a.eval() + b; }
template <typename SA, typename B>
auto operator-(const AdjointView<SA>& a, const MatrixBase<B>& b) -> decltype(a.eval() - b) { return a.eval() - b; }

// selfadjointView<Uplo>(A) * B and A * diag(d) for dense objects
template <typename S, int Uplo, typename B>
Matrix<internal::prod_t<S, typename B::Scalar>, Dynamic, Dynamic> operator*(const DenseSelfAdjointView<S, Uplo>& a, const MatrixBase<B>& b)
{
    using RS = internal::prod_t<S, typename B::Scalar>;
    const Index n = a.m.rows();
    Matrix<RS, Dynamic, Dynamic> res(n, b.cols());
    for (Index j = 0; j < b.cols(); j++)
        for (Index k = 0; k < n; k++)
        {
            const auto bk = b.coeff(k, j);
            for (Index i = 0; i < n; i++)
                res(i, j) += a.at(i, k) * bk;
        }
    return res;
}
template <typename A, typename S>
Matrix<internal::prod_t<typename A::Scalar, S>, Dynamic, Dynamic> operator*(const MatrixBase<A>& a, const DiagonalWrapper<S>& d)
{
    Matrix<internal::prod_t<typename A::Scalar, S>, Dynamic, Dynamic> res(a.rows(), a.cols());
    for (Index j = 0; j < a.cols(); j++)
        for (Index i = 0; i < a.rows(); i++)
            res(i, j) = a.coeff(i, j) * d.d.coeff(j);
    return res;
}

template <typename Derived>
std::ostream& operator<<(std::ostream& os, const MatrixBase<Derived>& m)
{
    for (Index i = 0; i < m.rows(); i++)
    {
        for (Index j = 0; j < m.cols(); j++)
            os << (j ? " " : "") << m.coeff(i, j);
        if (i + 1 < m.rows())
            os << "\n";
    }
    return os;
}
template <typename S>
std::ostream& operator<<(std::ostream& os, const AdjointView<S>& a) { return os << a.eval(); }

template <typename Derived>
Matrix<typename MatrixBase<Derived>::Scalar, Dynamic, Dynamic> MatrixBase<Derived>::inverse() const
{
    const Index n = rows();
    Matrix<Scalar, Dynamic, Dynamic> a(derived()), inv = Matrix<Scalar, Dynamic, Dynamic>::Identity(n, n);
    for (Index k = 0; k < n; k++)
    {
        Index p = k;
        for (Index i = k + 1; i < n; i++)
            if (std::abs(a(i, k)) > std::abs(a(p, k)))
                p = i;
        for (Index j = 0; j < n; j++)
        {
            std::swap(a(k, j), a(p, j));
            std::swap(inv(k, j), inv(p, j));
        }
        const Scalar piv = a(k, k);
        for (Index j = 0; j < n; j++)
        {
            a(k, j) /= piv;
            inv(k, j) /= piv;
        }
        for (Index i = 0; i < n; i++)
        {
            if (i == k)
                continue;
            const Scalar f = a(i, k);
            if (f != Scalar(0))
                for (Index j = 0; j < n; j++)
                {
                    a(i, j) -= f * a(k, j);
                    inv(i, j) -= f * inv(k, j);
                }
        }
    }
    return inv;
}

// ------------------------------------------------------------------------------------------------
// Arrays (coefficient-wise world): only what HermEigsBase / GenEigsBase / UpperHessenbergQR touch
// ------------------------------------------------------------------------------------------------
template <typename S, int R = Dynamic, int C = 1>
class Array
{
    std::vector<S> m_a;

public:
    using Scalar = S;
    using RealScalar = typename NumTraits<S>::Real;
    Array() {}
    explicit Array(Index n) : m_a(static_cast<size_t>(n), S()) {}
    Array(const ArrayRef<S>& r);
    Index size() const { return static_cast<Index>(m_a.size()); }
    void resize(Index n) { m_a.assign(static_cast<size_t>(n), S()); }
    void setZero() { std::fill(m_a.begin(), m_a.end(), S()); }
    // std::vector<bool> has no data(); bool arrays are stored as unsigned char underneath
    S* data() { return m_a.data(); }
    const S* data() const { return m_a.data(); }
    S& coeffRef(Index i) { return m_a[static_cast<size_t>(i)]; }
    const S& coeff(Index i) const { return m_a[static_cast<size_t>(i)]; }
    S& operator[](Index i) { return m_a[static_cast<size_t>(i)]; }
    const S& operator[](Index i) const { return m_a[static_cast<size_t>(i)]; }
    S& operator()(Index i) { return m_a[static_cast<size_t>(i)]; }
    const S& operator()(Index i) const { return m_a[static_cast<size_t>(i)]; }
    void swap(Array& o) { m_a.swap(o.m_a); }

    Array<RealScalar, R, C> abs() const
    {
        Array<RealScalar, R, C> res(size());
        for (Index i = 0; i < size(); i++)
            res[i] = std::abs(m_a[i]);
        return res;
    }
    Array max(const S& v) const
    {
        Array res(size());
        for (Index i = 0; i < size(); i++)
            res[i] = (std::max)(m_a[i], v);
        return res;
    }
    Array operator*(const S& v) const
    {
        Array res(size());
        for (Index i = 0; i < size(); i++)
            res[i] = m_a[i] * v;
        return res;
    }
    friend Array operator*(const S& v, const Array& a) { return a * v; }
    Array operator-(const S& v) const
    {
        Array res(size());
        for (Index i = 0; i < size(); i++)
            res[i] = m_a[i] - v;
        return res;
    }
    Array operator+(const S& v) const
    {
        Array res(size());
        for (Index i = 0; i < size(); i++)
            res[i] = m_a[i] + v;
        return res;
    }
    Array operator/(const S& v) const
    {
        Array res(size());
        for (Index i = 0; i < size(); i++)
            res[i] = m_a[i] / v;
        return res;
    }
    S maxCoeff() const
    {
        S m = m_a.at(0);
        for (Index i = 1; i < size(); i++)
            if (m_a[i] > m)
                m = m_a[i];
        return m;
    }
    Array log10() const
    {
        Array res(size());
        for (Index i = 0; i < size(); i++)
            res[i] = std::log10(m_a[i]);
        return res;
    }
    S mean() const { return sum() / S(size()); }
    Index count() const
    {
        Index n = 0;
        for (Index i = 0; i < size(); i++)
            n += m_a[i] ? 1 : 0;
        return n;
    }
    template <typename T>
    Array<T, R, C> cast() const
    {
        Array<T, R, C> res(size());
        for (Index i = 0; i < size(); i++)
            res[i] = T(m_a[i]);
        return res;
    }
    S sum() const
    {
        S acc(0);
        for (Index i = 0; i < size(); i++)
            acc += m_a[i];
        return acc;
    }
};

// bool arrays: element type unsigned char underneath so that operator[] returns a real reference
template <int R, int C>
class Array<bool, R, C>
{
    std::vector<unsigned char> m_a;

public:
    using Scalar = bool;
    Array() {}
    explicit Array(Index n) : m_a(static_cast<size_t>(n), 0) {}
    Index size() const { return static_cast<Index>(m_a.size()); }
    void resize(Index n) { m_a.assign(static_cast<size_t>(n), 0); }
    void setZero() { std::fill(m_a.begin(), m_a.end(), 0); }
    struct BoolRef
    {
        unsigned char& b;
        BoolRef& operator=(bool v)
        {
            b = v ? 1 : 0;
            return *this;
        }
        BoolRef& operator=(const BoolRef& o)
        {
            b = o.b;
            return *this;
        }
        operator bool() const { return b != 0; }
    };
    BoolRef operator[](Index i) { return BoolRef{m_a[static_cast<size_t>(i)]}; }
    bool operator[](Index i) const { return m_a[static_cast<size_t>(i)] != 0; }
    void swap(Array& o) { m_a.swap(o.m_a); }
    Index count() const
    {
        Index n = 0;
        for (unsigned char b : m_a)
            n += b ? 1 : 0;
        return n;
    }
    template <typename T>
    Array<T, R, C> cast() const
    {
        Array<T, R, C> res(size());
        for (Index i = 0; i < size(); i++)
            res[i] = T(m_a[static_cast<size_t>(i)] ? 1 : 0);
        return res;
    }
};

template <typename S, int R, int C>
Array<bool, R, C> operator<(const Array<S, R, C>& a, const Array<S, R, C>& b)
{
    if (a.size() != b.size())
        throw std::logic_error("Eigen stand-in: array size mismatch");
    Array<bool, R, C> res(a.size());
    for (Index i = 0; i < a.size(); i++)
        res[i] = a[i] < b[i];
    return res;
}

// a strided window onto a matrix / vector, seen coefficient-wise
template <typename S>
class ArrayRef
{
    S* m_p;
    Index m_n, m_s;

public:
    using RealScalar = typename NumTraits<S>::Real;
    ArrayRef(S* p, Index n, Index s) : m_p(p), m_n(n), m_s(s) {}
    Index size() const { return m_n; }
    const S& operator[](Index i) const { return m_p[i * m_s]; }
    ArrayRef& operator-=(const S& v)
    {
        for (Index i = 0; i < m_n; i++)
            m_p[i * m_s] -= v;
        return *this;
    }
    ArrayRef& operator+=(const S& v)
    {
        for (Index i = 0; i < m_n; i++)
            m_p[i * m_s] += v;
        return *this;
    }
    ArrayRef& operator=(const Array<S, Dynamic, 1>& a)
    {
        if (a.size() != m_n)
            throw std::logic_error("Eigen stand-in: array size mismatch");
        for (Index i = 0; i < m_n; i++)
            m_p[i * m_s] = a[i];
        return *this;
    }
    Array<S, Dynamic, 1> operator-(const S& v) const
    {
        Array<S, Dynamic, 1> res(m_n);
        for (Index i = 0; i < m_n; i++)
            res[i] = m_p[i * m_s] - v;
        return res;
    }
    Array<RealScalar, Dynamic, 1> abs() const
    {
        Array<RealScalar, Dynamic, 1> res(m_n);
        for (Index i = 0; i < m_n; i++)
            res[i] = std::abs(m_p[i * m_s]);
        return res;
    }
    Array<S, Dynamic, 1> operator/(const S& v) const
    {
        Array<S, Dynamic, 1> res(m_n);
        for (Index i = 0; i < m_n; i++)
            res[i] = m_p[i * m_s] / v;
        return res;
    }
};
// scalar / array, coefficient-wise (SymEigsShiftSolver.h:167: lambda = 1 / nu + sigma)
template <typename S>
Array<S, Dynamic, 1> operator/(const S& v, const ArrayRef<S>& a)
{
    Array<S, Dynamic, 1> res(a.size());
    for (Index i = 0; i < a.size(); i++)
        res[i] = v / a[i];
    return res;
}

template <typename S, int R, int C>
Array<S, R, C>::Array(const ArrayRef<S>& r) : m_a(static_cast<size_t>(r.size()))
{
    for (Index i = 0; i < r.size(); i++)
        m_a[static_cast<size_t>(i)] = r[i];
}

// Eigen::ComplexSchur, the part UpperHessenbergEigen<std::complex<T>> uses (UpperHessenbergEigen.h:328-454):
// computeFromHessenberg -> reduceToTriangularForm.  Restated from Eigen 3.4.0's published ComplexSchur.h:
// deflation test |T(i+1,i)|_1 <= eps (|T(i,i)|_1 + |T(i+1,i+1)|_1), Wilkinson-type shift from the trailing 2x2
// block, EISPACK comqr exceptional shifts at iterations 10 and 20, at most 30 iterations per row.
template <typename MatrixType>
class ComplexSchur
{
public:
    using ComplexScalar = typename MatrixType::Scalar;
    using RealScalar = typename NumTraits<ComplexScalar>::Real;
    using ComplexMatrixType = Matrix<ComplexScalar, Dynamic, Dynamic>;

    ComplexSchur() : m_info(Success) {}

    template <typename HessMatrixType, typename OrthMatrixType>
    ComplexSchur& computeFromHessenberg(const HessMatrixType& matrixH, const OrthMatrixType& matrixQ, bool computeU = true)
    {
        m_matT = matrixH;
        if (computeU)
            m_matU = matrixQ;
        reduceToTriangularForm(computeU);
        return *this;
    }
    const ComplexMatrixType& matrixT() const { return m_matT; }
    const ComplexMatrixType& matrixU() const { return m_matU; }
    ComputationInfo info() const { return m_info; }

private:
    ComplexMatrixType m_matT, m_matU;
    ComputationInfo m_info;

    bool subdiagonalEntryIsNeglegible(Index i)
    {
        RealScalar d = numext::norm1(m_matT.coeff(i, i)) + numext::norm1(m_matT.coeff(i + 1, i + 1));
        RealScalar sd = numext::norm1(m_matT.coeff(i + 1, i));
        if (std::abs(sd) <= std::abs(d) * NumTraits<RealScalar>::epsilon())  // internal::isMuchSmallerThan
        {
            m_matT.coeffRef(i + 1, i) = ComplexScalar(0);
            return true;
        }
        return false;
    }

    ComplexScalar computeShift(Index iu, Index iter)
    {
        using std::abs;
        if (iter == 10 || iter == 20)
            return abs(numext::real(m_matT.coeff(iu, iu - 1))) + abs(numext::real(m_matT.coeff(iu - 1, iu - 2)));
        Matrix<ComplexScalar, 2, 2> t = m_matT.template block<2, 2>(iu - 1, iu - 1);
        RealScalar normt = t.cwiseAbs().sum();
        t /= normt;
        ComplexScalar b = t.coeff(0, 1) * t.coeff(1, 0);
        ComplexScalar c = t.coeff(0, 0) - t.coeff(1, 1);
        ComplexScalar disc = std::sqrt(c * c + RealScalar(4) * b);
        ComplexScalar det = t.coeff(0, 0) * t.coeff(1, 1) - b;
        ComplexScalar trace = t.coeff(0, 0) + t.coeff(1, 1);
        ComplexScalar eival1 = (trace + disc) / RealScalar(2);
        ComplexScalar eival2 = (trace - disc) / RealScalar(2);
        RealScalar eival1_norm = numext::norm1(eival1);
        RealScalar eival2_norm = numext::norm1(eival2);
        if (eival1_norm > eival2_norm)
            eival2 = det / eival1;
        else if (eival2_norm != RealScalar(0))
            eival1 = det / eival2;
        if (numext::norm1(eival1 - t.coeff(1, 1)) < numext::norm1(eival2 - t.coeff(1, 1)))
            return normt * eival1;
        return normt * eival2;
    }

    void reduceToTriangularForm(bool computeU)
    {
        const Index maxIterations = 30 * m_matT.rows();
        Index iu = m_matT.cols() - 1;
        Index il;
        Index iter = 0;
        Index totalIter = 0;
        while (true)
        {
            while (iu > 0)
            {
                if (!subdiagonalEntryIsNeglegible(iu - 1))
                    break;
                iter = 0;
                --iu;
            }
            if (iu == 0)
                break;
            iter++;
            totalIter++;
            if (totalIter > maxIterations)
                break;
            il = iu - 1;
            while (il > 0 && !subdiagonalEntryIsNeglegible(il - 1))
                --il;
            ComplexScalar shift = computeShift(iu, iter);
            JacobiRotation<ComplexScalar> rot;
            rot.makeGivens(m_matT.coeff(il, il) - shift, m_matT.coeff(il + 1, il));
            m_matT.rightCols(m_matT.cols() - il).applyOnTheLeft(il, il + 1, rot.adjoint());
            m_matT.topRows((std::min)(il + 2, iu) + 1).applyOnTheRight(il, il + 1, rot);
            if (computeU)
                m_matU.applyOnTheRight(il, il + 1, rot);
            for (Index i = il + 1; i < iu; i++)
            {
                rot.makeGivens(m_matT.coeffRef(i, i - 1), m_matT.coeffRef(i + 1, i - 1), &m_matT.coeffRef(i, i - 1));
                m_matT.coeffRef(i + 1, i - 1) = ComplexScalar(0);
                m_matT.rightCols(m_matT.cols() - i).applyOnTheLeft(i, i + 1, rot.adjoint());  // column i - 1 already holds (r, 0)
                m_matT.topRows((std::min)(i + 2, iu) + 1).applyOnTheRight(i, i + 1, rot);
                if (computeU)
                    m_matU.applyOnTheRight(i, i + 1, rot);
            }
        }
        m_info = (totalIter <= maxIterations) ? Success : NoConvergence;
    }
};

// ------------------------------------------------------------------------------------------------
// SparseCore subset: compressed storage + the two products of MatOp/Sparse{Sym,Gen}MatProd.h
// ------------------------------------------------------------------------------------------------
template <typename Derived>
class SparseMatrixBase
{
public:
    const Derived& derived() const { return *static_cast<const Derived*>(this); }
    Index rows() const { return derived().sm_rows(); }
    Index cols() const { return derived().sm_cols(); }
};

template <typename S, int Flags = ColMajor, typename StorageIndex = int>
class SparseMatrix;

// the compressed arrays of a matrix somebody else owns
template <typename S, int Flags, typename StorageIndex>
struct SparseData
{
    Index rows = 0, cols = 0, nnz = 0;
    const StorageIndex* outer = nullptr;
    const StorageIndex* inner = nullptr;
    const S* values = nullptr;
};

template <typename S, int Flags, typename StorageIndex, int Uplo>
class SparseSelfAdjointView
{
public:
    SparseData<S, Flags, StorageIndex> d;
};

template <typename S, int Flags, typename StorageIndex>
class SparseCompressedBase
{
protected:
    SparseData<S, Flags, StorageIndex> m_d;

public:
    using Scalar = S;
    enum { IsRowMajor = (Flags & RowMajor) ? 1 : 0 };
    Index rows() const { return m_d.rows; }
    Index cols() const { return m_d.cols; }
    Index sm_rows() const { return m_d.rows; }
    Index sm_cols() const { return m_d.cols; }
    Index nonZeros() const { return m_d.nnz; }
    Index outerSize() const { return IsRowMajor ? m_d.rows : m_d.cols; }
    const SparseData<S, Flags, StorageIndex>& raw() const { return m_d; }
    // SparseMatrix::coeff: search inside the outer vector (inner indices ascending or not)
    S coeff(Index i, Index j) const
    {
        const Index o = IsRowMajor ? i : j, in = IsRowMajor ? j : i;
        S acc(0);
        for (Index k = m_d.outer[o]; k < m_d.outer[o + 1]; k++)
            if (m_d.inner[k] == in)
                acc += m_d.values[k];
        return acc;
    }
    template <int Uplo>
    SparseSelfAdjointView<S, Flags, StorageIndex, Uplo> selfadjointView() const
    {
        SparseSelfAdjointView<S, Flags, StorageIndex, Uplo> v;
        v.d = m_d;
        return v;
    }
};

template <typename S, typename StorageIndex = int>
class Triplet
{
    StorageIndex m_r, m_c;
    S m_v;

public:
    Triplet() : m_r(0), m_c(0), m_v(0) {}
    Triplet(const StorageIndex& i, const StorageIndex& j, const S& v = S(0)) : m_r(i), m_c(j), m_v(v) {}
    const StorageIndex& row() const { return m_r; }
    const StorageIndex& col() const { return m_c; }
    const S& value() const { return m_v; }
};

// Owning sparse matrix in Eigen's storage scheme: compressed (outer[k] .. outer[k + 1]) or, after reserve() / insert(), uncompressed
// (outer[k] .. outer[k] + innerNonZeros[k], free room behind each inner vector).  Inner vectors are kept sorted, as Eigen's insert() does.
template <typename S, int Flags, typename StorageIndex>
class SparseMatrix : public SparseMatrixBase<SparseMatrix<S, Flags, StorageIndex>>
{
public:
    using PlainObject = SparseMatrix;
    using Scalar = S;
    enum { IsRowMajor = (Flags & RowMajor) ? 1 : 0 };

private:
    Index m_rows = 0, m_cols = 0;
    std::vector<StorageIndex> m_outer;     // outerSize + 1
    std::vector<StorageIndex> m_innernnz;  // outerSize; empty <=> compressed
    std::vector<StorageIndex> m_inner;
    std::vector<S> m_values;
    mutable SparseData<S, Flags, StorageIndex> m_view;
    mutable std::vector<StorageIndex> m_couter, m_cinner;  // compressed copy handed to Ref<> while *this is uncompressed
    mutable std::vector<S> m_cvalues;

    Index outer_of(Index i, Index j) const { return IsRowMajor ? i : j; }
    Index inner_of(Index i, Index j) const { return IsRowMajor ? j : i; }
    Index count(Index o) const { return m_innernnz.empty() ? Index(m_outer[o + 1] - m_outer[o]) : Index(m_innernnz[o]); }

    void uncompress(Index extra_per_outer)
    {
        const Index no = outerSize();
        std::vector<StorageIndex> outer(no + 1), nnz(no);
        Index pos = 0;
        for (Index o = 0; o < no; o++)
        {
            outer[o] = StorageIndex(pos);
            nnz[o] = StorageIndex(count(o));
            pos += count(o) + extra_per_outer;
        }
        outer[no] = StorageIndex(pos);
        std::vector<StorageIndex> inner(pos);
        std::vector<S> values(pos);
        for (Index o = 0; o < no; o++)
            for (Index k = 0; k < count(o); k++)
            {
                inner[outer[o] + k] = m_inner[m_outer[o] + k];
                values[outer[o] + k] = m_values[m_outer[o] + k];
            }
        m_outer.swap(outer);
        m_innernnz.swap(nnz);
        m_inner.swap(inner);
        m_values.swap(values);
    }

public:
    SparseMatrix() { m_outer.assign(1, 0); }
    SparseMatrix(Index rows, Index cols) { resize(rows, cols); }
    // the full (Hermitian) matrix a selfadjointView<Uplo> stands for: the stored triangle plus its conjugated mirror, real diagonal
    template <int Uplo>
    SparseMatrix(const SparseSelfAdjointView<S, Flags, StorageIndex, Uplo>& v)
    {
        const auto& d = v.d;
        resize(d.rows, d.cols);
        std::vector<Triplet<S, StorageIndex>> t;
        const Index no = IsRowMajor ? d.rows : d.cols;
        for (Index o = 0; o < no; o++)
            for (Index k = d.outer[o]; k < d.outer[o + 1]; k++)
            {
                const Index in = d.inner[k];
                const Index r = IsRowMajor ? o : in, c = IsRowMajor ? in : o;
                if ((Uplo == Lower) ? r < c : r > c)
                    continue;
                if (r == c)
                    t.emplace_back(StorageIndex(r), StorageIndex(c), S(numext::real(d.values[k])));
                else
                {
                    t.emplace_back(StorageIndex(r), StorageIndex(c), d.values[k]);
                    t.emplace_back(StorageIndex(c), StorageIndex(r), numext::conj(d.values[k]));
                }
            }
        setFromTriplets(t.begin(), t.end());
    }
    SparseMatrix transpose() const
    {
        SparseMatrix R(m_cols, m_rows);
        std::vector<Triplet<S, StorageIndex>> t;
        const auto& d = raw();
        for (Index o = 0; o < outerSize(); o++)
            for (Index k = d.outer[o]; k < d.outer[o + 1]; k++)
                t.emplace_back(StorageIndex(IsRowMajor ? d.inner[k] : o), StorageIndex(IsRowMajor ? o : d.inner[k]), d.values[k]);
        R.setFromTriplets(t.begin(), t.end());
        return R;
    }
    void setIdentity()
    {
        std::vector<Triplet<S, StorageIndex>> t;
        for (Index i = 0; i < (std::min)(m_rows, m_cols); i++)
            t.emplace_back(StorageIndex(i), StorageIndex(i), S(1));
        setFromTriplets(t.begin(), t.end());
    }
    // eager sparse arithmetic: alpha * A + beta * B as a new compressed matrix (entries that cancel stay stored, as in Eigen)
    static SparseMatrix combine(const S& alpha, const SparseMatrix& A, const S& beta, const SparseMatrix& Bm)
    {
        SparseMatrix R(A.rows(), A.cols());
        std::vector<Triplet<S, StorageIndex>> t;
        auto push = [&](const SparseMatrix& X, const S& f) {
            const auto& d = X.raw();
            for (Index o = 0; o < X.outerSize(); o++)
                for (Index k = d.outer[o]; k < d.outer[o + 1]; k++)
                    t.emplace_back(StorageIndex(IsRowMajor ? o : d.inner[k]), StorageIndex(IsRowMajor ? d.inner[k] : o), f * d.values[k]);
        };
        push(A, alpha);
        push(Bm, beta);
        R.setFromTriplets(t.begin(), t.end());
        return R;
    }
    friend SparseMatrix operator*(const S& f, const SparseMatrix& A) { return combine(f, A, S(0), SparseMatrix(A.rows(), A.cols())); }
    friend SparseMatrix operator-(const SparseMatrix& A, const SparseMatrix& Bm) { return combine(S(1), A, S(-1), Bm); }
    friend SparseMatrix operator+(const SparseMatrix& A, const SparseMatrix& Bm) { return combine(S(1), A, S(1), Bm); }

    void resize(Index rows, Index cols)
    {
        m_rows = rows;
        m_cols = cols;
        m_outer.assign(size_t(outerSize() + 1), 0);
        m_innernnz.clear();
        m_inner.clear();
        m_values.clear();
    }
    Index rows() const { return m_rows; }
    Index cols() const { return m_cols; }
    Index sm_rows() const { return m_rows; }
    Index sm_cols() const { return m_cols; }
    Index outerSize() const { return IsRowMajor ? m_rows : m_cols; }
    Index innerSize() const { return IsRowMajor ? m_cols : m_rows; }
    bool isCompressed() const { return m_innernnz.empty(); }
    Index nonZeros() const
    {
        Index n = 0;
        for (Index o = 0; o < outerSize(); o++)
            n += count(o);
        return n;
    }
    const StorageIndex* outerIndexPtr() const { return m_outer.data(); }
    const StorageIndex* innerIndexPtr() const { return m_inner.data(); }
    const StorageIndex* innerNonZeroPtr() const { return m_innernnz.empty() ? nullptr : m_innernnz.data(); }
    const S* valuePtr() const { return m_values.data(); }

    // reserve(sizes): room for sizes[k] more entries in inner vector k; switches to uncompressed mode
    template <typename SizesType>
    void reserve(const SizesType& sizes)
    {
        const Index no = outerSize();
        std::vector<StorageIndex> outer(no + 1), nnz(no);
        Index pos = 0;
        for (Index o = 0; o < no; o++)
        {
            outer[o] = StorageIndex(pos);
            nnz[o] = StorageIndex(count(o));
            pos += count(o) + Index(sizes[o]);
        }
        outer[no] = StorageIndex(pos);
        std::vector<StorageIndex> inner(pos);
        std::vector<S> values(pos);
        for (Index o = 0; o < no; o++)
            for (Index k = 0; k < count(o); k++)
            {
                inner[outer[o] + k] = m_inner[m_outer[o] + k];
                values[outer[o] + k] = m_values[m_outer[o] + k];
            }
        m_outer.swap(outer);
        m_innernnz.swap(nnz);
        m_inner.swap(inner);
        m_values.swap(values);
    }
    void reserve(Index) {}  // reserve(nnz): a capacity hint in Eigen; nothing to do here

    // insert(i, j): the entry must not exist yet; returns a reference to its (zero) value
    S& insert(Index i, Index j)
    {
        const Index o = outer_of(i, j), in = inner_of(i, j);
        if (isCompressed())
            uncompress(2);
        if (m_outer[o] + m_innernnz[o] >= m_outer[o + 1])
        {
            // inner vector full: give every inner vector more room
            std::vector<Index> more(size_t(outerSize()), Index(2));
            more[size_t(o)] = (std::max)(Index(2), Index(m_innernnz[o]));
            reserve(more);
        }
        Index k = m_outer[o] + m_innernnz[o];
        while (k > m_outer[o] && m_inner[k - 1] > in)
        {
            m_inner[k] = m_inner[k - 1];
            m_values[k] = m_values[k - 1];
            k--;
        }
        if (k > m_outer[o] && m_inner[k - 1] == in)
            throw std::logic_error("Eigen stand-in: insert() of an existing entry");
        m_inner[k] = StorageIndex(in);
        m_values[k] = S(0);
        m_innernnz[o]++;
        return m_values[k];
    }
    S& coeffRef(Index i, Index j)
    {
        const Index o = outer_of(i, j), in = inner_of(i, j);
        for (Index k = m_outer[o]; k < m_outer[o] + count(o); k++)
            if (m_inner[k] == in)
                return m_values[k];
        return insert(i, j);
    }
    S coeff(Index i, Index j) const
    {
        const Index o = outer_of(i, j), in = inner_of(i, j);
        for (Index k = m_outer[o]; k < m_outer[o] + count(o); k++)
            if (m_inner[k] == in)
                return m_values[k];
        return S(0);
    }

    void makeCompressed()
    {
        if (isCompressed())
            return;
        const Index no = outerSize();
        Index pos = 0;
        for (Index o = 0; o < no; o++)
        {
            const Index start = m_outer[o], cnt = m_innernnz[o];
            for (Index k = 0; k < cnt; k++)
            {
                m_inner[pos + k] = m_inner[start + k];
                m_values[pos + k] = m_values[start + k];
            }
            m_outer[o] = StorageIndex(pos);
            pos += cnt;
        }
        m_outer[no] = StorageIndex(pos);
        m_inner.resize(pos);
        m_values.resize(pos);
        m_innernnz.clear();
    }

    // setFromTriplets(begin, end): duplicates are summed, result compressed
    template <typename It>
    void setFromTriplets(It begin, It end)
    {
        const Index no = outerSize();
        std::vector<std::vector<std::pair<StorageIndex, S>>> cols(no);
        for (It it = begin; it != end; ++it)
            cols[size_t(outer_of(it->row(), it->col()))].emplace_back(StorageIndex(inner_of(it->row(), it->col())), it->value());
        m_outer.assign(size_t(no + 1), 0);
        m_inner.clear();
        m_values.clear();
        m_innernnz.clear();
        for (Index o = 0; o < no; o++)
        {
            auto& c = cols[size_t(o)];
            std::stable_sort(c.begin(), c.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
            for (size_t k = 0; k < c.size(); k++)
            {
                if (k > 0 && c[k].first == c[k - 1].first)
                    m_values.back() += c[k].second;
                else
                {
                    m_inner.push_back(c[k].first);
                    m_values.push_back(c[k].second);
                }
            }
            m_outer[size_t(o + 1)] = StorageIndex(m_inner.size());
        }
    }

    // what Ref<const SparseMatrix> binds to: the arrays themselves when compressed, a packed copy otherwise
    const SparseData<S, Flags, StorageIndex>& raw() const
    {
        m_view.rows = m_rows;
        m_view.cols = m_cols;
        if (isCompressed())
        {
            m_view.nnz = Index(m_inner.size());
            m_view.outer = m_outer.data();
            m_view.inner = m_inner.data();
            m_view.values = m_values.data();
            return m_view;
        }
        const Index no = outerSize();
        m_couter.assign(size_t(no + 1), 0);
        m_cinner.clear();
        m_cvalues.clear();
        for (Index o = 0; o < no; o++)
        {
            for (Index k = 0; k < m_innernnz[o]; k++)
            {
                m_cinner.push_back(m_inner[m_outer[o] + k]);
                m_cvalues.push_back(m_values[m_outer[o] + k]);
            }
            m_couter[size_t(o + 1)] = StorageIndex(m_cinner.size());
        }
        m_view.nnz = Index(m_cinner.size());
        m_view.outer = m_couter.data();
        m_view.inner = m_cinner.data();
        m_view.values = m_cvalues.data();
        return m_view;
    }
    template <int Uplo>
    SparseSelfAdjointView<S, Flags, StorageIndex, Uplo> selfadjointView() const
    {
        SparseSelfAdjointView<S, Flags, StorageIndex, Uplo> v;
        v.d = raw();
        return v;
    }
};

template <typename S, int Flags, typename StorageIndex, typename B>
Matrix<internal::prod_t<S, typename B::Scalar>, Dynamic, Dynamic> operator*(const SparseMatrix<S, Flags, StorageIndex>& A, const MatrixBase<B>& x)
{
    Ref<const SparseMatrix<S, Flags, StorageIndex>> r(A);
    return static_cast<const SparseCompressedBase<S, Flags, StorageIndex>&>(r) * x;
}

// Map<const SparseMatrix>(rows, cols, nnz, outerIndexPtr, innerIndexPtr, valuePtr): Eigen's constructor
template <typename S, int Flags, typename StorageIndex>
class Map<const SparseMatrix<S, Flags, StorageIndex>, void>
    : public SparseCompressedBase<S, Flags, StorageIndex>, public SparseMatrixBase<Map<const SparseMatrix<S, Flags, StorageIndex>, void>>
{
public:
    using PlainObject = SparseMatrix<S, Flags, StorageIndex>;
    using Scalar = S;
    using SparseCompressedBase<S, Flags, StorageIndex>::rows;
    using SparseCompressedBase<S, Flags, StorageIndex>::cols;
    Map(Index rows, Index cols, Index nnz, const StorageIndex* outer, const StorageIndex* inner, const S* values)
    {
        this->m_d.rows = rows;
        this->m_d.cols = cols;
        this->m_d.nnz = nnz;
        this->m_d.outer = outer;
        this->m_d.inner = inner;
        this->m_d.values = values;
    }
};

template <typename S, int Flags, typename StorageIndex>
class Ref<const SparseMatrix<S, Flags, StorageIndex>, void> : public SparseCompressedBase<S, Flags, StorageIndex>
{
public:
    using Scalar = S;
    template <typename Derived>
    Ref(const SparseMatrixBase<Derived>& m)
    {
        this->m_d = m.derived().raw();
    }
};

// y = A x, A general compressed (SparseGenMatProd.h:86): row-major = one dot product per row,
// column-major = scaled columns accumulated in column order (Eigen's sparse_time_dense_product)
template <typename S, int Flags, typename StorageIndex, typename B>
Matrix<internal::prod_t<S, typename B::Scalar>, Dynamic, Dynamic> operator*(const SparseCompressedBase<S, Flags, StorageIndex>& A, const MatrixBase<B>& x)
{
    using RS = internal::prod_t<S, typename B::Scalar>;
    const auto& d = A.raw();
    Matrix<RS, Dynamic, Dynamic> y(d.rows, x.cols());
    for (Index c = 0; c < x.cols(); c++)
    {
        if (Flags & RowMajor)
        {
            for (Index i = 0; i < d.rows; i++)
            {
                RS acc(0);
                for (Index k = d.outer[i]; k < d.outer[i + 1]; k++)
                    acc += d.values[k] * x.coeff(d.inner[k], c);
                y(i, c) = acc;
            }
        }
        else
        {
            for (Index j = 0; j < d.cols; j++)
            {
                const auto xj = x.coeff(j, c);
                for (Index k = d.outer[j]; k < d.outer[j + 1]; k++)
                    y(d.inner[k], c) += d.values[k] * xj;
            }
        }
    }
    return y;
}

// y = selfadjointView<Uplo>(A) x (SparseSymMatProd.h:87): only the Uplo triangle of A is read; each
// stored off-diagonal entry contributes to two rows (Eigen's sparse_selfadjoint_time_dense_product)
template <typename S, int Flags, typename StorageIndex, int Uplo, typename B>
Matrix<internal::prod_t<S, typename B::Scalar>, Dynamic, Dynamic> operator*(const SparseSelfAdjointView<S, Flags, StorageIndex, Uplo>& A, const MatrixBase<B>& x)
{
    using RS = internal::prod_t<S, typename B::Scalar>;
    const auto& d = A.d;
    const bool row_major = (Flags & RowMajor) != 0;
    const Index nouter = row_major ? d.rows : d.cols;
    Matrix<RS, Dynamic, Dynamic> y(d.rows, x.cols());
    for (Index c = 0; c < x.cols(); c++)
    {
        for (Index o = 0; o < nouter; o++)
        {
            const auto xo = x.coeff(o, c);
            RS acc(0);
            for (Index k = d.outer[o]; k < d.outer[o + 1]; k++)
            {
                const Index in = d.inner[k];
                // (row, col) of the stored entry
                const Index r = row_major ? o : in, cc = row_major ? in : o;
                const bool in_triangle = (Uplo == Lower) ? (r >= cc) : (r <= cc);
                if (!in_triangle)
                    continue;
                const S v = d.values[k];
                if (in == o)
                    y(o, c) += numext::real(v) * xo;
                else
                {
                    // entry A(r, cc) = v and its mirror A(cc, r) = conj(v); seen from outer index o:
                    // the entry's own row/column gets v * x(other), the mirrored one conj(v) * x(o)
                    if (row_major)
                    {
                        acc += v * x.coeff(in, c);                // row o, column in
                        y(in, c) += numext::conj(v) * xo;        // row in, column o
                    }
                    else
                    {
                        y(in, c) += v * xo;                       // row in, column o
                        acc += numext::conj(v) * x.coeff(in, c);  // row o, column in
                    }
                }
            }
            y(o, c) += acc;
        }
    }
    return y;
}

// Eigen::SparseLU as MatOp/SparseSymShiftSolve.h uses it (:51, :91-94, :108): compute() / info() / solve().  Stand-in: LU with partial
// pivoting on LAPACK band storage (the DGBTF2 / DGBTRS algorithm), with the band limits read off the pattern -- a direct P A = L U solve
// like SuperLU's; the factors differ from Eigen's supernodal ones, the solve result does not (to rounding).  A zero pivot gives
// info() == NumericalIssue, which the reference turns into "factorization failed with the given shift".
template <typename MatrixType>
class SparseLU
{
    using S = typename MatrixType::Scalar;
    Index m_n = 0, m_kl = 0, m_ku = 0, m_ld = 0;
    std::vector<S> m_ab;
    std::vector<Index> m_ipiv;
    ComputationInfo m_info = InvalidInput;
    S& ab(Index i, Index j) { return m_ab[size_t((m_kl + m_ku + i - j) + j * m_ld)]; }  // A(i, j), max(0, j - ku - kl) <= i <= min(n - 1, j + kl)
    const S& ab(Index i, Index j) const { return m_ab[size_t((m_kl + m_ku + i - j) + j * m_ld)]; }

public:
    SparseLU() {}
    void isSymmetric(bool) {}
    ComputationInfo info() const { return m_info; }
    void compute(const MatrixType& A)
    {
        const auto& d = A.raw();
        const bool row_major = MatrixType::IsRowMajor != 0;
        m_n = A.rows();
        m_kl = m_ku = 0;
        for (Index o = 0; o < A.outerSize(); o++)
            for (Index k = d.outer[o]; k < d.outer[o + 1]; k++)
            {
                const Index r = row_major ? o : d.inner[k], c = row_major ? d.inner[k] : o;
                m_kl = (std::max)(m_kl, r - c);
                m_ku = (std::max)(m_ku, c - r);
            }
        m_ld = 2 * m_kl + m_ku + 1;
        m_ab.assign(size_t(m_ld * m_n), S(0));
        for (Index o = 0; o < A.outerSize(); o++)
            for (Index k = d.outer[o]; k < d.outer[o + 1]; k++)
            {
                const Index r = row_major ? o : d.inner[k], c = row_major ? d.inner[k] : o;
                ab(r, c) += d.values[k];
            }
        m_ipiv.assign(size_t(m_n), 0);
        m_info = Success;
        for (Index j = 0; j < m_n; j++)
        {
            const Index last = (std::min)(m_n - 1, j + m_kl);
            Index p = j;
            for (Index i = j + 1; i <= last; i++)
                if (std::abs(ab(i, j)) > std::abs(ab(p, j)))
                    p = i;
            m_ipiv[size_t(j)] = p;
            if (ab(p, j) == S(0))
            {
                m_info = NumericalIssue;
                return;
            }
            const Index cmax = (std::min)(m_n - 1, j + m_ku + m_kl);
            if (p != j)
                for (Index c = j; c <= cmax; c++)
                    std::swap(ab(j, c), ab(p, c));
            const S inv = S(1) / ab(j, j);
            for (Index i = j + 1; i <= last; i++)
            {
                const S l = ab(i, j) * inv;
                ab(i, j) = l;
                if (l != S(0))
                    for (Index c = j + 1; c <= cmax; c++)
                        ab(i, c) -= l * ab(j, c);
            }
        }
    }
    template <typename Rhs>
    Matrix<S, Dynamic, Dynamic> solve(const MatrixBase<Rhs>& b) const
    {
        if (m_info != Success)
            throw std::logic_error("Eigen stand-in: SparseLU::solve() without a successful compute()");
        Matrix<S, Dynamic, Dynamic> x(b);
        for (Index col = 0; col < x.cols(); col++)
        {
            for (Index j = 0; j < m_n; j++)
            {
                const Index p = m_ipiv[size_t(j)];
                if (p != j)
                    std::swap(x(j, col), x(p, col));
                const S xj = x(j, col);
                if (xj != S(0))
                    for (Index i = j + 1; i <= (std::min)(m_n - 1, j + m_kl); i++)
                        x(i, col) -= ab(i, j) * xj;
            }
            for (Index j = m_n - 1; j >= 0; j--)
            {
                x(j, col) /= ab(j, j);
                const S xj = x(j, col);
                if (xj != S(0))
                    for (Index i = (std::max)(Index(0), j - m_ku - m_kl); i < j; i++)
                        x(i, col) -= ab(i, j) * xj;
            }
        }
        return x;
    }
};

// Eigen::SelfAdjointEigenSolver as the reference's tests use it for "true" eigenvalues (test/Example1.cpp:40-41 ...): cyclic Jacobi on the
// lower triangle, eigenvalues ascending with matching eigenvector columns.  Real symmetric matrices only.
template <typename MatrixType>
class SelfAdjointEigenSolver
{
    using S = typename MatrixType::Scalar;
    Matrix<S, Dynamic, 1> m_vals;
    Matrix<S, Dynamic, Dynamic> m_vecs;

public:
    SelfAdjointEigenSolver() {}
    template <typename Derived>
    explicit SelfAdjointEigenSolver(const MatrixBase<Derived>& A) { compute(A); }
    template <typename Derived>
    SelfAdjointEigenSolver& compute(const MatrixBase<Derived>& A)
    {
        const Index n = A.rows();
        Matrix<S, Dynamic, Dynamic> a(n, n), v = Matrix<S, Dynamic, Dynamic>::Identity(n, n);
        for (Index j = 0; j < n; j++)
            for (Index i = 0; i < n; i++)
                a(i, j) = i >= j ? A.coeff(i, j) : A.coeff(j, i);
        for (int sweep = 0; sweep < 100; sweep++)
        {
            S off(0);
            for (Index q = 1; q < n; q++)
                for (Index p = 0; p < q; p++)
                    off += a(p, q) * a(p, q);
            if (off <= S(1e-32) * a.squaredNorm() || off == S(0))
                break;
            for (Index p = 0; p < n - 1; p++)
                for (Index q = p + 1; q < n; q++)
                {
                    if (a(p, q) == S(0))
                        continue;
                    const S theta = (a(q, q) - a(p, p)) / (S(2) * a(p, q));
                    const S t = (theta >= S(0) ? S(1) : S(-1)) / (std::abs(theta) + std::sqrt(theta * theta + S(1)));
                    const S c = S(1) / std::sqrt(t * t + S(1)), sn = t * c;
                    for (Index k = 0; k < n; k++)
                    {
                        const S akp = a(k, p), akq = a(k, q);
                        a(k, p) = c * akp - sn * akq;
                        a(k, q) = sn * akp + c * akq;
                    }
                    for (Index k = 0; k < n; k++)
                    {
                        const S apk = a(p, k), aqk = a(q, k);
                        a(p, k) = c * apk - sn * aqk;
                        a(q, k) = sn * apk + c * aqk;
                    }
                    for (Index k = 0; k < n; k++)
                    {
                        const S vkp = v(k, p), vkq = v(k, q);
                        v(k, p) = c * vkp - sn * vkq;
                        v(k, q) = sn * vkp + c * vkq;
                    }
                }
        }
        std::vector<Index> order(static_cast<size_t>(n));
        for (Index i = 0; i < n; i++)
            order[size_t(i)] = i;
        std::stable_sort(order.begin(), order.end(), [&](Index x, Index y) { return a(x, x) < a(y, y); });
        m_vals.resize(n);
        m_vecs.resize(n, n);
        for (Index j = 0; j < n; j++)
        {
            m_vals[j] = a(order[size_t(j)], order[size_t(j)]);
            for (Index i = 0; i < n; i++)
                m_vecs(i, j) = v(i, order[size_t(j)]);
        }
        return *this;
    }
    const Matrix<S, Dynamic, 1>& eigenvalues() const { return m_vals; }
    const Matrix<S, Dynamic, Dynamic>& eigenvectors() const { return m_vecs; }
    ComputationInfo info() const { return Success; }
};

// Eigen::HouseholderQR (test/QR.cpp:145-146 takes householderQ() as the reference Q): Householder reflections, Q accumulated explicitly
template <typename MatrixType>
class HouseholderQR
{
    using S = typename MatrixType::Scalar;
    Matrix<S, Dynamic, Dynamic> m_q, m_r;

public:
    template <typename Derived>
    explicit HouseholderQR(const MatrixBase<Derived>& A)
    {
        const Index m = A.rows(), n = A.cols();
        m_r = A;
        m_q = Matrix<S, Dynamic, Dynamic>::Identity(m, m);
        for (Index k = 0; k < (std::min)(m - 1, n); k++)
        {
            Matrix<S, Dynamic, 1> x(m - k), ess(m - k - 1);
            for (Index i = k; i < m; i++)
                x[i - k] = m_r(i, k);
            S tau;
            typename NumTraits<S>::Real beta;
            x.makeHouseholder(ess, tau, beta);
            if (tau == S(0))
                continue;
            // H = I - tau v v^H with v = (1, ess); R <- H R, Q <- Q H
            auto vi = [&](Index i) { return i == 0 ? S(1) : ess[i - 1]; };
            for (Index j = 0; j < n; j++)
            {
                S dot(0);
                for (Index i = 0; i < m - k; i++)
                    dot += numext::conj(vi(i)) * m_r(k + i, j);
                for (Index i = 0; i < m - k; i++)
                    m_r(k + i, j) -= tau * vi(i) * dot;
            }
            for (Index i = 0; i < m; i++)
            {
                S dot(0);
                for (Index j = 0; j < m - k; j++)
                    dot += m_q(i, k + j) * vi(j);
                for (Index j = 0; j < m - k; j++)
                    m_q(i, k + j) -= numext::conj(tau) * dot * numext::conj(vi(j));
            }
        }
    }
    const Matrix<S, Dynamic, Dynamic>& householderQ() const { return m_q; }
    Matrix<S, Dynamic, Dynamic> matrixQR() const { return m_r; }
};

using MatrixXd = Matrix<double, Dynamic, Dynamic>;
using VectorXd = Matrix<double, Dynamic, 1>;
using MatrixXf = Matrix<float, Dynamic, Dynamic>;
using VectorXf = Matrix<float, Dynamic, 1>;
using MatrixXcd = Matrix<std::complex<double>, Dynamic, Dynamic>;
using VectorXcd = Matrix<std::complex<double>, Dynamic, 1>;
using VectorXi = Matrix<int, Dynamic, 1>;
using ArrayXd = Array<double, Dynamic, 1>;

}  // namespace Eigen

// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product path.
//
// CPU restatement (Eigen-free, C++17) of the implicitly restarted Lanczos / Arnoldi drivers of
// yixuan/spectra @ db1d5cc: Arnoldi.h, Lanczos.h, HermEigsBase.h (a.k.a. SymEigsBase),
// GenEigsBase.h, SelectionRule.h, SimpleRandom.h, MatOp/internal/ArnoldiOp.h (B = I),
// MatOp/Sparse{Sym,Gen}MatProd.h, SymEigsShiftSolver.h (back-transform only).
// See dense.hpp for the parity statement.  File:line citations are relative to
// /root/reference/include/Spectra/.
#pragma once

#include <functional>
#include <utility>

#include "dense.hpp"

#ifdef _OPENMP
#include <omp.h>
#endif

namespace oracle {

// ---------------------------------------------------------------------------------------------
// Util/SelectionRule.h :33-58
// ---------------------------------------------------------------------------------------------
enum class SortRule : int
{
    LargestMagn = 0,
    LargestReal,
    LargestImag,
    LargestAlge,
    SmallestMagn,
    SmallestReal,
    SmallestImag,
    SmallestAlge,
    BothEnds
};

// Util/CompInfo.h :17-30
enum class CompInfo : int
{
    Successful = 0,
    NotComputed,
    NotConverging,
    NumericalIssue
};

// SortingTarget<double, Rule>::get  (SelectionRule.h:68-192)
inline double sorting_target(SortRule rule, double v)
{
    switch (rule)
    {
        case SortRule::LargestMagn:
            return -std::abs(v);
        case SortRule::LargestAlge:
        case SortRule::BothEnds:
            return -v;
        case SortRule::SmallestMagn:
            return std::abs(v);
        case SortRule::SmallestAlge:
            return v;
        default:
            throw std::invalid_argument("incompatible selection rule");
    }
}
inline double sorting_target(SortRule rule, const Complex& v)
{
    switch (rule)
    {
        case SortRule::LargestMagn:
            return -std::abs(v);
        case SortRule::LargestReal:
            return -v.real();
        case SortRule::LargestImag:
            return -std::abs(v.imag());
        case SortRule::SmallestMagn:
            return std::abs(v);
        case SortRule::SmallestReal:
            return v.real();
        case SortRule::SmallestImag:
            return std::abs(v.imag());
        default:
            throw std::invalid_argument("incompatible selection rule");
    }
}

// SortEigenvalue<T, Rule>  (SelectionRule.h:195-224): std::sort of the index vector (not stable).
template <typename T>
std::vector<Index> sort_eigenvalue(SortRule rule, const T* evals, Index size)
{
    std::vector<Index> ind(size);
    for (Index i = 0; i < size; i++)
        ind[i] = i;
    std::sort(ind.begin(), ind.end(), [&](Index i, Index j) { return sorting_target(rule, evals[i]) < sorting_target(rule, evals[j]); });
    return ind;
}

// argsort for real values  (SelectionRule.h:227-287)
inline std::vector<Index> argsort(SortRule selection, const double* values, Index len)
{
    std::vector<Index> ind;
    switch (selection)
    {
        case SortRule::LargestMagn:
            ind = sort_eigenvalue(SortRule::LargestMagn, values, len);
            break;
        case SortRule::BothEnds:
        case SortRule::LargestAlge:
            ind = sort_eigenvalue(SortRule::LargestAlge, values, len);
            break;
        case SortRule::SmallestMagn:
            ind = sort_eigenvalue(SortRule::SmallestMagn, values, len);
            break;
        case SortRule::SmallestAlge:
            ind = sort_eigenvalue(SortRule::SmallestAlge, values, len);
            break;
        default:
            throw std::invalid_argument("unsupported selection rule");
    }
    if (selection == SortRule::BothEnds)
    {
        std::vector<Index> ind_copy(ind);
        for (Index i = 0; i < len; i++)
        {
            if (i % 2 == 0)
                ind[i] = ind_copy[i / 2];
            else
                ind[i] = ind_copy[len - 1 - i / 2];
        }
    }
    return ind;
}

// ---------------------------------------------------------------------------------------------
// Util/SimpleRandom.h :30-123
// ---------------------------------------------------------------------------------------------
inline long next_long_rand(long seed)
{
    constexpr unsigned int m_a = 16807;
    constexpr unsigned long m_max = 2147483647L;
    unsigned long lo, hi;
    lo = (unsigned long) m_a * (unsigned long) (seed & 0xFFFFUL);
    hi = (unsigned long) m_a * (unsigned long) ((unsigned long) seed >> 16);
    lo += (hi & 0x7FFF) << 16;
    if (lo > m_max)
    {
        lo &= m_max;
        ++lo;
    }
    lo += hi >> 15;
    if (lo > m_max)
    {
        lo &= m_max;
        ++lo;
    }
    return (long) lo;
}

class SimpleRandom
{
    long m_rand;

public:
    explicit SimpleRandom(unsigned long init_seed)
    {
        constexpr unsigned long m_max = 2147483647L;
        m_rand = init_seed ? (init_seed & m_max) : 1;
    }
    double random()
    {
        constexpr unsigned long m_max = 2147483647L;
        m_rand = next_long_rand(m_rand);
        return double(m_rand) / double(m_max) - 0.5;
    }
    void random_vec(double* v, Index len)
    {
        for (Index i = 0; i < len; i++)
            v[i] = random();
    }
};

// ---------------------------------------------------------------------------------------------
// Operator: MatOp/SparseSymMatProd.h:83-88, MatOp/SparseGenMatProd.h:82-87.
// The oracle holds an explicit full CSR; the "selfadjointView<Uplo>" expansion and the CSC ->
// CSR transpose are done by build_full_csr() below.
// ---------------------------------------------------------------------------------------------
struct CsrOp
{
    Index n = 0;
    std::vector<int64_t> rowptr;
    std::vector<int32_t> col;
    std::vector<double> val;
    int threads = 1;

    Index rows() const { return n; }
    void perform_op(const double* x, double* y) const
    {
#ifdef _OPENMP
#pragma omp parallel for num_threads(threads) schedule(static) if (threads > 1)
#endif
        for (Index i = 0; i < n; i++)
        {
            double s = 0.0;
            for (int64_t p = rowptr[i]; p < rowptr[i + 1]; p++)
                s += val[p] * x[col[p]];
            y[i] = s;
        }
    }
};

// mode: 0 = general (use every stored entry), 1 = symmetric from Lower triangle, 2 = symmetric
// from Upper triangle.  order: 0 = ColMajor (outer = column), 1 = RowMajor (outer = row).
// Result: full CSR with columns ascending inside each row, duplicates summed.
inline CsrOp build_full_csr(Index n, const int64_t* outer, const int32_t* inner, const double* val, int order, int mode)
{
    struct Ent
    {
        int32_t c;
        double v;
    };
    std::vector<std::vector<Ent>> rows(n);
    for (Index o = 0; o < n; o++)
    {
        for (int64_t p = outer[o]; p < outer[o + 1]; p++)
        {
            const Index in = inner[p];
            const Index i = (order == 0) ? in : o;  // row
            const Index j = (order == 0) ? o : in;  // col
            if (mode == 0)
                rows[i].push_back({int32_t(j), val[p]});
            else
            {
                const bool used = (mode == 1) ? (i >= j) : (i <= j);
                if (!used)
                    continue;
                rows[i].push_back({int32_t(j), val[p]});
                if (i != j)
                    rows[j].push_back({int32_t(i), val[p]});
            }
        }
    }
    CsrOp op;
    op.n = n;
    op.rowptr.assign(n + 1, 0);
    for (Index i = 0; i < n; i++)
    {
        auto& r = rows[i];
        std::stable_sort(r.begin(), r.end(), [](const Ent& a, const Ent& b) { return a.c < b.c; });
        size_t w = 0;
        for (size_t k = 0; k < r.size(); k++)
        {
            if (w > 0 && r[w - 1].c == r[k].c)
                r[w - 1].v += r[k].v;
            else
                r[w++] = r[k];
        }
        r.resize(w);
        op.rowptr[i + 1] = op.rowptr[i] + int64_t(w);
    }
    op.col.resize(op.rowptr[n]);
    op.val.resize(op.rowptr[n]);
    for (Index i = 0; i < n; i++)
    {
        int64_t p = op.rowptr[i];
        for (auto& e : rows[i])
        {
            op.col[p] = e.c;
            op.val[p] = e.v;
            p++;
        }
    }
    return op;
}

// Generic user operator (the OpType concept, SymEigsSolver.h:99-114)
struct FnOp
{
    Index n = 0;
    std::function<void(const double*, double*)> fn;
    Index rows() const { return n; }
    void perform_op(const double* x, double* y) const { fn(x, y); }
};

// ---------------------------------------------------------------------------------------------
// BLAS-1/2 used by ArnoldiOp<Op, IdentityBOp> (ArnoldiOp.h:136-155): plain dot, X^T y, sqrt(sum x^2)
// ---------------------------------------------------------------------------------------------
struct Blas
{
    int threads = 1;
    double dot(const double* x, const double* y, Index n) const
    {
        double s = 0.0;
#ifdef _OPENMP
#pragma omp parallel for num_threads(threads) reduction(+ : s) schedule(static) if (threads > 1)
#endif
        for (Index i = 0; i < n; i++)
            s += x[i] * y[i];
        return s;
    }
    double norm(const double* x, Index n) const { return std::sqrt(dot(x, x, n)); }
    // res = X[:, :k]' y
    void adjoint_product(const double* X, Index ld, Index k, const double* y, Index n, double* res) const
    {
        for (Index j = 0; j < k; j++)
            res[j] = dot(X + j * ld, y, n);
    }
    // f -= X[:, :k] c
    void sub_gemv(const double* X, Index ld, Index k, const double* c, double* f, Index n) const
    {
#ifdef _OPENMP
#pragma omp parallel for num_threads(threads) schedule(static) if (threads > 1)
#endif
        for (Index i = 0; i < n; i++)
        {
            double s = 0.0;
            for (Index j = 0; j < k; j++)
                s += X[i + j * ld] * c[j];
            f[i] -= s;
        }
    }
};

struct Stats
{
    Index reorth_passes = 0;   // correction passes inside factorize_from
    Index expand_calls = 0;    // expand_basis invocations
    Index restarts = 0;        // restart() invocations
    Index lanczos_steps = 0;   // factorize_from loop bodies executed
};

// ---------------------------------------------------------------------------------------------
// LinAlg/Arnoldi.h :27-343
// ---------------------------------------------------------------------------------------------
template <typename Op>
class Arnoldi
{
protected:
    const double m_near_0 = kNear0;
    const double m_eps = kEps;
    const Op& m_op;
    Blas m_blas;
    const Index m_n;
    const Index m_m;
    Index m_k = 0;
    Mat m_fac_V, m_fac_H;
    std::vector<double> m_fac_f;
    double m_beta = 0.0;

public:
    Stats stats;
    Index op_limit = -1;  // >= 0: stop factorising once op_counter reaches this (bench sampling only)

protected:
    // :66-115
    void expand_basis(const double* V, Index ncolV, Index seed, std::vector<double>& f, double& fnorm, Index& op_counter)
    {
        stats.expand_calls++;
        std::vector<double> v(m_n), Vf(ncolV);
        for (Index iter = 0; iter < 5; iter++)
        {
            SimpleRandom rng(seed + 123 * iter);
            if (iter == 0)
            {
                rng.random_vec(v.data(), m_n);
                m_op.perform_op(v.data(), f.data());
                op_counter++;
            }
            else
            {
                rng.random_vec(f.data(), m_n);
            }
            m_blas.adjoint_product(V, m_n, ncolV, f.data(), m_n, Vf.data());
            m_blas.sub_gemv(V, m_n, ncolV, Vf.data(), f.data(), m_n);
            fnorm = m_blas.norm(f.data(), m_n);

            m_blas.adjoint_product(V, m_n, ncolV, f.data(), m_n, Vf.data());
            double ortho_err = 0.0;
            for (Index j = 0; j < ncolV; j++)
                ortho_err = std::max(ortho_err, std::abs(Vf[j]));
            int count = 0;
            while (count < 3 && ortho_err >= m_eps * fnorm)
            {
                m_blas.sub_gemv(V, m_n, ncolV, Vf.data(), f.data(), m_n);
                fnorm = m_blas.norm(f.data(), m_n);
                m_blas.adjoint_product(V, m_n, ncolV, f.data(), m_n, Vf.data());
                ortho_err = 0.0;
                for (Index j = 0; j < ncolV; j++)
                    ortho_err = std::max(ortho_err, std::abs(Vf[j]));
                count++;
            }
            if (ortho_err < m_eps * fnorm)
                return;
        }
    }

    static double max_abs(const double* v, Index n)
    {
        // cwiseAbs().maxCoeff() on an empty vector is not reached on this path
        double m = 0.0;
        for (Index i = 0; i < n; i++)
            m = std::max(m, std::abs(v[i]));
        return m;
    }

    // keep the leading from_k x from_k block of H  (Lanczos.h:85-86, Arnoldi.h:219-220)
    void trim_H(Index from_k)
    {
        for (Index j = 0; j < m_m; j++)
            for (Index i = 0; i < m_m; i++)
                if (i >= from_k || j >= from_k)
                    m_fac_H(i, j) = 0.0;
    }

public:
    Arnoldi(const Op& op, Index m, int threads = 1) : m_op(op), m_n(op.rows()), m_m(m) { m_blas.threads = threads; }
    virtual ~Arnoldi() {}

    const Mat& matrix_V() const { return m_fac_V; }
    const Mat& matrix_H() const { return m_fac_H; }
    const std::vector<double>& vector_f() const { return m_fac_f; }
    double f_norm() const { return m_beta; }
    Index subspace_dim() const { return m_k; }

    // :136-195
    void init(const double* v0, Index& op_counter)
    {
        m_fac_V.resize(m_n, m_m);
        m_fac_H.resize(m_m, m_m);
        m_fac_f.assign(m_n, 0.0);

        const double v0norm = m_blas.norm(v0, m_n);
        if (v0norm < m_near_0)
            throw std::invalid_argument("initial residual vector cannot be zero");

        double* v = m_fac_V.col(0);
        m_op.perform_op(v0, v);
        op_counter++;

        const double vnorm = m_blas.norm(v, m_n);
        if (vnorm < m_near_0)
        {
            for (Index i = 0; i < m_n; i++)
                v[i] = v0[i] / v0norm;
        }
        else
        {
            for (Index i = 0; i < m_n; i++)
                v[i] /= vnorm;
        }

        std::vector<double> w(m_n);
        m_op.perform_op(v, w.data());
        op_counter++;

        m_fac_H(0, 0) = m_blas.dot(v, w.data(), m_n);
        const double h00 = m_fac_H(0, 0);
        for (Index i = 0; i < m_n; i++)
            m_fac_f[i] = w[i] - v[i] * h00;

        if (max_abs(m_fac_f.data(), m_n) < m_eps * std::abs(h00))
        {
            std::fill(m_fac_f.begin(), m_fac_f.end(), 0.0);
            m_beta = 0.0;
        }
        else
        {
            m_beta = m_blas.norm(m_fac_f.data(), m_n);
        }
        m_k = 1;
    }

    // :198-295
    virtual void factorize_from(Index from_k, Index to_m, Index& op_counter)
    {
        if (to_m <= from_k)
            return;
        if (from_k > m_k)
        {
            std::string msg = "Arnoldi: from_k (= " + std::to_string(from_k) + ") is larger than the current subspace dimension (= " + std::to_string(m_k) + ")";
            throw std::invalid_argument(msg);
        }
        const double beta_thresh = m_eps * std::sqrt(double(m_n));
        std::vector<double> Vf(to_m), w(m_n);
        trim_H(from_k);

        for (Index i = from_k; i <= to_m - 1; i++)
        {
            if (op_limit >= 0 && op_counter >= op_limit)
                return;
            stats.lanczos_steps++;
            bool restart = false;
            if (m_beta < m_near_0)
            {
                expand_basis(m_fac_V.data(), i, 2 * i, m_fac_f, m_beta, op_counter);
                restart = true;
            }
            double* vi = m_fac_V.col(i);
            for (Index r = 0; r < m_n; r++)
                vi[r] = m_fac_f[r] / m_beta;
            m_fac_H(i, i - 1) = restart ? 0.0 : m_beta;

            m_op.perform_op(vi, w.data());
            op_counter++;

            const Index i1 = i + 1;
            double* h = &m_fac_H(0, i);
            m_blas.adjoint_product(m_fac_V.data(), m_n, i1, w.data(), m_n, h);

            // f <- w - V * h
            for (Index r = 0; r < m_n; r++)
                m_fac_f[r] = w[r];
            m_blas.sub_gemv(m_fac_V.data(), m_n, i1, h, m_fac_f.data(), m_n);
            m_beta = m_blas.norm(m_fac_f.data(), m_n);

            double hnorm = 0.0;
            for (Index j = 0; j < i1; j++)
                hnorm += h[j] * h[j];
            hnorm = std::sqrt(hnorm);
            if (m_beta > 0.717 * hnorm)
                continue;

            m_blas.adjoint_product(m_fac_V.data(), m_n, i1, m_fac_f.data(), m_n, Vf.data());
            double ortho_err = max_abs(Vf.data(), i1);
            int count = 0;
            while (count < 5 && ortho_err > m_eps * m_beta)
            {
                if (m_beta < beta_thresh)
                {
                    std::fill(m_fac_f.begin(), m_fac_f.end(), 0.0);
                    m_beta = 0.0;
                    break;
                }
                stats.reorth_passes++;
                m_blas.sub_gemv(m_fac_V.data(), m_n, i1, Vf.data(), m_fac_f.data(), m_n);
                for (Index j = 0; j < i1; j++)
                    h[j] += Vf[j];
                m_beta = m_blas.norm(m_fac_f.data(), m_n);
                m_blas.adjoint_product(m_fac_V.data(), m_n, i1, m_fac_f.data(), m_n, Vf.data());
                ortho_err = max_abs(Vf.data(), i1);
                count++;
            }
        }
        m_k = to_m;
    }

    // :299-310
    void compress_H(const DoubleShiftQR& decomp)
    {
        decomp.matrix_QtHQ(m_fac_H);
        m_k -= 2;
    }
    void compress_H(const UpperHessenbergQR& decomp)
    {
        decomp.matrix_QtHQ(m_fac_H);
        m_k--;
    }

    // :320-340
    void compress_V(const Mat& Q)
    {
        Mat Vs(m_n, m_k + 1);
        const int threads = m_blas.threads;
        for (Index i = 0; i < m_k; i++)
        {
            const Index nnz = m_m - m_k + i + 1;
            const double* q = Q.col(i);
            double* dst = Vs.col(i);
#ifdef _OPENMP
#pragma omp parallel for num_threads(threads) schedule(static) if (threads > 1)
#endif
            for (Index r = 0; r < m_n; r++)
            {
                double s = 0.0;
                for (Index j = 0; j < nnz; j++)
                    s += m_fac_V(r, j) * q[j];
                dst[r] = s;
            }
        }
        {
            const double* q = Q.col(m_k);
            double* dst = Vs.col(m_k);
#ifdef _OPENMP
#pragma omp parallel for num_threads(threads) schedule(static) if (threads > 1)
#endif
            for (Index r = 0; r < m_n; r++)
            {
                double s = 0.0;
                for (Index j = 0; j < m_m; j++)
                    s += m_fac_V(r, j) * q[j];
                dst[r] = s;
            }
        }
        for (Index j = 0; j <= m_k; j++)
            std::copy(Vs.col(j), Vs.col(j) + m_n, m_fac_V.col(j));

        const double qmk = Q(m_m - 1, m_k - 1);
        const double hk = m_fac_H(m_k, m_k - 1);
        const double* vk = m_fac_V.col(m_k);
        for (Index r = 0; r < m_n; r++)
            m_fac_f[r] = m_fac_f[r] * qmk + vk[r] * hk;
        m_beta = m_blas.norm(m_fac_f.data(), m_n);
    }
};

// ---------------------------------------------------------------------------------------------
// LinAlg/Lanczos.h :27-218
// ---------------------------------------------------------------------------------------------
template <typename Op>
class Lanczos : public Arnoldi<Op>
{
    using Base = Arnoldi<Op>;
    using Base::m_beta;
    using Base::m_blas;
    using Base::m_eps;
    using Base::m_fac_f;
    using Base::m_fac_H;
    using Base::m_fac_V;
    using Base::m_k;
    using Base::m_m;
    using Base::m_n;
    using Base::m_near_0;
    using Base::m_op;

public:
    using Base::op_limit;
    using Base::stats;
    Lanczos(const Op& op, Index m, int threads = 1) : Base(op, m, threads) {}

    // :62-187
    void factorize_from(Index from_k, Index to_m, Index& op_counter) override
    {
        if (to_m <= from_k)
            return;
        if (from_k > m_k)
        {
            std::string msg = "Lanczos: from_k (= " + std::to_string(from_k) + ") is larger than the current subspace dimension (= " + std::to_string(m_k) + ")";
            throw std::invalid_argument(msg);
        }
        const double beta_thresh = m_eps * std::sqrt(double(m_n));
        const double eps_sqrt = std::sqrt(m_eps);
        std::vector<double> Vf(to_m), w(m_n);
        this->trim_H(from_k);

        for (Index i = from_k; i <= to_m - 1; i++)
        {
            if (op_limit >= 0 && op_counter >= op_limit)
                return;
            stats.lanczos_steps++;
            bool restart = (m_beta < m_near_0);
            double* v = m_fac_V.col(i);
            if (!restart)
            {
                for (Index r = 0; r < m_n; r++)
                    v[r] = m_fac_f[r] / m_beta;
                if (m_beta < eps_sqrt)
                {
                    const double Viv = m_blas.dot(m_fac_V.col(i - 1), v, m_n);
                    restart = (std::abs(Viv) > eps_sqrt);
                }
            }
            if (restart)
            {
                this->expand_basis(m_fac_V.data(), i, 2 * i, m_fac_f, m_beta, op_counter);
                for (Index r = 0; r < m_n; r++)
                    v[r] = m_fac_f[r] / m_beta;
            }

            m_fac_H(i, i - 1) = restart ? 0.0 : m_beta;
            m_fac_H(i - 1, i) = m_fac_H(i, i - 1);

            m_op.perform_op(v, w.data());
            op_counter++;

            if (!restart)
            {
                const double hi = m_fac_H(i, i - 1);
                const double* vp = m_fac_V.col(i - 1);
                for (Index r = 0; r < m_n; r++)
                    w[r] -= hi * vp[r];
            }

            m_fac_H(i, i) = m_blas.dot(v, w.data(), m_n);
            const double hii = m_fac_H(i, i);
            for (Index r = 0; r < m_n; r++)
                m_fac_f[r] = w[r] - hii * v[r];
            m_beta = m_blas.norm(m_fac_f.data(), m_n);

            const Index i1 = i + 1;
            m_blas.adjoint_product(m_fac_V.data(), m_n, i1, m_fac_f.data(), m_n, Vf.data());
            double ortho_err = Base::max_abs(Vf.data(), i1);
            int count = 0;
            while (count < 5 && ortho_err > m_eps * m_beta)
            {
                if (m_beta < beta_thresh)
                {
                    std::fill(m_fac_f.begin(), m_fac_f.end(), 0.0);
                    m_beta = 0.0;
                    break;
                }
                stats.reorth_passes++;
                m_blas.sub_gemv(m_fac_V.data(), m_n, i1, Vf.data(), m_fac_f.data(), m_n);
                m_fac_H(i - 1, i) += Vf[i - 1];
                m_fac_H(i, i - 1) = m_fac_H(i - 1, i);
                m_fac_H(i, i) += Vf[i];
                m_beta = m_blas.norm(m_fac_f.data(), m_n);
                m_blas.adjoint_product(m_fac_V.data(), m_n, i1, m_fac_f.data(), m_n, Vf.data());
                ortho_err = Base::max_abs(Vf.data(), i1);
                count++;
            }
        }
        m_k = to_m;
    }

    // :198-202
    void compress_H(const TridiagQR& decomp)
    {
        decomp.matrix_QtHQ(m_fac_H);
        m_k--;
    }
};

// ---------------------------------------------------------------------------------------------
// HermEigsBase.h :43-479 (== SymEigsBase) + SymEigsSolver.h:133-160 + SymEigsShiftSolver.h:148-196
// ---------------------------------------------------------------------------------------------
template <typename Op>
class SymEigsSolver
{
protected:
    const Op& m_op;
    const Index m_n, m_nev, m_ncv;
    Index m_nmatop = 0, m_niter = 0;
    Lanczos<Op> m_fac;
    std::vector<double> m_ritz_val;
    Mat m_ritz_vec;
    std::vector<double> m_ritz_est;
    std::vector<char> m_ritz_conv;
    CompInfo m_info = CompInfo::NotComputed;
    bool m_shift_mode = false;  // SymEigsShiftSolver: lambda = 1/nu + sigma in sort_ritzpair
    double m_sigma = 0.0;
    int m_threads = 1;

    // :105-155
    void restart(Index k, SortRule selection)
    {
        if (k >= m_ncv)
            return;
        m_fac.stats.restarts++;
        TridiagQR decomp(m_ncv);
        Mat Q(m_ncv, m_ncv);
        Q.set_identity();

        const Index nshift = m_ncv - k;
        std::vector<double> shifts(m_ritz_val.end() - nshift, m_ritz_val.end());
        std::sort(shifts.begin(), shifts.end(), [](const double& v1, const double& v2) { return std::abs(v1) > std::abs(v2); });

        for (Index i = 0; i < nshift; i++)
        {
            decomp.compute(m_fac.matrix_H(), shifts[i]);
            decomp.apply_YQ(Q);
            m_fac.compress_H(decomp);
        }
        m_fac.compress_V(Q);
        m_fac.factorize_from(k, m_ncv, m_nmatop);
        retrieve_ritzpair(selection);
    }

    // :158-175
    Index num_converged(double tol)
    {
        const double eps23 = std::pow(kEps, 2.0 / 3.0);
        Index cnt = 0;
        for (Index i = 0; i < m_nev; i++)
        {
            const double thresh = tol * std::max(std::abs(m_ritz_val[i]), eps23);
            const double resid = std::abs(m_ritz_est[i]) * m_fac.f_norm();
            m_ritz_conv[i] = (resid < thresh);
            cnt += m_ritz_conv[i] ? 1 : 0;
        }
        return cnt;
    }

    // :178-202
    Index nev_adjusted(Index nconv)
    {
        Index nev_new = m_nev;
        for (Index i = m_nev; i < m_ncv; i++)
            if (std::abs(m_ritz_est[i]) < kNear0)
                nev_new++;
        nev_new += std::min(nconv, (m_ncv - nev_new) / 2);
        if (nev_new == 1 && m_ncv >= 6)
            nev_new = m_ncv / 2;
        else if (nev_new == 1 && m_ncv > 2)
            nev_new = 2;
        if (nev_new > m_ncv - 1)
            nev_new = m_ncv - 1;
        return nev_new;
    }

    // :205-224
    void retrieve_ritzpair(SortRule selection)
    {
        TridiagEigen decomp(m_fac.matrix_H());
        const std::vector<double>& evals = decomp.eigenvalues();
        const Mat& evecs = decomp.eigenvectors();
        std::vector<Index> ind = argsort(selection, evals.data(), m_ncv);
        for (Index i = 0; i < m_ncv; i++)
        {
            m_ritz_val[i] = evals[ind[i]];
            m_ritz_est[i] = evecs(m_ncv - 1, ind[i]);
        }
        for (Index i = 0; i < m_nev; i++)
            std::copy(evecs.col(ind[i]), evecs.col(ind[i]) + m_ncv, m_ritz_vec.col(i));
    }

    // :229-251 (+ SymEigsShiftSolver.h:163-169)
    void sort_ritzpair(SortRule sort_rule)
    {
        if (m_shift_mode)
            for (Index i = 0; i < m_nev; i++)
                m_ritz_val[i] = 1.0 / m_ritz_val[i] + m_sigma;

        if ((sort_rule != SortRule::LargestAlge) && (sort_rule != SortRule::LargestMagn) && (sort_rule != SortRule::SmallestAlge) && (sort_rule != SortRule::SmallestMagn))
            throw std::invalid_argument("unsupported sorting rule");
        std::vector<Index> ind = argsort(sort_rule, m_ritz_val.data(), m_nev);
        std::vector<double> new_ritz_val(m_ncv, 0.0);  // tail is uninitialised in the reference
        Mat new_ritz_vec(m_ncv, m_nev);
        std::vector<char> new_ritz_conv(m_nev);
        for (Index i = 0; i < m_nev; i++)
        {
            new_ritz_val[i] = m_ritz_val[ind[i]];
            std::copy(m_ritz_vec.col(ind[i]), m_ritz_vec.col(ind[i]) + m_ncv, new_ritz_vec.col(i));
            new_ritz_conv[i] = m_ritz_conv[ind[i]];
        }
        m_ritz_val.swap(new_ritz_val);
        std::swap(m_ritz_vec, new_ritz_vec);
        m_ritz_conv.swap(new_ritz_conv);
    }

public:
    // :257-272
    SymEigsSolver(const Op& op, Index nev, Index ncv, int threads = 1) :
        m_op(op), m_n(op.rows()), m_nev(nev), m_ncv(ncv > m_n ? m_n : ncv), m_fac(op, m_ncv, threads), m_threads(threads)
    {
        if (nev < 1 || nev > m_n - 1)
            throw std::invalid_argument("nev must satisfy 1 <= nev <= n - 1, n is the size of matrix");
        if (ncv <= nev || ncv > m_n)
            throw std::invalid_argument("ncv must satisfy nev < ncv <= n, n is the size of matrix");
    }

    void set_shift_mode(double sigma)
    {
        m_shift_mode = true;
        m_sigma = sigma;
    }
    void set_op_limit(Index lim) { m_fac.op_limit = lim; }

    // :309-328
    void init(const double* init_resid)
    {
        m_ritz_val.assign(m_ncv, 0.0);
        m_ritz_vec.resize(m_ncv, m_nev);
        m_ritz_est.assign(m_ncv, 0.0);
        m_ritz_conv.assign(m_nev, 0);
        m_nmatop = 0;
        m_niter = 0;
        m_fac.init(init_resid, m_nmatop);
    }
    // :337-342
    void init()
    {
        SimpleRandom rng(0);
        std::vector<double> init_resid(m_n);
        rng.random_vec(init_resid.data(), m_n);
        init(init_resid.data());
    }

    // :366-390
    Index compute(SortRule selection = SortRule::LargestMagn, Index maxit = 1000, double tol = 1e-10, SortRule sorting = SortRule::LargestAlge)
    {
        m_fac.factorize_from(1, m_ncv, m_nmatop);
        if (m_fac.op_limit >= 0 && m_nmatop >= m_fac.op_limit)
            return 0;  // bounded bench sample: stop here
        retrieve_ritzpair(selection);
        Index i, nconv = 0, nev_adj;
        for (i = 0; i < maxit; i++)
        {
            nconv = num_converged(tol);
            if (nconv >= m_nev)
                break;
            nev_adj = nev_adjusted(nconv);
            restart(nev_adj, selection);
            if (m_fac.op_limit >= 0 && m_nmatop >= m_fac.op_limit)
                return 0;
        }
        sort_ritzpair(sorting);
        m_niter += (i + 1);
        m_info = (nconv >= m_nev) ? CompInfo::Successful : CompInfo::NotConverging;
        return std::min(m_nev, nconv);
    }

    CompInfo info() const { return m_info; }
    Index num_iterations() const { return m_niter; }
    Index num_operations() const { return m_nmatop; }
    const Stats& stats() const { return m_fac.stats; }
    const Lanczos<Op>& factorization() const { return m_fac; }
    Lanczos<Op>& factorization() { return m_fac; }
    Index& op_counter() { return m_nmatop; }

    // :417-436
    std::vector<double> eigenvalues() const
    {
        std::vector<double> res;
        for (Index i = 0; i < m_nev; i++)
            if (m_ritz_conv[i])
                res.push_back(m_ritz_val[i]);
        return res;
    }

    // :447-470  (n x nvec, column-major)
    Mat eigenvectors(Index nvec) const
    {
        Index nconv = 0;
        for (Index i = 0; i < m_nev; i++)
            nconv += m_ritz_conv[i] ? 1 : 0;
        nvec = std::min(nvec, nconv);
        Mat res(m_n, nvec);
        if (!nvec)
            return res;
        Mat ritz_vec_conv(m_ncv, nvec);
        Index j = 0;
        for (Index i = 0; i < m_nev && j < nvec; i++)
        {
            if (m_ritz_conv[i])
            {
                std::copy(m_ritz_vec.col(i), m_ritz_vec.col(i) + m_ncv, ritz_vec_conv.col(j));
                j++;
            }
        }
        const Mat& V = m_fac.matrix_V();
        const int threads = m_threads;
        for (Index c = 0; c < nvec; c++)
        {
            double* dst = res.col(c);
            const double* s = ritz_vec_conv.col(c);
#ifdef _OPENMP
#pragma omp parallel for num_threads(threads) schedule(static) if (threads > 1)
#endif
            for (Index r = 0; r < m_n; r++)
            {
                double acc = 0.0;
                for (Index k = 0; k < m_ncv; k++)
                    acc += V(r, k) * s[k];
                dst[r] = acc;
            }
        }
        return res;
    }
    Mat eigenvectors() const { return eigenvectors(m_nev); }
};

// ---------------------------------------------------------------------------------------------
// GenEigsBase.h :43-612 (real specialisation) + GenEigsSolver.h:158-186
// ---------------------------------------------------------------------------------------------
template <typename Op>
class GenEigsSolver
{
protected:
    const Op& m_op;
    const Index m_n, m_nev, m_ncv;
    Index m_nmatop = 0, m_niter = 0;
    Arnoldi<Op> m_fac;
    std::vector<Complex> m_ritz_val;
    CMat m_ritz_vec;
    std::vector<Complex> m_ritz_est;
    std::vector<char> m_ritz_conv;
    CompInfo m_info = CompInfo::NotComputed;
    int m_threads = 1;

    static bool is_complex(const Complex& v) { return v.imag() != 0.0; }
    static bool is_conj(const Complex& v1, const Complex& v2) { return v1 == std::conj(v2); }

    // RestartArnoldi<double,...>::run  :60-107
    void restart_arnoldi_run(Index k, Mat& Q)
    {
        const Index ncv = m_ncv;
        DoubleShiftQR decomp_ds(ncv);
        UpperHessenbergQR decomp_hb(ncv);
        for (Index i = k; i < ncv; i++)
        {
            // the reference reads ritz_val[i + 1] unguarded (:70); conjugates are adjacent
            if (is_complex(m_ritz_val[i]) && i + 1 < ncv && is_conj(m_ritz_val[i], m_ritz_val[i + 1]))
            {
                const double s = 2.0 * m_ritz_val[i].real();
                const double t = std::norm(m_ritz_val[i]);
                decomp_ds.compute(m_fac.matrix_H(), s, t);
                decomp_ds.apply_YQ(Q);
                m_fac.compress_H(decomp_ds);
                i++;
            }
            else
            {
                decomp_hb.compute(m_fac.matrix_H(), m_ritz_val[i].real());
                decomp_hb.apply_YQ(Q);
                m_fac.compress_H(decomp_hb);
            }
        }
    }

    // :204-222
    void restart(Index k, SortRule selection)
    {
        if (k >= m_ncv)
            return;
        m_fac.stats.restarts++;
        Mat Q(m_ncv, m_ncv);
        Q.set_identity();
        restart_arnoldi_run(k, Q);
        m_fac.compress_V(Q);
        m_fac.factorize_from(k, m_ncv, m_nmatop);
        retrieve_ritzpair(selection);
    }

    // :225-242
    Index num_converged(double tol)
    {
        const double eps23 = std::pow(kEps, 2.0 / 3.0);
        Index cnt = 0;
        for (Index i = 0; i < m_nev; i++)
        {
            const double thresh = tol * std::max(std::abs(m_ritz_val[i]), eps23);
            const double resid = std::abs(m_ritz_est[i]) * m_fac.f_norm();
            m_ritz_conv[i] = (resid < thresh);
            cnt += m_ritz_conv[i] ? 1 : 0;
        }
        return cnt;
    }

    // :245-277
    Index nev_adjusted(Index nconv)
    {
        Index nev_new = m_nev;
        for (Index i = m_nev; i < m_ncv; i++)
            if (std::abs(m_ritz_est[i]) < kNear0)
                nev_new++;
        nev_new += std::min(nconv, (m_ncv - nev_new) / 2);
        if (nev_new == 1 && m_ncv >= 6)
            nev_new = m_ncv / 2;
        else if (nev_new == 1 && m_ncv > 3)
            nev_new = 2;
        if (nev_new > m_ncv - 2)
            nev_new = m_ncv - 2;
        if (is_complex(m_ritz_val[nev_new - 1]) && is_conj(m_ritz_val[nev_new - 1], m_ritz_val[nev_new]))
            nev_new++;
        return nev_new;
    }

    static void check_rule(SortRule rule, const char* what)
    {
        switch (rule)
        {
            case SortRule::LargestMagn:
            case SortRule::LargestReal:
            case SortRule::LargestImag:
            case SortRule::SmallestMagn:
            case SortRule::SmallestReal:
            case SortRule::SmallestImag:
                return;
            default:
                throw std::invalid_argument(what);
        }
    }

    // :280-340
    void retrieve_ritzpair(SortRule selection)
    {
        UpperHessenbergEigen decomp(m_fac.matrix_H());
        const std::vector<Complex>& evals = decomp.eigenvalues();
        CMat evecs = decomp.eigenvectors();
        check_rule(selection, "unsupported selection rule");
        std::vector<Index> ind = sort_eigenvalue(selection, evals.data(), m_ncv);
        for (Index i = 0; i < m_ncv; i++)
        {
            m_ritz_val[i] = evals[ind[i]];
            m_ritz_est[i] = evecs(m_ncv - 1, ind[i]);
        }
        for (Index i = 0; i < m_nev; i++)
            for (Index r = 0; r < m_ncv; r++)
                m_ritz_vec(r, i) = evecs(r, ind[i]);
    }

    // :345-404
    void sort_ritzpair(SortRule sort_rule)
    {
        check_rule(sort_rule, "unsupported sorting rule");
        std::vector<Index> ind = sort_eigenvalue(sort_rule, m_ritz_val.data(), m_nev);
        std::vector<Complex> new_ritz_val(m_ncv);
        CMat new_ritz_vec(m_ncv, m_nev);
        std::vector<char> new_ritz_conv(m_nev);
        for (Index i = 0; i < m_nev; i++)
        {
            new_ritz_val[i] = m_ritz_val[ind[i]];
            for (Index r = 0; r < m_ncv; r++)
                new_ritz_vec(r, i) = m_ritz_vec(r, ind[i]);
            new_ritz_conv[i] = m_ritz_conv[ind[i]];
        }
        m_ritz_val.swap(new_ritz_val);
        std::swap(m_ritz_vec, new_ritz_vec);
        m_ritz_conv.swap(new_ritz_conv);
    }

public:
    // :409-424
    GenEigsSolver(const Op& op, Index nev, Index ncv, int threads = 1) :
        m_op(op), m_n(op.rows()), m_nev(nev), m_ncv(ncv > m_n ? m_n : ncv), m_fac(op, m_ncv, threads), m_threads(threads)
    {
        if (nev < 1 || nev > m_n - 2)
            throw std::invalid_argument("nev must satisfy 1 <= nev <= n - 2, n is the size of matrix");
        if (ncv < nev + 2 || ncv > m_n)
            throw std::invalid_argument("ncv must satisfy nev + 2 <= ncv <= n, n is the size of matrix");
    }
    void set_op_limit(Index lim) { m_fac.op_limit = lim; }

    // :442-475
    void init(const double* init_resid)
    {
        m_ritz_val.assign(m_ncv, Complex(0, 0));
        m_ritz_vec.resize(m_ncv, m_nev);
        m_ritz_est.assign(m_ncv, Complex(0, 0));
        m_ritz_conv.assign(m_nev, 0);
        m_nmatop = 0;
        m_niter = 0;
        m_fac.init(init_resid, m_nmatop);
    }
    void init()
    {
        SimpleRandom rng(0);
        std::vector<double> init_resid(m_n);
        rng.random_vec(init_resid.data(), m_n);
        init(init_resid.data());
    }

    // :501-525
    Index compute(SortRule selection = SortRule::LargestMagn, Index maxit = 1000, double tol = 1e-10, SortRule sorting = SortRule::LargestMagn)
    {
        m_fac.factorize_from(1, m_ncv, m_nmatop);
        if (m_fac.op_limit >= 0 && m_nmatop >= m_fac.op_limit)
            return 0;
        retrieve_ritzpair(selection);
        Index i, nconv = 0, nev_adj;
        for (i = 0; i < maxit; i++)
        {
            nconv = num_converged(tol);
            if (nconv >= m_nev)
                break;
            nev_adj = nev_adjusted(nconv);
            restart(nev_adj, selection);
            if (m_fac.op_limit >= 0 && m_nmatop >= m_fac.op_limit)
                return 0;
        }
        sort_ritzpair(sorting);
        m_niter += (i + 1);
        m_info = (nconv >= m_nev) ? CompInfo::Successful : CompInfo::NotConverging;
        return std::min(m_nev, nconv);
    }

    CompInfo info() const { return m_info; }
    Index num_iterations() const { return m_niter; }
    Index num_operations() const { return m_nmatop; }
    const Stats& stats() const { return m_fac.stats; }
    const Arnoldi<Op>& factorization() const { return m_fac; }
    Arnoldi<Op>& factorization() { return m_fac; }
    Index& op_counter() { return m_nmatop; }

    // :531-551
    std::vector<Complex> eigenvalues() const
    {
        std::vector<Complex> res;
        for (Index i = 0; i < m_nev; i++)
            if (m_ritz_conv[i])
                res.push_back(m_ritz_val[i]);
        return res;
    }

    // :561-603
    CMat eigenvectors(Index nvec) const
    {
        Index nconv = 0;
        for (Index i = 0; i < m_nev; i++)
            nconv += m_ritz_conv[i] ? 1 : 0;
        nvec = std::min(nvec, nconv);
        CMat res(m_n, nvec);
        if (!nvec)
            return res;
        CMat ritz_vec_conv(m_ncv, nvec);
        Index j = 0;
        for (Index i = 0; i < m_nev && j < nvec; i++)
        {
            if (m_ritz_conv[i])
            {
                for (Index r = 0; r < m_ncv; r++)
                    ritz_vec_conv(r, j) = m_ritz_vec(r, i);
                j++;
            }
        }
        const Mat& V = m_fac.matrix_V();
        for (Index c = 0; c < nvec; c++)
            for (Index r = 0; r < m_n; r++)
            {
                double re = 0.0, im = 0.0;
                for (Index k = 0; k < m_ncv; k++)
                {
                    re += V(r, k) * ritz_vec_conv(k, c).real();
                    im += V(r, k) * ritz_vec_conv(k, c).imag();
                }
                res(r, c) = Complex(re, im);
            }
        return res;
    }
    CMat eigenvectors() const { return eigenvectors(m_nev); }
};

}  // namespace oracle

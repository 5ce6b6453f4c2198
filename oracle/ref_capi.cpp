// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product path.
//
// C entry points over the REFERENCE'S OWN HEADERS (yixuan/spectra @ db1d5cc, compiled from
// /root/reference/include where they lie; nothing is copied into this repository).  Built by
// `make -C oracle ref` into oracle/_ref/libspectra_ref.so when /root/reference exists.
//
// Eigen 3.4, which the reference needs, is not installed in this image; <Eigen/...> resolves to the
// stand-in under oracle/eigen_standin/ (see Eigen/src/standin.h for what that means for parity: the
// reference's control flow, constants, shifts, deflation and restart rules are its own compiled
// code; dot / axpy / gemv loops and their summation order are the stand-in's).
//
// The functions mirror oracle/capi.cpp tier by tier so that tests/test_oracle_vs_reference.py can
// put the restatement (oracle/*.hpp) next to the reference on identical inputs.
#include <Spectra/SymEigsSolver.h>
#include <Spectra/GenEigsSolver.h>
#include <Spectra/HermEigsSolver.h>
#include <Spectra/SymEigsShiftSolver.h>
#include <Spectra/MatOp/SparseSymShiftSolve.h>
#include <Spectra/MatOp/SparseHermMatProd.h>
#include <Spectra/MatOp/SparseSymMatProd.h>
#include <Spectra/MatOp/SparseGenMatProd.h>
#include <Spectra/LinAlg/DoubleShiftQR.h>
#include <Spectra/LinAlg/UpperHessenbergQR.h>
#include <Spectra/LinAlg/TridiagEigen.h>
#include <Spectra/LinAlg/UpperHessenbergEigen.h>
#include <Spectra/LinAlg/UpperHessenbergSchur.h>
#include <Spectra/LinAlg/Givens.h>
#include <Spectra/LinAlg/Lanczos.h>
#include <Spectra/LinAlg/Arnoldi.h>
#include <Spectra/Util/SimpleRandom.h>
#include <Spectra/Util/SelectionRule.h>
#include <Spectra/Util/Version.h>

#include <chrono>
#include <cstring>
#include <string>

namespace {

using Eigen::Index;
using Matrix = Eigen::Matrix<double, Eigen::Dynamic, Eigen::Dynamic>;
using Vector = Eigen::Matrix<double, Eigen::Dynamic, 1>;
using MapConstMat = Eigen::Map<const Matrix>;
using MapMat = Eigen::Map<Matrix>;
using MapConstVec = Eigen::Map<const Vector>;
using Complex = std::complex<double>;

thread_local std::string g_err;

double now_s()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

void mat_to(const Matrix& m, double* out)
{
    if (out)
        std::memcpy(out, m.data(), sizeof(double) * size_t(m.rows() * m.cols()));
}

template <typename CM>
void cmat_to(const CM& m, double* out)
{
    if (!out)
        return;
    for (Index j = 0; j < m.cols(); j++)
        for (Index i = 0; i < m.rows(); i++)
        {
            out[2 * (i + j * m.rows())] = m(i, j).real();
            out[2 * (i + j * m.rows()) + 1] = m(i, j).imag();
        }
}

struct RefResult
{
    int64_t nconv, niter, nops;
    int32_t info;
    double seconds;
};

// a compressed matrix the caller owns (32-bit StorageIndex: Eigen's default `int`)
struct Compressed
{
    int64_t n, nnz;
    const int32_t* outer;
    const int32_t* inner;
    const double* val;
};

template <int Flags>
using SpMap = Eigen::Map<const Eigen::SparseMatrix<double, Flags, int>>;

}  // namespace

#define REF_TRY try {
#define REF_CATCH                                \
    return 0;                                    \
    }                                            \
    catch (const std::invalid_argument& e)       \
    {                                            \
        g_err = e.what();                        \
        return 1;                                \
    }                                            \
    catch (const std::logic_error& e)            \
    {                                            \
        g_err = e.what();                        \
        return 2;                                \
    }                                            \
    catch (const std::runtime_error& e)          \
    {                                            \
        g_err = e.what();                        \
        return 3;                                \
    }                                            \
    catch (const std::exception& e)              \
    {                                            \
        g_err = e.what();                        \
        return 4;                                \
    }

// dispatch on (uplo, order) for the symmetric operator and on order for the general one
#define REF_SYM_DISPATCH(uplo, order, BODY)                                                      \
    if ((uplo) == 0 && (order) == 0) { BODY(Eigen::Lower, Eigen::ColMajor) }                    \
    else if ((uplo) == 1 && (order) == 0) { BODY(Eigen::Upper, Eigen::ColMajor) }               \
    else if ((uplo) == 0 && (order) == 1) { BODY(Eigen::Lower, Eigen::RowMajor) }               \
    else if ((uplo) == 1 && (order) == 1) { BODY(Eigen::Upper, Eigen::RowMajor) }               \
    else throw std::invalid_argument("ref: uplo / order must be 0 or 1");

extern "C" {

const char* ref_last_error() { return g_err.c_str(); }

// "reference <major>.<minor>.<patch> over the Eigen stand-in"
const char* ref_version()
{
    static std::string v = std::string("spectra ") + std::to_string(SPECTRA_MAJOR_VERSION) + "." + std::to_string(SPECTRA_MINOR_VERSION) + "." +
        std::to_string(SPECTRA_PATCH_VERSION) + " (reference headers, Eigen stand-in)";
    return v.c_str();
}

// ---- Util ---------------------------------------------------------------------------------------
void ref_simple_random(uint64_t seed, int64_t n, double* out)
{
    Spectra::SimpleRandom<double> rng(seed);
    Vector v = rng.random_vec(n);
    std::memcpy(out, v.data(), sizeof(double) * size_t(n));
}

int ref_argsort(int selection, const double* values, int64_t len, int64_t* ind)
{
    REF_TRY
    Vector v(len);
    for (int64_t i = 0; i < len; i++)
        v[i] = values[i];
    auto r = Spectra::argsort(Spectra::SortRule(selection), v, len);
    for (int64_t i = 0; i < len; i++)
        ind[i] = r[i];
    REF_CATCH
}

int ref_argsort_complex(int selection, const double* values_ri, int64_t len, int64_t* ind)
{
    REF_TRY
    std::vector<Complex> c(len);
    for (int64_t i = 0; i < len; i++)
        c[i] = Complex(values_ri[2 * i], values_ri[2 * i + 1]);
    std::vector<Index> r;
    using Spectra::SortEigenvalue;
    using Spectra::SortRule;
    switch (SortRule(selection))
    {
        case SortRule::LargestMagn: { SortEigenvalue<Complex, SortRule::LargestMagn> s(c.data(), len); s.swap(r); break; }
        case SortRule::LargestReal: { SortEigenvalue<Complex, SortRule::LargestReal> s(c.data(), len); s.swap(r); break; }
        case SortRule::LargestImag: { SortEigenvalue<Complex, SortRule::LargestImag> s(c.data(), len); s.swap(r); break; }
        case SortRule::SmallestMagn: { SortEigenvalue<Complex, SortRule::SmallestMagn> s(c.data(), len); s.swap(r); break; }
        case SortRule::SmallestReal: { SortEigenvalue<Complex, SortRule::SmallestReal> s(c.data(), len); s.swap(r); break; }
        case SortRule::SmallestImag: { SortEigenvalue<Complex, SortRule::SmallestImag> s(c.data(), len); s.swap(r); break; }
        default: throw std::invalid_argument("unsupported selection rule");
    }
    for (int64_t i = 0; i < len; i++)
        ind[i] = r[i];
    REF_CATCH
}

// ---- small dense kernels ------------------------------------------------------------------------
void ref_givens(double x, double y, double* r, double* c, double* s) { Spectra::Givens<double>::compute_rotation(x, y, *r, *c, *s); }

// kind 0: TridiagQR, 1: UpperHessenbergQR
int ref_shifted_qr(int kind, int64_t m, const double* H, double shift, double* R, double* QtHQ, double* Q)
{
    REF_TRY
    MapConstMat h(H, m, m);
    Matrix q = Matrix::Identity(m, m), d;
    if (kind == 0)
    {
        Spectra::TridiagQR<double> dec(m);
        dec.compute(h, shift);
        mat_to(dec.matrix_R(), R);
        dec.matrix_QtHQ(d);
        dec.apply_YQ(q);
    }
    else
    {
        Spectra::UpperHessenbergQR<double> dec(m);
        dec.compute(h, shift);
        mat_to(dec.matrix_R(), R);
        dec.matrix_QtHQ(d);
        dec.apply_YQ(q);
    }
    mat_to(d, QtHQ);
    mat_to(q, Q);
    REF_CATCH
}

int ref_double_shift_qr(int64_t m, const double* H, double s, double t, double* QtHQ, double* Q)
{
    REF_TRY
    MapConstMat h(H, m, m);
    Spectra::DoubleShiftQR<double> dec(m);
    dec.compute(h, s, t);
    Matrix d(m, m), q = Matrix::Identity(m, m);
    dec.matrix_QtHQ(d);
    dec.apply_YQ(q);
    mat_to(d, QtHQ);
    mat_to(q, Q);
    REF_CATCH
}

int ref_tridiag_eigen(int64_t m, const double* H, double* evals, double* evecs)
{
    REF_TRY
    MapConstMat h(H, m, m);
    Spectra::TridiagEigen<double> dec(h);
    std::memcpy(evals, dec.eigenvalues().data(), sizeof(double) * size_t(m));
    mat_to(dec.eigenvectors(), evecs);
    REF_CATCH
}

int ref_hess_schur(int64_t m, const double* H, double* T, double* U)
{
    REF_TRY
    MapConstMat h(H, m, m);
    Spectra::UpperHessenbergSchur<double> dec;
    dec.compute(h);
    mat_to(dec.matrix_T(), T);
    mat_to(dec.matrix_U(), U);
    REF_CATCH
}

int ref_hess_eigen(int64_t m, const double* H, double* evals, double* evecs)
{
    REF_TRY
    MapConstMat h(H, m, m);
    Spectra::UpperHessenbergEigen<double> dec(h);
    const auto& ev = dec.eigenvalues();
    for (int64_t i = 0; i < m; i++)
    {
        evals[2 * i] = ev[i].real();
        evals[2 * i + 1] = ev[i].imag();
    }
    cmat_to(dec.eigenvectors(), evecs);
    REF_CATCH
}

// ---- operators ----------------------------------------------------------------------------------
// sym != 0: SparseSymMatProd<double, uplo, order>; else SparseGenMatProd<double, order>.  order 0 = ColMajor
// (outer = column pointers), 1 = RowMajor.  uplo 0 = Lower, 1 = Upper.
int ref_spmv(int sym, int uplo, int order, const Compressed* A, const double* x, double* y)
{
    REF_TRY
    if (sym)
    {
#define BODY(U, F)                                                  \
    SpMap<F> mat(A->n, A->n, A->nnz, A->outer, A->inner, A->val);   \
    Spectra::SparseSymMatProd<double, U, F> op(mat);                \
    op.perform_op(x, y);
        REF_SYM_DISPATCH(uplo, order, BODY)
#undef BODY
    }
    else if (order == 0)
    {
        SpMap<Eigen::ColMajor> mat(A->n, A->n, A->nnz, A->outer, A->inner, A->val);
        Spectra::SparseGenMatProd<double, Eigen::ColMajor> op(mat);
        op.perform_op(x, y);
    }
    else
    {
        SpMap<Eigen::RowMajor> mat(A->n, A->n, A->nnz, A->outer, A->inner, A->val);
        Spectra::SparseGenMatProd<double, Eigen::RowMajor> op(mat);
        op.perform_op(x, y);
    }
    REF_CATCH
}

// A(i, j) through the operator's operator() (SparseSymMatProd.h:103, SparseGenMatProd.h:102)
int ref_coeff(int sym, int order, const Compressed* A, int64_t i, int64_t j, double* out)
{
    REF_TRY
    if (order == 0)
    {
        SpMap<Eigen::ColMajor> mat(A->n, A->n, A->nnz, A->outer, A->inner, A->val);
        if (sym)
        {
            Spectra::SparseSymMatProd<double, Eigen::Lower, Eigen::ColMajor> op(mat);
            *out = op(i, j);
        }
        else
        {
            Spectra::SparseGenMatProd<double, Eigen::ColMajor> op(mat);
            *out = op(i, j);
        }
    }
    else
    {
        SpMap<Eigen::RowMajor> mat(A->n, A->n, A->nnz, A->outer, A->inner, A->val);
        if (sym)
        {
            Spectra::SparseSymMatProd<double, Eigen::Lower, Eigen::RowMajor> op(mat);
            *out = op(i, j);
        }
        else
        {
            Spectra::SparseGenMatProd<double, Eigen::RowMajor> op(mat);
            *out = op(i, j);
        }
    }
    REF_CATCH
}

// ---- factorisation tier (test/Arnoldi.cpp flow) -------------------------------------------------
// kind 0: Lanczos over SparseSymMatProd<Lower, ColMajor>; 1: Arnoldi over SparseGenMatProd<order>.
// init(v0 or SimpleRandom(0)), factorize_from(1, mid), factorize_from(mid, m).
int ref_factorize(int kind, int order, const Compressed* A, int64_t m, const double* v0, int64_t mid, double* V, double* H, double* f, double* beta,
                  int64_t* nops, double* seconds)
{
    REF_TRY
    const Index n = A->n;
    Vector init(n);
    if (v0)
        std::memcpy(init.data(), v0, sizeof(double) * size_t(n));
    else
    {
        Spectra::SimpleRandom<double> rng(0);
        rng.random_vec(init);
    }
    Index cnt = 0;
    MapConstVec v0map(init.data(), n);
    // V / H / f may be null (bench.py's bounded CPU sample only wants the time of init + the first m - 1 steps)
    auto run = [&](auto& fac) {
        const double t0 = now_s();
        fac.init(v0map, cnt);
        fac.factorize_from(1, mid, cnt);
        fac.factorize_from(mid, m, cnt);
        if (seconds)
            *seconds = now_s() - t0;
        mat_to(fac.matrix_V(), V);
        mat_to(fac.matrix_H(), H);
        if (f)
            std::memcpy(f, fac.vector_f().data(), sizeof(double) * size_t(n));
        *beta = fac.f_norm();
    };
    if (kind == 0)
    {
        SpMap<Eigen::ColMajor> mat(n, n, A->nnz, A->outer, A->inner, A->val);
        using Op = Spectra::SparseSymMatProd<double, Eigen::Lower, Eigen::ColMajor>;
        using AOp = Spectra::ArnoldiOp<Op, Spectra::IdentityBOp>;
        Op op(mat);
        Spectra::IdentityBOp bop;
        Spectra::Lanczos<AOp> fac(AOp(op, bop), m);
        run(fac);
    }
    else if (order == 0)
    {
        SpMap<Eigen::ColMajor> mat(n, n, A->nnz, A->outer, A->inner, A->val);
        using Op = Spectra::SparseGenMatProd<double, Eigen::ColMajor>;
        using AOp = Spectra::ArnoldiOp<Op, Spectra::IdentityBOp>;
        Op op(mat);
        Spectra::IdentityBOp bop;
        Spectra::Arnoldi<AOp> fac(AOp(op, bop), m);
        run(fac);
    }
    else
    {
        SpMap<Eigen::RowMajor> mat(n, n, A->nnz, A->outer, A->inner, A->val);
        using Op = Spectra::SparseGenMatProd<double, Eigen::RowMajor>;
        using AOp = Spectra::ArnoldiOp<Op, Spectra::IdentityBOp>;
        Op op(mat);
        Spectra::IdentityBOp bop;
        Spectra::Arnoldi<AOp> fac(AOp(op, bop), m);
        run(fac);
    }
    *nops = cnt;
    REF_CATCH
}

// ---- solver tier --------------------------------------------------------------------------------
// SymEigsSolver<SparseSymMatProd<double, uplo, order>>: the README / test/SymEigs.cpp flow
int ref_sym_eigs(int uplo, int order, const Compressed* A, int64_t nev, int64_t ncv, int selection, int64_t maxit, double tol, int sorting,
                 const double* init_resid, double* evals, double* evecs, RefResult* res)
{
    REF_TRY
#define BODY(U, F)                                                                                      \
    SpMap<F> mat(A->n, A->n, A->nnz, A->outer, A->inner, A->val);                                       \
    using Op = Spectra::SparseSymMatProd<double, U, F>;                                                 \
    Op op(mat);                                                                                         \
    Spectra::SymEigsSolver<Op> eigs(op, nev, ncv);                                                      \
    const double t0 = now_s();                                                                          \
    if (init_resid)                                                                                     \
        eigs.init(init_resid);                                                                          \
    else                                                                                                \
        eigs.init();                                                                                    \
    const Index nconv = eigs.compute(Spectra::SortRule(selection), maxit, tol, Spectra::SortRule(sorting)); \
    res->seconds = now_s() - t0;                                                                        \
    res->nconv = nconv;                                                                                 \
    res->niter = eigs.num_iterations();                                                                 \
    res->nops = eigs.num_operations();                                                                  \
    res->info = int32_t(eigs.info());                                                                   \
    Vector ev = eigs.eigenvalues();                                                                     \
    if (evals)                                                                                          \
        std::memcpy(evals, ev.data(), sizeof(double) * size_t(ev.size()));                              \
    if (evecs)                                                                                          \
        mat_to(eigs.eigenvectors(), evecs);
    REF_SYM_DISPATCH(uplo, order, BODY)
#undef BODY
    REF_CATCH
}

// SymEigsSolver over a user-defined operator type (the OpType concept, SymEigsSolver.h:99-130)
namespace {
struct CallbackOp
{
    using Scalar = double;
    Index n;
    void (*fn)(const double*, double*, void*);
    void* user;
    Index rows() const { return n; }
    Index cols() const { return n; }
    void perform_op(const double* x, double* y) const { fn(x, y, user); }
};
}  // namespace

int ref_sym_eigs_userop(int64_t n, void (*fn)(const double*, double*, void*), void* user, int64_t nev, int64_t ncv, int selection, int64_t maxit, double tol,
                        int sorting, const double* init_resid, double* evals, double* evecs, RefResult* res)
{
    REF_TRY
    CallbackOp op{Index(n), fn, user};
    Spectra::SymEigsSolver<CallbackOp> eigs(op, nev, ncv);
    const double t0 = now_s();
    if (init_resid)
        eigs.init(init_resid);
    else
        eigs.init();
    const Index nconv = eigs.compute(Spectra::SortRule(selection), maxit, tol, Spectra::SortRule(sorting));
    res->seconds = now_s() - t0;
    res->nconv = nconv;
    res->niter = eigs.num_iterations();
    res->nops = eigs.num_operations();
    res->info = int32_t(eigs.info());
    Vector ev = eigs.eigenvalues();
    if (evals)
        std::memcpy(evals, ev.data(), sizeof(double) * size_t(ev.size()));
    if (evecs)
        mat_to(eigs.eigenvectors(), evecs);
    REF_CATCH
}

// GenEigsSolver<SparseGenMatProd<double, order>>: evals / evecs interleaved (re, im)
int ref_gen_eigs(int order, const Compressed* A, int64_t nev, int64_t ncv, int selection, int64_t maxit, double tol, int sorting, const double* init_resid,
                 double* evals, double* evecs, RefResult* res)
{
    REF_TRY
    auto run = [&](auto& eigs) {
        const double t0 = now_s();
        if (init_resid)
            eigs.init(init_resid);
        else
            eigs.init();
        const Index nconv = eigs.compute(Spectra::SortRule(selection), maxit, tol, Spectra::SortRule(sorting));
        res->seconds = now_s() - t0;
        res->nconv = nconv;
        res->niter = eigs.num_iterations();
        res->nops = eigs.num_operations();
        res->info = int32_t(eigs.info());
        auto ev = eigs.eigenvalues();
        for (Index i = 0; i < ev.size(); i++)
        {
            evals[2 * i] = ev[i].real();
            evals[2 * i + 1] = ev[i].imag();
        }
        if (evecs)
            cmat_to(eigs.eigenvectors(), evecs);
    };
    if (order == 0)
    {
        SpMap<Eigen::ColMajor> mat(A->n, A->n, A->nnz, A->outer, A->inner, A->val);
        using Op = Spectra::SparseGenMatProd<double, Eigen::ColMajor>;
        Op op(mat);
        Spectra::GenEigsSolver<Op> eigs(op, nev, ncv);
        run(eigs);
    }
    else
    {
        SpMap<Eigen::RowMajor> mat(A->n, A->n, A->nnz, A->outer, A->inner, A->val);
        using Op = Spectra::SparseGenMatProd<double, Eigen::RowMajor>;
        Op op(mat);
        Spectra::GenEigsSolver<Op> eigs(op, nev, ncv);
        run(eigs);
    }
    REF_CATCH
}

}  // extern "C"

// ---- complex Hermitian path ------------------------------------------------------
// HermEigsSolver<SparseHermMatProd<std::complex<double>, Uplo, ColMajor>> (HermEigsSolver.h:121-122,
// MatOp/SparseHermMatProd.h:21-89).  val / init_resid / evecs are interleaved (re, im); eigenvalues are real.
namespace {
using CVector = Eigen::Matrix<Complex, Eigen::Dynamic, 1>;
using CMatrix = Eigen::Matrix<Complex, Eigen::Dynamic, Eigen::Dynamic>;

struct CallbackOpZ
{
    using Scalar = Complex;
    Index n;
    void (*fn)(const double*, double*, void*);
    void* user;
    Index rows() const { return n; }
    Index cols() const { return n; }
    void perform_op(const Complex* x, Complex* y) const { fn(reinterpret_cast<const double*>(x), reinterpret_cast<double*>(y), user); }
};

template <typename Eigs>
void run_herm(Eigs& eigs, int selection, int64_t maxit, double tol, int sorting, const double* init_resid, double* evals, double* evecs, RefResult* res)
{
    const double t0 = now_s();
    if (init_resid)
        eigs.init(reinterpret_cast<const Complex*>(init_resid));
    else
        eigs.init();
    const Index nconv = eigs.compute(Spectra::SortRule(selection), maxit, tol, Spectra::SortRule(sorting));
    res->seconds = now_s() - t0;
    res->nconv = nconv;
    res->niter = eigs.num_iterations();
    res->nops = eigs.num_operations();
    res->info = int32_t(eigs.info());
    Vector ev = eigs.eigenvalues();
    if (evals)
        std::memcpy(evals, ev.data(), sizeof(double) * size_t(ev.size()));
    if (evecs)
        cmat_to(eigs.eigenvectors(), evecs);
}
}  // namespace

extern "C" {

int ref_herm_eigs(int uplo, const Compressed* A, int64_t nev, int64_t ncv, int selection, int64_t maxit, double tol, int sorting, const double* init_resid,
                  double* evals, double* evecs, RefResult* res)
{
    REF_TRY
    Eigen::Map<const Eigen::SparseMatrix<Complex, Eigen::ColMajor, int>> mat(A->n, A->n, A->nnz, A->outer, A->inner, reinterpret_cast<const Complex*>(A->val));
    if (uplo == 0)
    {
        using Op = Spectra::SparseHermMatProd<Complex, Eigen::Lower>;
        Op op(mat);
        Spectra::HermEigsSolver<Op> eigs(op, nev, ncv);
        run_herm(eigs, selection, maxit, tol, sorting, init_resid, evals, evecs, res);
    }
    else
    {
        using Op = Spectra::SparseHermMatProd<Complex, Eigen::Upper>;
        Op op(mat);
        Spectra::HermEigsSolver<Op> eigs(op, nev, ncv);
        run_herm(eigs, selection, maxit, tol, sorting, init_resid, evals, evecs, res);
    }
    REF_CATCH
}

int ref_herm_eigs_userop(int64_t n, void (*fn)(const double*, double*, void*), void* user, int64_t nev, int64_t ncv, int selection, int64_t maxit, double tol,
                         int sorting, const double* init_resid, double* evals, double* evecs, RefResult* res)
{
    REF_TRY
    CallbackOpZ op{Index(n), fn, user};
    Spectra::HermEigsSolver<CallbackOpZ> eigs(op, nev, ncv);
    run_herm(eigs, selection, maxit, tol, sorting, init_resid, evals, evecs, res);
    REF_CATCH
}

// y = selfadjointView<Uplo>(A) x through SparseHermMatProd::perform_op
int ref_herm_spmv(int uplo, const Compressed* A, const double* x, double* y)
{
    REF_TRY
    Eigen::Map<const Eigen::SparseMatrix<Complex, Eigen::ColMajor, int>> mat(A->n, A->n, A->nnz, A->outer, A->inner, reinterpret_cast<const Complex*>(A->val));
    if (uplo == 0)
    {
        Spectra::SparseHermMatProd<Complex, Eigen::Lower> op(mat);
        op.perform_op(reinterpret_cast<const Complex*>(x), reinterpret_cast<Complex*>(y));
    }
    else
    {
        Spectra::SparseHermMatProd<Complex, Eigen::Upper> op(mat);
        op.perform_op(reinterpret_cast<const Complex*>(x), reinterpret_cast<Complex*>(y));
    }
    REF_CATCH
}

void ref_simple_random_complex(uint64_t seed, int64_t n, double* out_ri)
{
    Spectra::SimpleRandom<Complex> rng(seed);
    CVector v = rng.random_vec(n);
    std::memcpy(out_ri, v.data(), sizeof(Complex) * size_t(n));
}

}  // extern "C"

// ---- complex general path --------------------------------------------------------
// GenEigsSolver<SparseGenMatProd<std::complex<double>, Flags>> (GenEigsBase.h:111-140 complex restart, UpperHessenbergQR<complex>,
// Givens<complex>, UpperHessenbergEigen<complex> over the stand-in's ComplexSchur).  Everything interleaved (re, im).
namespace {
template <typename Eigs>
void run_gen_z(Eigs& eigs, int selection, int64_t maxit, double tol, int sorting, const double* init_resid, double* evals, double* evecs, RefResult* res)
{
    const double t0 = now_s();
    if (init_resid)
        eigs.init(reinterpret_cast<const Complex*>(init_resid));
    else
        eigs.init();
    const Index nconv = eigs.compute(Spectra::SortRule(selection), maxit, tol, Spectra::SortRule(sorting));
    res->seconds = now_s() - t0;
    res->nconv = nconv;
    res->niter = eigs.num_iterations();
    res->nops = eigs.num_operations();
    res->info = int32_t(eigs.info());
    auto ev = eigs.eigenvalues();
    for (Index i = 0; i < ev.size(); i++)
    {
        evals[2 * i] = ev[i].real();
        evals[2 * i + 1] = ev[i].imag();
    }
    if (evecs)
        cmat_to(eigs.eigenvectors(), evecs);
}
}  // namespace

extern "C" {

int ref_gen_eigs_complex(int order, const Compressed* A, int64_t nev, int64_t ncv, int selection, int64_t maxit, double tol, int sorting,
                         const double* init_resid, double* evals, double* evecs, RefResult* res)
{
    REF_TRY
    if (order == 0)
    {
        Eigen::Map<const Eigen::SparseMatrix<Complex, Eigen::ColMajor, int>> mat(A->n, A->n, A->nnz, A->outer, A->inner, reinterpret_cast<const Complex*>(A->val));
        using Op = Spectra::SparseGenMatProd<Complex, Eigen::ColMajor>;
        Op op(mat);
        Spectra::GenEigsSolver<Op> eigs(op, nev, ncv);
        run_gen_z(eigs, selection, maxit, tol, sorting, init_resid, evals, evecs, res);
    }
    else
    {
        Eigen::Map<const Eigen::SparseMatrix<Complex, Eigen::RowMajor, int>> mat(A->n, A->n, A->nnz, A->outer, A->inner, reinterpret_cast<const Complex*>(A->val));
        using Op = Spectra::SparseGenMatProd<Complex, Eigen::RowMajor>;
        Op op(mat);
        Spectra::GenEigsSolver<Op> eigs(op, nev, ncv);
        run_gen_z(eigs, selection, maxit, tol, sorting, init_resid, evals, evecs, res);
    }
    REF_CATCH
}

int ref_gen_eigs_complex_userop(int64_t n, void (*fn)(const double*, double*, void*), void* user, int64_t nev, int64_t ncv, int selection, int64_t maxit,
                                double tol, int sorting, const double* init_resid, double* evals, double* evecs, RefResult* res)
{
    REF_TRY
    CallbackOpZ op{Index(n), fn, user};
    Spectra::GenEigsSolver<CallbackOpZ> eigs(op, nev, ncv);
    run_gen_z(eigs, selection, maxit, tol, sorting, init_resid, evals, evecs, res);
    REF_CATCH
}

// UpperHessenbergEigen<std::complex<double>> (UpperHessenbergEigen.h:328-454): H, evals, evecs interleaved, column-major
int ref_hess_eigen_complex(int64_t m, const double* H_ri, double* evals, double* evecs)
{
    REF_TRY
    Eigen::Map<const CMatrix> h(reinterpret_cast<const Complex*>(H_ri), m, m);
    Spectra::UpperHessenbergEigen<Complex> dec(h);
    const auto& ev = dec.eigenvalues();
    for (int64_t i = 0; i < m; i++)
    {
        evals[2 * i] = ev[i].real();
        evals[2 * i + 1] = ev[i].imag();
    }
    cmat_to(dec.eigenvectors(), evecs);
    REF_CATCH
}

// UpperHessenbergQR<std::complex<double>> with a complex shift (:136-255, :383-417): R, Q^H H Q, Q = I G_1 G_2 ...
int ref_shifted_qr_complex(int64_t m, const double* H_ri, double shift_re, double shift_im, double* R, double* QtHQ, double* Q)
{
    REF_TRY
    Eigen::Map<const CMatrix> h(reinterpret_cast<const Complex*>(H_ri), m, m);
    CMatrix q = CMatrix::Identity(m, m), d;
    Spectra::UpperHessenbergQR<Complex> dec(m);
    dec.compute(h, Complex(shift_re, shift_im));
    cmat_to(dec.matrix_R(), R);
    dec.matrix_QtHQ(d);
    dec.apply_YQ(q);
    cmat_to(d, QtHQ);
    cmat_to(q, Q);
    REF_CATCH
}

// Givens<std::complex<double>>::compute_rotation (Givens.h:218-335)
void ref_givens_complex(double xr, double xi, double yr, double yi, double* r_ri, double* c, double* s_ri)
{
    Complex r, s;
    double cc;
    Spectra::Givens<Complex>::compute_rotation(Complex(xr, xi), Complex(yr, yi), r, cc, s);
    r_ri[0] = r.real();
    r_ri[1] = r.imag();
    *c = cc;
    s_ri[0] = s.real();
    s_ri[1] = s.imag();
}

}  // extern "C"

// ---- shift-and-invert (BASELINE config 5) ------------------------------------------
// SymEigsShiftSolver<SparseSymShiftSolve<double, Uplo, ColMajor>> (SymEigsShiftSolver.h:148-196, MatOp/SparseSymShiftSolve.h:30-110).  The
// reference's set_shift() builds A - sigma I from the stored triangle and hands it to Eigen::SparseLU -- here the stand-in's band LU with
// partial pivoting -- and sort_ritzpair() maps nu back to lambda = 1 / nu + sigma.
extern "C" {

int ref_shift_solve(int uplo, const Compressed* A, double sigma, const double* x, double* y)
{
    REF_TRY
    SpMap<Eigen::ColMajor> mat(A->n, A->n, A->nnz, A->outer, A->inner, A->val);
    if (uplo == 0)
    {
        Spectra::SparseSymShiftSolve<double, Eigen::Lower> op(mat);
        op.set_shift(sigma);
        op.perform_op(x, y);
    }
    else
    {
        Spectra::SparseSymShiftSolve<double, Eigen::Upper> op(mat);
        op.set_shift(sigma);
        op.perform_op(x, y);
    }
    REF_CATCH
}

int ref_sym_shift_eigs(int uplo, const Compressed* A, double sigma, int64_t nev, int64_t ncv, int selection, int64_t maxit, double tol, int sorting,
                       const double* init_resid, double* evals, double* evecs, RefResult* res)
{
    REF_TRY
    SpMap<Eigen::ColMajor> mat(A->n, A->n, A->nnz, A->outer, A->inner, A->val);
    auto run = [&](auto& op) {
        using Op = typename std::remove_reference<decltype(op)>::type;
        Spectra::SymEigsShiftSolver<Op> eigs(op, nev, ncv, sigma);
        const double t0 = now_s();
        if (init_resid)
            eigs.init(init_resid);
        else
            eigs.init();
        const Index nconv = eigs.compute(Spectra::SortRule(selection), maxit, tol, Spectra::SortRule(sorting));
        res->seconds = now_s() - t0;
        res->nconv = nconv;
        res->niter = eigs.num_iterations();
        res->nops = eigs.num_operations();
        res->info = int32_t(eigs.info());
        Vector ev = eigs.eigenvalues();
        if (evals)
            std::memcpy(evals, ev.data(), sizeof(double) * size_t(ev.size()));
        if (evecs)
            mat_to(eigs.eigenvectors(), evecs);
    };
    if (uplo == 0)
    {
        Spectra::SparseSymShiftSolve<double, Eigen::Lower> op(mat);
        run(op);
    }
    else
    {
        Spectra::SparseSymShiftSolve<double, Eigen::Upper> op(mat);
        run(op);
    }
    REF_CATCH
}

}  // extern "C"

// ORACLE — TEST INFRASTRUCTURE ONLY.  C ABI over the CPU restatement (dense.hpp, solver.hpp) so
// that tests/ and bench.py's CPU-baseline legs can drive it through ctypes.  Nothing in the
// shipped product (spectra_b200/, include/) links, loads or calls this library.
#include <chrono>
#include <cstring>
#include <random>

#include "band.hpp"
#include "solver.hpp"

using namespace oracle;

namespace {
thread_local std::string g_err;
int fail(const std::exception& e, int code)
{
    g_err = e.what();
    return code;
}
#define ORACLE_TRY try {
#define ORACLE_CATCH                                      \
    }                                                     \
    catch (const std::invalid_argument& e) { return fail(e, 1); } \
    catch (const std::logic_error& e) { return fail(e, 2); }      \
    catch (const std::runtime_error& e) { return fail(e, 3); }    \
    catch (const std::exception& e) { return fail(e, 4); }        \
    return 0;

Mat mat_from(const double* p, Index r, Index c)
{
    Mat m(r, c);
    std::memcpy(m.data(), p, sizeof(double) * size_t(r * c));
    return m;
}
void mat_to(const Mat& m, double* p)
{
    if (p)
        std::memcpy(p, m.data(), sizeof(double) * size_t(m.r * m.c));
}
double now_s()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct UserOp
{
    Index n;
    void (*fn)(const double*, double*, void*);
    void* user;
    Index rows() const { return n; }
    void perform_op(const double* x, double* y) const { fn(x, y, user); }
};
}  // namespace

extern "C" {

const char* oracle_last_error() { return g_err.c_str(); }

int oracle_max_threads()
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// ---- operator -------------------------------------------------------------------------------
// order: 0 ColMajor / 1 RowMajor; mode: 0 general, 1 sym-Lower, 2 sym-Upper
int oracle_csr_create(int64_t n, const int64_t* outer, const int32_t* inner, const double* val, int order, int mode, void** out)
{
    ORACLE_TRY
    *out = new CsrOp(build_full_csr(n, outer, inner, val, order, mode));
    ORACLE_CATCH
}
// adopt an existing full CSR (rows ascending, no expansion) — used for the large synthetic inputs
int oracle_csr_adopt(int64_t n, const int64_t* rowptr, const int32_t* col, const double* val, void** out)
{
    ORACLE_TRY
    auto* op = new CsrOp();
    op->n = n;
    op->rowptr.assign(rowptr, rowptr + n + 1);
    op->col.assign(col, col + rowptr[n]);
    op->val.assign(val, val + rowptr[n]);
    *out = op;
    ORACLE_CATCH
}
void oracle_csr_free(void* h) { delete static_cast<CsrOp*>(h); }
int64_t oracle_csr_nnz(void* h) { return static_cast<CsrOp*>(h)->rowptr.back(); }
void oracle_csr_export(void* h, int64_t* rowptr, int32_t* col, double* val)
{
    auto* op = static_cast<CsrOp*>(h);
    std::memcpy(rowptr, op->rowptr.data(), sizeof(int64_t) * op->rowptr.size());
    std::memcpy(col, op->col.data(), sizeof(int32_t) * op->col.size());
    std::memcpy(val, op->val.data(), sizeof(double) * op->val.size());
}
void oracle_csr_set_threads(void* h, int threads) { static_cast<CsrOp*>(h)->threads = threads; }
void oracle_spmv(void* h, const double* x, double* y) { static_cast<CsrOp*>(h)->perform_op(x, y); }

// ---- fixtures -------------------------------------------------------------------------------
// SimpleRandom(seed).random_vec(n)   Util/SimpleRandom.h:80-123
void oracle_simple_random(uint64_t seed, int64_t n, double* out)
{
    SimpleRandom rng(seed);
    rng.random_vec(out, n);
}

// gen_sparse_data(n, prob) of the reference tests (test/SymEigs.cpp:25-42, test/GenEigs.cpp:21-36):
// std::default_random_engine seeded 0 + uniform_real_distribution<double>(0,1), row-major visiting
// order.  Returns the number of entries; arrays must hold n*n entries.
int64_t oracle_gen_sparse_data(int64_t n, double prob, int32_t* rows, int32_t* cols, double* vals)
{
    std::default_random_engine gen;
    gen.seed(0);
    std::uniform_real_distribution<double> distr(0.0, 1.0);
    int64_t cnt = 0;
    for (int64_t i = 0; i < n; i++)
        for (int64_t j = 0; j < n; j++)
            if (distr(gen) < prob)
            {
                rows[cnt] = int32_t(i);
                cols[cnt] = int32_t(j);
                vals[cnt] = distr(gen) - 0.5;
                cnt++;
            }
    return cnt;
}

// gen_sparse_data(n, prob) of test/HermEigs.cpp:27-50: same engine and visiting order; an accepted position draws the real part and,
// off the diagonal, the imaginary part.  vals_ri holds interleaved (re, im) pairs; arrays must hold n*n entries.
int64_t oracle_gen_sparse_data_herm(int64_t n, double prob, int32_t* rows, int32_t* cols, double* vals_ri)
{
    std::default_random_engine gen;
    gen.seed(0);
    std::uniform_real_distribution<double> distr(0.0, 1.0);
    int64_t cnt = 0;
    for (int64_t i = 0; i < n; i++)
        for (int64_t j = 0; j < n; j++)
            if (distr(gen) < prob)
            {
                const double re = distr(gen) - 0.5;
                const double im = (i == j) ? 0.0 : (distr(gen) - 0.5);
                rows[cnt] = int32_t(i);
                cols[cnt] = int32_t(j);
                vals_ri[2 * cnt] = re;
                vals_ri[2 * cnt + 1] = im;
                cnt++;
            }
    return cnt;
}

// gen_sparse_data(n, prob) of test/ComplexEigs.cpp:20-39: same engine and visiting order; every accepted position draws re and im
int64_t oracle_gen_sparse_data_complex(int64_t n, double prob, int32_t* rows, int32_t* cols, double* vals_ri)
{
    std::default_random_engine gen;
    gen.seed(0);
    std::uniform_real_distribution<double> distr(0.0, 1.0);
    int64_t cnt = 0;
    for (int64_t i = 0; i < n; i++)
        for (int64_t j = 0; j < n; j++)
            if (distr(gen) < prob)
            {
                rows[cnt] = int32_t(i);
                cols[cnt] = int32_t(j);
                vals_ri[2 * cnt] = distr(gen) - 0.5;
                vals_ri[2 * cnt + 1] = distr(gen) - 0.5;
                cnt++;
            }
    return cnt;
}

// ---- small dense kernels --------------------------------------------------------------------
void oracle_givens(double x, double y, double* r, double* c, double* s) { givens_rotation(x, y, *r, *c, *s); }

// kind: 0 TridiagQR, 1 UpperHessenbergQR.  Q = I * G1 * G2 ... (apply_YQ on identity).
int oracle_shifted_qr(int kind, int64_t m, const double* H, double shift, double* R, double* QtHQ, double* Q)
{
    ORACLE_TRY
    Mat h = mat_from(H, m, m);
    Mat q(m, m), d;
    q.set_identity();
    if (kind == 0)
    {
        TridiagQR dec(m);
        dec.compute(h, shift);
        mat_to(dec.matrix_R(), R);
        dec.matrix_QtHQ(d);
        dec.apply_YQ(q);
    }
    else
    {
        UpperHessenbergQR dec(m);
        dec.compute(h, shift);
        mat_to(dec.matrix_R(), R);
        dec.matrix_QtHQ(d);
        dec.apply_YQ(q);
    }
    mat_to(d, QtHQ);
    mat_to(q, Q);
    ORACLE_CATCH
}

int oracle_double_shift_qr(int64_t m, const double* H, double s, double t, double* QtHQ, double* Q)
{
    ORACLE_TRY
    Mat h = mat_from(H, m, m);
    DoubleShiftQR dec(m);
    dec.compute(h, s, t);
    Mat d, q(m, m);
    q.set_identity();
    dec.matrix_QtHQ(d);
    dec.apply_YQ(q);
    mat_to(d, QtHQ);
    mat_to(q, Q);
    ORACLE_CATCH
}

int oracle_tridiag_eigen(int64_t m, const double* H, double* evals, double* evecs)
{
    ORACLE_TRY
    TridiagEigen dec(mat_from(H, m, m));
    std::memcpy(evals, dec.eigenvalues().data(), sizeof(double) * m);
    mat_to(dec.eigenvectors(), evecs);
    ORACLE_CATCH
}

int oracle_hess_schur(int64_t m, const double* H, double* T, double* U)
{
    ORACLE_TRY
    UpperHessenbergSchur dec;
    dec.compute(mat_from(H, m, m));
    mat_to(dec.matrix_T(), T);
    mat_to(dec.matrix_U(), U);
    ORACLE_CATCH
}

// evals/evecs as interleaved (re, im) pairs, evecs column-major m x m
int oracle_hess_eigen(int64_t m, const double* H, double* evals, double* evecs)
{
    ORACLE_TRY
    UpperHessenbergEigen dec(mat_from(H, m, m));
    const auto& ev = dec.eigenvalues();
    for (int64_t i = 0; i < m; i++)
    {
        evals[2 * i] = ev[i].real();
        evals[2 * i + 1] = ev[i].imag();
    }
    CMat V = dec.eigenvectors();
    for (int64_t i = 0; i < m * m; i++)
    {
        evecs[2 * i] = V.a[i].real();
        evecs[2 * i + 1] = V.a[i].imag();
    }
    ORACLE_CATCH
}

// argsort(selection, values, len)  SelectionRule.h:227-287
int oracle_argsort(int selection, const double* values, int64_t len, int64_t* ind)
{
    ORACLE_TRY
    auto v = argsort(SortRule(selection), values, len);
    for (int64_t i = 0; i < len; i++)
        ind[i] = v[i];
    ORACLE_CATCH
}
int oracle_argsort_complex(int selection, const double* values_ri, int64_t len, int64_t* ind)
{
    ORACLE_TRY
    std::vector<Complex> c(len);
    for (int64_t i = 0; i < len; i++)
        c[i] = Complex(values_ri[2 * i], values_ri[2 * i + 1]);
    auto v = sort_eigenvalue(SortRule(selection), c.data(), len);
    for (int64_t i = 0; i < len; i++)
        ind[i] = v[i];
    ORACLE_CATCH
}

// ---- factorisation-level hooks (test/Arnoldi.cpp style) ---------------------------------------
// kind 0: Lanczos, 1: Arnoldi.  init(v0 or SimpleRandom(0)) then factorize_from(1, mid) and
// factorize_from(mid, m).  Outputs V (n x m), H (m x m), f (n), beta.
int oracle_factorize(int kind, void* csr, int64_t m, const double* v0, int64_t mid, double* V, double* H, double* f, double* beta, int64_t* nops,
                     int64_t* stats4)
{
    ORACLE_TRY
    auto* op = static_cast<CsrOp*>(csr);
    const Index n = op->n;
    std::vector<double> init(n);
    if (v0)
        std::memcpy(init.data(), v0, sizeof(double) * n);
    else
    {
        SimpleRandom rng(0);
        rng.random_vec(init.data(), n);
    }
    Index cnt = 0;
    auto run = [&](Arnoldi<CsrOp>& fac) {
        fac.init(init.data(), cnt);
        fac.factorize_from(1, mid, cnt);
        fac.factorize_from(mid, m, cnt);
        mat_to(fac.matrix_V(), V);
        mat_to(fac.matrix_H(), H);
        std::memcpy(f, fac.vector_f().data(), sizeof(double) * n);
        *beta = fac.f_norm();
        if (stats4)
        {
            stats4[0] = fac.stats.reorth_passes;
            stats4[1] = fac.stats.expand_calls;
            stats4[2] = fac.stats.restarts;
            stats4[3] = fac.stats.lanczos_steps;
        }
    };
    if (kind == 0)
    {
        Lanczos<CsrOp> fac(*op, m, op->threads);
        run(fac);
    }
    else
    {
        Arnoldi<CsrOp> fac(*op, m, op->threads);
        run(fac);
    }
    *nops = cnt;
    ORACLE_CATCH
}

// One restart "prepare" step of HermEigsBase (retrieve_ritzpair :205-224, num_converged :158-175,
// nev_adjusted :178-202, shift loop of restart :105-147) on a given tridiagonal H and beta.
int oracle_sym_restart_prepare(int64_t m, const double* H, double beta, int64_t nev, int selection, double tol, double* ritz_val, double* ritz_est,
                               double* ritz_vec, int32_t* conv, int64_t* nconv_out, int64_t* k_out, double* Q_out, double* Hnew_out)
{
    ORACLE_TRY
    Mat h = mat_from(H, m, m);
    TridiagEigen dec(h);
    const auto& evals = dec.eigenvalues();
    const Mat& evecs = dec.eigenvectors();
    auto ind = argsort(SortRule(selection), evals.data(), m);
    std::vector<double> rv(m), re(m);
    for (Index i = 0; i < m; i++)
    {
        rv[i] = evals[ind[i]];
        re[i] = evecs(m - 1, ind[i]);
        ritz_val[i] = rv[i];
        ritz_est[i] = re[i];
    }
    for (Index i = 0; i < nev; i++)
        std::memcpy(ritz_vec + i * m, evecs.col(ind[i]), sizeof(double) * m);
    const double eps23 = std::pow(kEps, 2.0 / 3.0);
    Index nconv = 0;
    for (Index i = 0; i < nev; i++)
    {
        const double thresh = tol * std::max(std::abs(rv[i]), eps23);
        const double resid = std::abs(re[i]) * beta;
        conv[i] = (resid < thresh) ? 1 : 0;
        nconv += conv[i];
    }
    *nconv_out = nconv;
    Index nev_new = nev;
    for (Index i = nev; i < m; i++)
        if (std::abs(re[i]) < kNear0)
            nev_new++;
    nev_new += std::min(nconv, (m - nev_new) / 2);
    if (nev_new == 1 && m >= 6)
        nev_new = m / 2;
    else if (nev_new == 1 && m > 2)
        nev_new = 2;
    if (nev_new > m - 1)
        nev_new = m - 1;
    *k_out = nev_new;

    const Index k = nev_new;
    TridiagQR decomp(m);
    Mat Q(m, m);
    Q.set_identity();
    const Index nshift = m - k;
    std::vector<double> shifts(rv.end() - nshift, rv.end());
    std::sort(shifts.begin(), shifts.end(), [](const double& a, const double& b) { return std::abs(a) > std::abs(b); });
    for (Index i = 0; i < nshift; i++)
    {
        decomp.compute(h, shifts[i]);
        decomp.apply_YQ(Q);
        decomp.matrix_QtHQ(h);
    }
    mat_to(Q, Q_out);
    mat_to(h, Hnew_out);
    ORACLE_CATCH
}

}  // extern "C"

// ---- full solvers ---------------------------------------------------------------------------
struct OracleResult
{
    int64_t nconv, niter, nops;
    int32_t info;
    int64_t reorth_passes, expand_calls, restarts, steps;
    double seconds;
};

template <typename OpT>
static void run_sym(const OpT& op, int64_t nev, int64_t ncv, int selection, int64_t maxit, double tol, int sorting, const double* init_resid, int shift_mode,
                    double sigma, int threads, int64_t op_limit, double* evals, double* evecs, OracleResult* res)
{
    SymEigsSolver<OpT> eigs(op, nev, ncv, threads);
    if (shift_mode)
        eigs.set_shift_mode(sigma);
    if (op_limit >= 0)
        eigs.set_op_limit(op_limit);
    const double t0 = now_s();
    if (init_resid)
        eigs.init(init_resid);
    else
        eigs.init();
    const Index nconv = eigs.compute(SortRule(selection), maxit, tol, SortRule(sorting));
    res->seconds = now_s() - t0;
    res->nconv = nconv;
    res->niter = eigs.num_iterations();
    res->nops = eigs.num_operations();
    res->info = int32_t(eigs.info());
    res->reorth_passes = eigs.stats().reorth_passes;
    res->expand_calls = eigs.stats().expand_calls;
    res->restarts = eigs.stats().restarts;
    res->steps = eigs.stats().lanczos_steps;
    if (op_limit >= 0)
        return;
    auto ev = eigs.eigenvalues();
    if (evals)
        std::memcpy(evals, ev.data(), sizeof(double) * ev.size());
    if (evecs)
        mat_to(eigs.eigenvectors(), evecs);
}

extern "C" {

int oracle_sym_eigs(void* csr, int64_t nev, int64_t ncv, int selection, int64_t maxit, double tol, int sorting, const double* init_resid, int shift_mode,
                    double sigma, int threads, int64_t op_limit, double* evals, double* evecs, OracleResult* res)
{
    ORACLE_TRY
    auto* op = static_cast<CsrOp*>(csr);
    op->threads = threads;
    run_sym(*op, nev, ncv, selection, maxit, tol, sorting, init_resid, shift_mode, sigma, threads, op_limit, evals, evecs, res);
    ORACLE_CATCH
}

// user-defined operator (the OpType concept): y = fn(x)
int oracle_sym_eigs_userop(int64_t n, void (*fn)(const double*, double*, void*), void* user, int64_t nev, int64_t ncv, int selection, int64_t maxit, double tol,
                           int sorting, const double* init_resid, int shift_mode, double sigma, double* evals, double* evecs, OracleResult* res)
{
    ORACLE_TRY
    UserOp op{n, fn, user};
    run_sym(op, nev, ncv, selection, maxit, tol, sorting, init_resid, shift_mode, sigma, 1, -1, evals, evecs, res);
    ORACLE_CATCH
}

// ---- shift-invert: SparseSymShiftSolve (band LU restatement, band.hpp) + SymEigsShiftSolver (SymEigsShiftSolver.h:148-196) ----
int oracle_band_create(void* csr, double sigma, void** out)
{
    ORACLE_TRY
    auto* A = static_cast<CsrOp*>(csr);
    auto* op = new BandLuOp();
    try
    {
        op->set_shift(*A, sigma);
    }
    catch (...)
    {
        delete op;
        throw;
    }
    *out = op;
    ORACLE_CATCH
}
int oracle_band_info(void* h, int64_t* kl, int64_t* ku)
{
    auto* op = static_cast<BandLuOp*>(h);
    *kl = op->kl;
    *ku = op->ku;
    return 0;
}
int oracle_band_solve(void* h, const double* x, double* y)
{
    ORACLE_TRY
    static_cast<BandLuOp*>(h)->perform_op(x, y);
    ORACLE_CATCH
}
void oracle_band_destroy(void* h) { delete static_cast<BandLuOp*>(h); }

int oracle_sym_shift_eigs(void* band, double sigma, int64_t nev, int64_t ncv, int selection, int64_t maxit, double tol, int sorting, const double* init_resid,
                          int64_t op_limit, double* evals, double* evecs, OracleResult* res)
{
    ORACLE_TRY
    auto* op = static_cast<BandLuOp*>(band);
    run_sym(*op, nev, ncv, selection, maxit, tol, sorting, init_resid, 1, sigma, 1, op_limit, evals, evecs, res);
    ORACLE_CATCH
}

// evals / evecs interleaved (re, im)
int oracle_gen_eigs(void* csr, int64_t nev, int64_t ncv, int selection, int64_t maxit, double tol, int sorting, const double* init_resid, int threads,
                    int64_t op_limit, double* evals, double* evecs, OracleResult* res)
{
    ORACLE_TRY
    auto* op = static_cast<CsrOp*>(csr);
    op->threads = threads;
    GenEigsSolver<CsrOp> eigs(*op, nev, ncv, threads);
    if (op_limit >= 0)
        eigs.set_op_limit(op_limit);
    const double t0 = now_s();
    if (init_resid)
        eigs.init(init_resid);
    else
        eigs.init();
    const Index nconv = eigs.compute(SortRule(selection), maxit, tol, SortRule(sorting));
    res->seconds = now_s() - t0;
    res->nconv = nconv;
    res->niter = eigs.num_iterations();
    res->nops = eigs.num_operations();
    res->info = int32_t(eigs.info());
    res->reorth_passes = eigs.stats().reorth_passes;
    res->expand_calls = eigs.stats().expand_calls;
    res->restarts = eigs.stats().restarts;
    res->steps = eigs.stats().lanczos_steps;
    if (op_limit >= 0)
        return 0;
    auto ev = eigs.eigenvalues();
    for (size_t i = 0; i < ev.size(); i++)
    {
        evals[2 * i] = ev[i].real();
        evals[2 * i + 1] = ev[i].imag();
    }
    if (evecs)
    {
        CMat V = eigs.eigenvectors();
        for (size_t i = 0; i < V.a.size(); i++)
        {
            evecs[2 * i] = V.a[i].real();
            evecs[2 * i + 1] = V.a[i].imag();
        }
    }
    ORACLE_CATCH
}

}  // extern "C"

// extern "C" boundary (include/spectra_b200.h).  Converts C++ exceptions to status codes.
#include <cstring>

#include "host.h"

struct sb200_sym_solver;
struct sb200_gen_solver;

namespace sb200 {
const char* last_error_cstr();
sb200_op* op_create_sparse(int64_t n, const void* outer, int outer_is_64, const int32_t* inner, const double* values, int storage_order, int matrix_mode,
                           sb200_comm* comm);
sb200_op* op_create_csr_slab(int64_t n, int64_t row0, int64_t nrows, const int64_t* rowptr_local, const int32_t* col, const double* values, sb200_comm* comm);
void op_spmv_device(sb200_op* op, const double* x_dev, double* y_dev);
sb200_op* op_create_callback(int64_t n, void (*fn)(const double*, double*, void*), void* user, bool is_complex);
sb200_op* op_create_shift_solve(int64_t n, const void* outer, int outer_is_64, const int32_t* inner, const double* values, int storage_order, int matrix_mode);
void op_perform_op_host(sb200_op* op, const double* x_host, double* y_host);
sb200_op* op_create_sparse_herm(int64_t n, const void* outer, int outer_is_64, const int32_t* inner, const double* values_ri, int storage_order, int matrix_mode);
float bench_gather(int64_t n, int64_t gathers, int repeat, double* checksum);
float bench_stream_gather(int64_t n, int64_t gathers, int band, int repeat, double* checksum);

sb200_sym_solver* sym_create(sb200_op* op, int64_t nev, int64_t ncv, bool shift_mode, double sigma);
void sym_init(sb200_sym_solver* s, const double* resid);
int64_t sym_compute(sb200_sym_solver* s, int selection, int64_t maxit, double tol, int sorting);
void sym_factorize_from(sb200_sym_solver* s, int64_t from_k, int64_t to_m);
void sym_get_factorization(sb200_sym_solver* s, double* Vh, double* Hh, double* fh, double* beta, int64_t* kk);
int64_t sym_eigenvalues(const sb200_sym_solver* s, double* out);
int64_t sym_eigenvectors_local(sb200_sym_solver* s, int64_t nvec, double* out);
int64_t sym_eigenvectors_full(sb200_sym_solver* s, int64_t nvec, double* out);
int sym_info(const sb200_sym_solver* s);
int64_t sym_niter(const sb200_sym_solver* s);
int64_t sym_nops(const sb200_sym_solver* s);
const sb200_stats& sym_stats(const sb200_sym_solver* s);
void sym_destroy(sb200_sym_solver* s);
void dense_sym_restart_host(int64_t m, const double* H, double beta, int64_t nev, int selection, double tol, double* ritz_val, double* ritz_est, int32_t* conv,
                            int64_t* nconv, int64_t* k, double* Q, double* Hnew);
void launch_tridiag_qr(const double* H, int m, double shift, double* QtHQ, double* Q, cudaStream_t stream);

// solver_gen.cu
sb200_gen_solver* gen_create(sb200_op* op, int64_t nev, int64_t ncv);
void gen_init(sb200_gen_solver* s, const double* resid);
int64_t gen_compute(sb200_gen_solver* s, int selection, int64_t maxit, double tol, int sorting);
void gen_factorize_from(sb200_gen_solver* s, int64_t from_k, int64_t to_m);
void gen_get_factorization(sb200_gen_solver* s, double* Vh, double* Hh, double* fh, double* beta, int64_t* kk);
int64_t gen_eigenvalues(const sb200_gen_solver* s, double* out_ri);
int64_t gen_eigenvectors(sb200_gen_solver* s, int64_t nvec, double* out_ri);
int gen_info(const sb200_gen_solver* s);
int64_t gen_niter(const sb200_gen_solver* s);
int64_t gen_nops(const sb200_gen_solver* s);
const sb200_stats& gen_stats(const sb200_gen_solver* s);
void gen_destroy(sb200_gen_solver* s);
void dense_hess_qr_host(int64_t m, const double* H, double shift, double* QtHQ, double* Q);
void dense_double_shift_qr_host(int64_t m, const double* H, double s, double t, double* QtHQ, double* Q);
void dense_hess_eigen_host(int64_t m, const double* H, double* evals_ri, double* evecs_ri);
// dense_gen_z.cu
void dense_hess_qr_z_host(int64_t m, const double* H_ri, double mu_re, double mu_im, double* QtHQ_ri, double* Q_ri);
void dense_hess_eigen_z_host(int64_t m, const double* H_ri, double* evals_ri, double* evecs_ri);
}  // namespace sb200

using namespace sb200;

#define ABI_TRY try {
#define ABI_CATCH                                  \
    }                                              \
    catch (const sb200::Error& e)                  \
    {                                              \
        sb200::set_last_error(e.what());           \
        return e.status;                           \
    }                                              \
    catch (const std::bad_alloc&)                  \
    {                                              \
        sb200::set_last_error("host out of memory"); \
        return SB200_RUNTIME;                      \
    }                                              \
    catch (const std::invalid_argument& e)         \
    {                                              \
        /* thrown by a user-defined C++ OpType inside the callback trampoline: keep its type across the C boundary */ \
        sb200::set_last_error(e.what());           \
        return SB200_INVALID_ARGUMENT;             \
    }                                              \
    catch (const std::logic_error& e)              \
    {                                              \
        sb200::set_last_error(e.what());           \
        return SB200_LOGIC;                        \
    }                                              \
    catch (const std::exception& e)                \
    {                                              \
        sb200::set_last_error(e.what());           \
        return SB200_RUNTIME;                      \
    }                                              \
    return SB200_OK;

#define ABI_NONNULL(p) SB200_REQUIRE((p) != nullptr, SB200_INVALID_ARGUMENT, "null pointer argument: " #p)

#pragma GCC visibility push(default)
extern "C" {

const char* sb200_last_error(void) { return sb200::last_error_cstr(); }
const char* sb200_version(void) { return "spectra_b200 0.1 (sm_100a)"; }

int sb200_device_info(int* device, int* sm_count, int* cc_major, int* cc_minor, int64_t* hbm_bytes)
{
    ABI_TRY
    const DeviceInfo& d = device_info();
    if (device)
        *device = d.device;
    if (sm_count)
        *sm_count = d.sm_count;
    if (cc_major)
        *cc_major = d.cc_major;
    if (cc_minor)
        *cc_minor = d.cc_minor;
    if (hbm_bytes)
        *hbm_bytes = (int64_t) d.total_mem;
    ABI_CATCH
}

int sb200_set_device(int device)
{
    ABI_TRY
    SB200_CUDA_CHECK(cudaSetDevice(device));
    ABI_CATCH
}

int sb200_set_profiling(int level)
{
    g_profiling_level = level;
    return SB200_OK;
}

// ---- communicator ----
int sb200_comm_unique_id(void* id128)
{
    ABI_TRY
    ABI_NONNULL(id128);
    nccl_unique_id(id128);
    ABI_CATCH
}
int sb200_comm_create(int rank, int nranks, const void* id128, sb200_comm** out)
{
    ABI_TRY
    ABI_NONNULL(out);
    SB200_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, SB200_INVALID_ARGUMENT, "bad rank / nranks");
    device_info();
    std::unique_ptr<sb200_comm> c(new sb200_comm());
    c->rank = rank;
    c->nranks = nranks;
    if (nranks > 1)
    {
        ABI_NONNULL(id128);
        nccl_comm_init(c.get(), id128);
    }
    *out = c.release();
    ABI_CATCH
}
int sb200_comm_rank(const sb200_comm* c, int* rank, int* nranks)
{
    ABI_TRY
    ABI_NONNULL(c);
    if (rank)
        *rank = c->rank;
    if (nranks)
        *nranks = c->nranks;
    ABI_CATCH
}
int sb200_comm_destroy(sb200_comm* c)
{
    ABI_TRY
    if (c)
    {
        nccl_comm_destroy(c);
        delete c;
    }
    ABI_CATCH
}

// ---- operator ----
int sb200_op_create_sparse(int64_t n, const void* outer, int outer_is_64, const int32_t* inner, const double* values, int storage_order, int matrix_mode,
                           sb200_comm* comm, sb200_op** out)
{
    ABI_TRY
    ABI_NONNULL(out);
    ABI_NONNULL(outer);
    *out = op_create_sparse(n, outer, outer_is_64, inner, values, storage_order, matrix_mode, (comm && comm->nranks > 1) ? comm : nullptr);
    ABI_CATCH
}
int sb200_op_create_sparse_herm(int64_t n, const void* outer, int outer_is_64, const int32_t* inner, const double* values_ri, int storage_order, int matrix_mode,
                                sb200_op** out)
{
    ABI_TRY
    ABI_NONNULL(out);
    ABI_NONNULL(outer);
    *out = op_create_sparse_herm(n, outer, outer_is_64, inner, values_ri, storage_order, matrix_mode);
    ABI_CATCH
}
int sb200_op_create_csr_slab(int64_t n, int64_t row0, int64_t nrows, const int64_t* rowptr_local, const int32_t* col, const double* values, sb200_comm* comm,
                             sb200_op** out)
{
    ABI_TRY
    ABI_NONNULL(out);
    ABI_NONNULL(rowptr_local);
    *out = op_create_csr_slab(n, row0, nrows, rowptr_local, col, values, (comm && comm->nranks > 1) ? comm : nullptr);
    ABI_CATCH
}
int sb200_op_create_callback(int64_t n, sb200_matvec_fn fn, void* user, sb200_op** out)
{
    ABI_TRY
    ABI_NONNULL(out);
    *out = op_create_callback(n, fn, user, false);
    ABI_CATCH
}
int sb200_op_create_callback_z(int64_t n, sb200_matvec_fn fn, void* user, sb200_op** out)
{
    ABI_TRY
    ABI_NONNULL(out);
    *out = op_create_callback(n, fn, user, true);
    ABI_CATCH
}
int sb200_op_create_shift_solve(int64_t n, const void* outer, int outer_is_64, const int32_t* inner, const double* values, int storage_order, int matrix_mode,
                                sb200_op** out)
{
    ABI_TRY
    ABI_NONNULL(out);
    ABI_NONNULL(outer);
    *out = op_create_shift_solve(n, outer, outer_is_64, inner, values, storage_order, matrix_mode);
    ABI_CATCH
}
int sb200_op_set_shift(sb200_op* op, double sigma)
{
    ABI_TRY
    ABI_NONNULL(op);
    band_set_shift(op, sigma);
    ABI_CATCH
}
int sb200_op_shift_solve_info(const sb200_op* op, int* half_bandwidth, int* block, int64_t* block_rows, int* levels)
{
    ABI_TRY
    ABI_NONNULL(op);
    band_info(op, half_bandwidth, block, block_rows, levels);
    ABI_CATCH
}
int sb200_op_shift_solve_refine(sb200_op* op, int steps)
{
    ABI_TRY
    ABI_NONNULL(op);
    band_set_refine(op, steps);
    ABI_CATCH
}
int sb200_op_shift_solve_status(const sb200_op* op, int* refine_steps, double* verify_residual, double* unrefined_residual)
{
    ABI_TRY
    ABI_NONNULL(op);
    band_status(op, refine_steps, verify_residual, unrefined_residual);
    ABI_CATCH
}
int sb200_op_rows(const sb200_op* op, int64_t* rows)
{
    ABI_TRY
    ABI_NONNULL(op);
    *rows = op->A.n;
    ABI_CATCH
}
int sb200_op_cols(const sb200_op* op, int64_t* cols)
{
    ABI_TRY
    ABI_NONNULL(op);
    *cols = op->A.n;
    ABI_CATCH
}
int sb200_op_local_rows(const sb200_op* op, int64_t* row0, int64_t* nrows)
{
    ABI_TRY
    ABI_NONNULL(op);
    if (row0)
        *row0 = op->A.row0;
    if (nrows)
        *nrows = op->A.nrows;
    ABI_CATCH
}
int sb200_op_nnz(const sb200_op* op, int64_t* nnz_local)
{
    ABI_TRY
    ABI_NONNULL(op);
    *nnz_local = op->A.nnz;
    ABI_CATCH
}
int sb200_op_layout_info(const sb200_op* op, int* format, int* col_blocks, int64_t* stored_entries)
{
    ABI_TRY
    ABI_NONNULL(op);
    const DeviceCsr& A = op->A;
    int64_t stored = 0;
    bool sell = op->plan.sell_threads > 0;
    if (A.blocks.empty())
        stored = sell ? A.sell.padded : A.nnz;
    else
        for (const CsrBlock& B : A.blocks)
            stored += sell ? B.sell.padded : B.nnz;
    if (format)
        *format = sell ? 1 : 0;
    if (col_blocks)
        *col_blocks = A.blocks.empty() ? 1 : (int) A.blocks.size();
    if (stored_entries)
        *stored_entries = stored;
    ABI_CATCH
}
int sb200_op_peer_mode(const sb200_op* op, int* peer)
{
    ABI_TRY
    ABI_NONNULL(op);
    ABI_NONNULL(peer);
    *peer = op->peer_mode() ? 1 : 0;
    ABI_CATCH
}
int sb200_bench_gather(int64_t n, int64_t gathers, int repeat, float* elapsed_ms, double* checksum)
{
    ABI_TRY
    device_info();
    const float ms = bench_gather(n, gathers, repeat, checksum);
    if (elapsed_ms)
        *elapsed_ms = ms;
    ABI_CATCH
}
int sb200_bench_stream_gather(int64_t n, int64_t gathers, int band, int repeat, float* elapsed_ms, double* checksum)
{
    ABI_TRY
    device_info();
    const float ms = bench_stream_gather(n, gathers, band, repeat, checksum);
    if (elapsed_ms)
        *elapsed_ms = ms;
    ABI_CATCH
}
int sb200_op_perform_op(sb200_op* op, const double* x_host, double* y_host)
{
    ABI_TRY
    ABI_NONNULL(op);
    ABI_NONNULL(x_host);
    ABI_NONNULL(y_host);
    op_perform_op_host(op, x_host, y_host);
    ABI_CATCH
}
int sb200_op_apply_matrix(sb200_op* op, const double* X_host, int64_t k, double* Y_host)
{
    ABI_TRY
    ABI_NONNULL(op);
    ABI_NONNULL(X_host);
    ABI_NONNULL(Y_host);
    const int cw = op->cplx ? 2 : 1;  // complex operators: interleaved (re, im) columns
    for (int64_t c = 0; c < k; c++)
        op_perform_op_host(op, X_host + c * op->A.n * cw, Y_host + c * op->A.nrows * cw);
    ABI_CATCH
}
int sb200_op_spmv_device(sb200_op* op, const double* x_dev, double* y_dev, int repeat, float* elapsed_ms)
{
    ABI_TRY
    ABI_NONNULL(op);
    if (repeat < 1)
        repeat = 1;
    DevBuf<double> xb, yb;
    if (!x_dev)
    {
        // benchmark convenience: x = ones, y = scratch
        std::vector<double> ones((size_t) op->A.n * (op->cplx ? 2 : 1), 1.0);  // complex operators: interleaved (re, im)
        xb.alloc(ones.size());
        SB200_CUDA_CHECK(cudaMemcpyAsync(xb.get(), ones.data(), sizeof(double) * ones.size(), cudaMemcpyHostToDevice, op->stream));
        SB200_CUDA_CHECK(cudaStreamSynchronize(op->stream));
        x_dev = xb.get();
    }
    if (!y_dev)
    {
        yb.alloc((size_t) std::max<int64_t>(op->A.nrows, 1) * (op->cplx ? 2 : 1));
        y_dev = yb.get();
    }
    SB200_CUDA_CHECK(cudaEventRecord(op->ev0, op->stream));
    for (int r = 0; r < repeat; r++)
        op_spmv_device(op, x_dev, y_dev);
    SB200_CUDA_CHECK(cudaEventRecord(op->ev1, op->stream));
    SB200_CUDA_CHECK(cudaEventSynchronize(op->ev1));
    if (elapsed_ms)
        SB200_CUDA_CHECK(cudaEventElapsedTime(elapsed_ms, op->ev0, op->ev1));
    ABI_CATCH
}
int sb200_op_destroy(sb200_op* op)
{
    ABI_TRY
    delete op;
    ABI_CATCH
}

// ---- SymEigsSolver ----
int sb200_sym_create(sb200_op* op, int64_t nev, int64_t ncv, sb200_sym_solver** out)
{
    ABI_TRY
    ABI_NONNULL(out);
    ABI_NONNULL(op);
    SB200_REQUIRE(!op->cplx, SB200_INVALID_ARGUMENT, "SymEigsSolver needs a real operator; use sb200_herm_create for a complex Hermitian one");
    *out = sym_create(op, nev, ncv, false, 0.0);
    ABI_CATCH
}
int sb200_herm_create(sb200_op* op, int64_t nev, int64_t ncv, sb200_sym_solver** out)
{
    ABI_TRY
    ABI_NONNULL(out);
    ABI_NONNULL(op);
    SB200_REQUIRE(op->cplx, SB200_INVALID_ARGUMENT, "HermEigsSolver needs a complex Hermitian operator (sb200_op_create_sparse_herm)");
    *out = sym_create(op, nev, ncv, false, 0.0);
    ABI_CATCH
}
int sb200_sym_create_shift(sb200_op* op, int64_t nev, int64_t ncv, double sigma, sb200_sym_solver** out)
{
    ABI_TRY
    ABI_NONNULL(out);
    ABI_NONNULL(op);
    // The reference makes a missing set_shift() a compile error (SymEigsShiftSolver.h:193); here the operator must be a shift-solve
    // operator (sb200_op_create_shift_solve) or a user callback that applies (A - sigma I)^{-1} itself, and it must be real.
    SB200_REQUIRE(!op->cplx, SB200_INVALID_ARGUMENT, "SymEigsShiftSolver needs a real operator");
    SB200_REQUIRE(op->band || op->cb, SB200_INVALID_ARGUMENT, "SymEigsShiftSolver needs a shift-solve operator (sb200_op_create_shift_solve) or a user-defined one");
    // Base(op, nev, ncv) argument checks first, then op.set_shift(sigma) (SymEigsShiftSolver.h:190-195)
    sb200_sym_solver* s = sym_create(op, nev, ncv, true, sigma);
    if (op->band)
    {
        try
        {
            band_set_shift(op, sigma);
        }
        catch (...)
        {
            sym_destroy(s);
            throw;
        }
    }
    *out = s;
    ABI_CATCH
}
int sb200_sym_init(sb200_sym_solver* s, const double* init_resid_or_null)
{
    ABI_TRY
    ABI_NONNULL(s);
    sym_init(s, init_resid_or_null);
    ABI_CATCH
}
int sb200_sym_compute(sb200_sym_solver* s, int selection, int64_t maxit, double tol, int sorting, int64_t* nconv)
{
    ABI_TRY
    ABI_NONNULL(s);
    const int64_t r = sym_compute(s, selection, maxit, tol, sorting);
    if (nconv)
        *nconv = r;
    ABI_CATCH
}
int sb200_sym_info(const sb200_sym_solver* s, int* info)
{
    ABI_TRY
    ABI_NONNULL(s);
    *info = sym_info(s);
    ABI_CATCH
}
int sb200_sym_num_iterations(const sb200_sym_solver* s, int64_t* niter)
{
    ABI_TRY
    ABI_NONNULL(s);
    *niter = sym_niter(s);
    ABI_CATCH
}
int sb200_sym_num_operations(const sb200_sym_solver* s, int64_t* nops)
{
    ABI_TRY
    ABI_NONNULL(s);
    *nops = sym_nops(s);
    ABI_CATCH
}
int sb200_sym_eigenvalues(const sb200_sym_solver* s, double* out, int64_t* count)
{
    ABI_TRY
    ABI_NONNULL(s);
    ABI_NONNULL(out);
    const int64_t c = sym_eigenvalues(s, out);
    if (count)
        *count = c;
    ABI_CATCH
}
int sb200_sym_eigenvectors(sb200_sym_solver* s, int64_t nvec, double* out, int64_t* ncols)
{
    ABI_TRY
    ABI_NONNULL(s);
    ABI_NONNULL(out);
    const int64_t c = sym_eigenvectors_full(s, nvec, out);
    if (ncols)
        *ncols = c;
    ABI_CATCH
}
int sb200_sym_eigenvectors_local(sb200_sym_solver* s, int64_t nvec, double* out, int64_t* ncols)
{
    ABI_TRY
    ABI_NONNULL(s);
    ABI_NONNULL(out);
    const int64_t c = sym_eigenvectors_local(s, nvec, out);
    if (ncols)
        *ncols = c;
    ABI_CATCH
}
int sb200_sym_destroy(sb200_sym_solver* s)
{
    ABI_TRY
    if (s)
        sym_destroy(s);
    ABI_CATCH
}
int sb200_sym_stats(const sb200_sym_solver* s, sb200_stats* out)
{
    ABI_TRY
    ABI_NONNULL(s);
    ABI_NONNULL(out);
    *out = sym_stats(s);
    ABI_CATCH
}
int sb200_sym_factorize_from(sb200_sym_solver* s, int64_t from_k, int64_t to_m)
{
    ABI_TRY
    ABI_NONNULL(s);
    sym_factorize_from(s, from_k, to_m);
    ABI_CATCH
}
int sb200_sym_get_factorization(sb200_sym_solver* s, double* V, double* H, double* f, double* beta, int64_t* k)
{
    ABI_TRY
    ABI_NONNULL(s);
    sym_get_factorization(s, V, H, f, beta, k);
    ABI_CATCH
}

// ---- GenEigsSolver ----
int sb200_gen_create(sb200_op* op, int64_t nev, int64_t ncv, sb200_gen_solver** out)
{
    ABI_TRY
    ABI_NONNULL(out);
    *out = gen_create(op, nev, ncv);
    ABI_CATCH
}
int sb200_gen_init(sb200_gen_solver* s, const double* init_resid_or_null)
{
    ABI_TRY
    ABI_NONNULL(s);
    gen_init(s, init_resid_or_null);
    ABI_CATCH
}
int sb200_gen_compute(sb200_gen_solver* s, int selection, int64_t maxit, double tol, int sorting, int64_t* nconv)
{
    ABI_TRY
    ABI_NONNULL(s);
    const int64_t r = gen_compute(s, selection, maxit, tol, sorting);
    if (nconv)
        *nconv = r;
    ABI_CATCH
}
int sb200_gen_info(const sb200_gen_solver* s, int* info)
{
    ABI_TRY
    ABI_NONNULL(s);
    *info = gen_info(s);
    ABI_CATCH
}
int sb200_gen_num_iterations(const sb200_gen_solver* s, int64_t* niter)
{
    ABI_TRY
    ABI_NONNULL(s);
    *niter = gen_niter(s);
    ABI_CATCH
}
int sb200_gen_num_operations(const sb200_gen_solver* s, int64_t* nops)
{
    ABI_TRY
    ABI_NONNULL(s);
    *nops = gen_nops(s);
    ABI_CATCH
}
int sb200_gen_eigenvalues(const sb200_gen_solver* s, double* out_ri, int64_t* count)
{
    ABI_TRY
    ABI_NONNULL(s);
    ABI_NONNULL(out_ri);
    const int64_t c = gen_eigenvalues(s, out_ri);
    if (count)
        *count = c;
    ABI_CATCH
}
int sb200_gen_eigenvectors(sb200_gen_solver* s, int64_t nvec, double* out_ri, int64_t* ncols)
{
    ABI_TRY
    ABI_NONNULL(s);
    ABI_NONNULL(out_ri);
    const int64_t c = gen_eigenvectors(s, nvec, out_ri);
    if (ncols)
        *ncols = c;
    ABI_CATCH
}
int sb200_gen_stats(const sb200_gen_solver* s, sb200_stats* out)
{
    ABI_TRY
    ABI_NONNULL(s);
    ABI_NONNULL(out);
    *out = gen_stats(s);
    ABI_CATCH
}
int sb200_gen_factorize_from(sb200_gen_solver* s, int64_t from_k, int64_t to_m)
{
    ABI_TRY
    ABI_NONNULL(s);
    gen_factorize_from(s, from_k, to_m);
    ABI_CATCH
}
int sb200_gen_get_factorization(sb200_gen_solver* s, double* V, double* H, double* f, double* beta, int64_t* k)
{
    ABI_TRY
    ABI_NONNULL(s);
    gen_get_factorization(s, V, H, f, beta, k);
    ABI_CATCH
}
int sb200_gen_destroy(sb200_gen_solver* s)
{
    ABI_TRY
    if (s)
        gen_destroy(s);
    ABI_CATCH
}

// ---- small dense kernels (unit tier) ----
int sb200_dense_tridiag_eigen(int64_t m, const double* H, double* evals, double* evecs)
{
    ABI_TRY
    ABI_NONNULL(H);
    ABI_NONNULL(evals);
    ABI_NONNULL(evecs);
    device_info();
    SB200_REQUIRE(m >= 1 && m <= kMaxNcv, SB200_INVALID_ARGUMENT, "matrix order out of range");
    DevBuf<double> dH(m * m), dE(m), dZ(m * m);
    DevBuf<int> dinfo(1);
    SB200_CUDA_CHECK(cudaMemcpy(dH.get(), H, sizeof(double) * m * m, cudaMemcpyHostToDevice));
    launch_tridiag_eigen(dH.get(), (int) m, dE.get(), dZ.get(), dinfo.get(), 0);
    SB200_CUDA_CHECK(cudaDeviceSynchronize());
    int info = 0;
    SB200_CUDA_CHECK(cudaMemcpy(&info, dinfo.get(), sizeof(int), cudaMemcpyDeviceToHost));
    if (info != 0)
        throw Error(SB200_RUNTIME, "TridiagEigen: eigen decomposition failed");
    SB200_CUDA_CHECK(cudaMemcpy(evals, dE.get(), sizeof(double) * m, cudaMemcpyDeviceToHost));
    SB200_CUDA_CHECK(cudaMemcpy(evecs, dZ.get(), sizeof(double) * m * m, cudaMemcpyDeviceToHost));
    ABI_CATCH
}
int sb200_dense_givens(int variant, int64_t count, const double* x, const double* y, double* r, double* c, double* s)
{
    ABI_TRY
    ABI_NONNULL(x);
    ABI_NONNULL(y);
    ABI_NONNULL(r);
    ABI_NONNULL(c);
    ABI_NONNULL(s);
    device_info();
    SB200_REQUIRE(variant >= 0 && variant <= 2 && count >= 0, SB200_INVALID_ARGUMENT, "givens: bad arguments");
    if (count > 0)
    {
        DevBuf<double> dx(count), dy(count), dr(count), dc(count), ds(count);
        SB200_CUDA_CHECK(cudaMemcpy(dx.get(), x, sizeof(double) * count, cudaMemcpyHostToDevice));
        SB200_CUDA_CHECK(cudaMemcpy(dy.get(), y, sizeof(double) * count, cudaMemcpyHostToDevice));
        launch_givens_batch(variant, count, dx.get(), dy.get(), dr.get(), dc.get(), ds.get(), 0);
        SB200_CUDA_CHECK(cudaDeviceSynchronize());
        SB200_CUDA_CHECK(cudaMemcpy(r, dr.get(), sizeof(double) * count, cudaMemcpyDeviceToHost));
        SB200_CUDA_CHECK(cudaMemcpy(c, dc.get(), sizeof(double) * count, cudaMemcpyDeviceToHost));
        SB200_CUDA_CHECK(cudaMemcpy(s, ds.get(), sizeof(double) * count, cudaMemcpyDeviceToHost));
    }
    ABI_CATCH
}
int sb200_dense_shifted_qr(int kind, int64_t m, const double* H, double shift, double* QtHQ, double* Q)
{
    ABI_TRY
    ABI_NONNULL(H);
    ABI_NONNULL(QtHQ);
    ABI_NONNULL(Q);
    device_info();
    SB200_REQUIRE(m >= 2 && m <= kMaxNcv, SB200_INVALID_ARGUMENT, "matrix order out of range");
    if (kind == 0)
    {
        DevBuf<double> dH(m * m), dD(m * m), dQ(m * m);
        SB200_CUDA_CHECK(cudaMemcpy(dH.get(), H, sizeof(double) * m * m, cudaMemcpyHostToDevice));
        launch_tridiag_qr(dH.get(), (int) m, shift, dD.get(), dQ.get(), 0);
        SB200_CUDA_CHECK(cudaDeviceSynchronize());
        SB200_CUDA_CHECK(cudaMemcpy(QtHQ, dD.get(), sizeof(double) * m * m, cudaMemcpyDeviceToHost));
        SB200_CUDA_CHECK(cudaMemcpy(Q, dQ.get(), sizeof(double) * m * m, cudaMemcpyDeviceToHost));
    }
    else
        dense_hess_qr_host(m, H, shift, QtHQ, Q);
    ABI_CATCH
}
int sb200_dense_double_shift_qr(int64_t m, const double* H, double s, double t, double* QtHQ, double* Q)
{
    ABI_TRY
    ABI_NONNULL(H);
    ABI_NONNULL(QtHQ);
    ABI_NONNULL(Q);
    dense_double_shift_qr_host(m, H, s, t, QtHQ, Q);
    ABI_CATCH
}
int sb200_dense_hess_eigen(int64_t m, const double* H, double* evals_ri, double* evecs_ri)
{
    ABI_TRY
    ABI_NONNULL(H);
    ABI_NONNULL(evals_ri);
    ABI_NONNULL(evecs_ri);
    dense_hess_eigen_host(m, H, evals_ri, evecs_ri);
    ABI_CATCH
}
int sb200_dense_shifted_qr_z(int64_t m, const double* H_ri, double shift_re, double shift_im, double* QtHQ_ri, double* Q_ri)
{
    ABI_TRY
    ABI_NONNULL(H_ri);
    ABI_NONNULL(QtHQ_ri);
    ABI_NONNULL(Q_ri);
    dense_hess_qr_z_host(m, H_ri, shift_re, shift_im, QtHQ_ri, Q_ri);
    ABI_CATCH
}
int sb200_dense_hess_eigen_z(int64_t m, const double* H_ri, double* evals_ri, double* evecs_ri)
{
    ABI_TRY
    ABI_NONNULL(H_ri);
    ABI_NONNULL(evals_ri);
    ABI_NONNULL(evecs_ri);
    dense_hess_eigen_z_host(m, H_ri, evals_ri, evecs_ri);
    ABI_CATCH
}
int sb200_dense_sym_restart(int64_t m, const double* H, double beta, int64_t nev, int selection, double tol, double* ritz_val, double* ritz_est,
                            int32_t* conv, int64_t* nconv, int64_t* k, double* Q, double* Hnew)
{
    ABI_TRY
    ABI_NONNULL(H);
    SB200_REQUIRE(m >= 2 && m <= kMaxNcv && nev >= 1 && nev < m, SB200_INVALID_ARGUMENT, "bad dimensions");
    dense_sym_restart_host(m, H, beta, nev, selection, tol, ritz_val, ritz_est, conv, nconv, k, Q, Hnew);
    ABI_CATCH
}
int sb200_dense_compress(int64_t n, int64_t m, int64_t kk, const double* V, const double* Q, const double* H, double* Vout, double* f, double* fnorm2,
                         int impl)
{
    ABI_TRY
    ABI_NONNULL(V);
    ABI_NONNULL(Q);
    ABI_NONNULL(Vout);
    SB200_REQUIRE(n >= 1 && m >= 1 && m <= kMaxNcv && kk >= 1 && kk <= m, SB200_INVALID_ARGUMENT, "bad dimensions");
    SB200_REQUIRE(!f || (H && fnorm2 && kk >= 2), SB200_INVALID_ARGUMENT, "residual update needs H, fnorm2 and kk >= 2");
    dense_compress_host(n, (int) m, (int) kk, V, Q, H, Vout, f, fnorm2, impl);
    ABI_CATCH
}

}  // extern "C"
#pragma GCC visibility pop

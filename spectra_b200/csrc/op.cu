// Device-resident operator behind sb200_op (replaces Sparse{Sym,Gen}MatProd).
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "host.h"

namespace sb200 {

int g_profiling_level = 0;

static thread_local std::string t_last_error;
void set_last_error(const std::string& msg) { t_last_error = msg; }
const char* last_error_cstr() { return t_last_error.c_str(); }

const DeviceInfo& device_info()
{
    static DeviceInfo info;
    static std::once_flag once;
    static std::string err;
    std::call_once(once, [] {
        int count = 0;
        cudaError_t e = cudaGetDeviceCount(&count);
        if (e != cudaSuccess || count == 0)
        {
            err = std::string("no CUDA device available: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
            cudaGetLastError();
            return;
        }
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess)
        {
            err = "cudaGetDevice failed";
            return;
        }
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, dev) != cudaSuccess)
        {
            err = "cudaGetDeviceProperties failed";
            return;
        }
        info.device = dev;
        info.sm_count = p.multiProcessorCount;
        info.cc_major = p.major;
        info.cc_minor = p.minor;
        info.total_mem = p.totalGlobalMem;
        info.l2_bytes = p.l2CacheSize;
        if (p.major < 10)
            err = "this library contains sm_100a code only; device compute capability is " + std::to_string(p.major) + "." + std::to_string(p.minor);
    });
    if (!err.empty())
        throw Error(SB200_CUDA, err);
    return info;
}

}  // namespace sb200

using namespace sb200;

sb200_op::~sb200_op()
{
    if (band)
        band_destroy(band);
    if (ev0)
        cudaEventDestroy(ev0);
    if (ev1)
        cudaEventDestroy(ev1);
    for (cudaEvent_t e : ev_chunk)
        cudaEventDestroy(e);
    if (ev_ready)
        cudaEventDestroy(ev_ready);
    if (comm_stream)
        cudaStreamDestroy(comm_stream);
    sb200::peer_window_destroy(comm, win_x);
    sb200::peer_window_destroy(comm, win_ctl);
    if (stream)
        cudaStreamDestroy(stream);
}

namespace sb200 {

// Number of partial all-gathers per operator application on a sharded operator: enough that one slice of the operand is
// L2-resident (choose_col_blocks), and at least 2 so that the second collective hides behind the SpMV of the first chunk.
// Measured on 4 GPUs at n = 1e7 (tools/mgpu_chunks.py): 1 chunk 571 ms, 2 chunks 519 ms, 4 chunks 550 ms, 8 chunks 618 ms
// per 541 operations -- more chunks shorten the exposed collective but split the rows into ever shorter pieces.
// SB200_AG_CHUNKS overrides.
static int choose_chunks(int64_t n)
{
    int k = std::max(choose_col_blocks(n), 2);
    if (const char* e = std::getenv("SB200_AG_CHUNKS"))
        k = std::atoi(e);
    return std::max(1, std::min(k, kMaxColBlocks));
}

static void finish_op(sb200_op* op)
{
    const int P = op->nranks();
    op->slab = (op->A.n + P - 1) / P;
    // test hook: SB200_FORCE_CHUNK_RANKS=R lays a single-GPU operator out as if it were one of R ranks' (chunk-major operand,
    // remapped column ids), so that the chunked split / permutation / per-block SpMV are covered without a second GPU
    int virt = 0;
    if (P == 1 && !op->cb)
        if (const char* e = std::getenv("SB200_FORCE_CHUNK_RANKS"))
            virt = std::max(0, std::atoi(e));
    if (P > 1 || virt > 0)
    {
        const int ranks = P > 1 ? P : virt;
        const int64_t slab = (op->A.n + ranks - 1) / ranks;
        split_column_chunks(op->A, choose_chunks(op->A.n), slab, ranks, op->stream);
        const int nb = (int) op->A.blocks.size();
        const size_t xcount = (size_t) (op->A.chunk_stride() * nb);
        // Peer mode (default on a multi-rank communicator, SB200_PEER=0 keeps the NCCL collectives): the operand buffer and the mailboxes
        // of the one-shot all-reduce are symmetric windows that every rank maps (peer.cu); falls back when the mapping is refused.
        const char* pe = std::getenv("SB200_PEER");
        if (P > 1 && !(pe && pe[0] == '0'))
        {
            if (peer_window_create(op->comm, sizeof(double) * xcount, op->win_x, op->stream))
            {
                if (!peer_window_create(op->comm, peer_ctl_bytes(P), op->win_ctl, op->stream))
                    peer_window_destroy(op->comm, op->win_x);
            }
        }
        if (op->peer_mode())
        {
            op->xc = static_cast<double*>(op->win_x.local);
            op->peer_seq.alloc(1);
            op->peer_seq.zero(op->stream);
        }
        else
        {
            op->x_chunks.alloc(xcount);
            op->x_chunks.zero(op->stream);
            op->xc = op->x_chunks.get();
        }
    }
    if (P > 1)
    {
        const int nb = (int) op->A.blocks.size();
        SB200_CUDA_CHECK(cudaStreamCreateWithFlags(&op->comm_stream, cudaStreamNonBlocking));
        SB200_CUDA_CHECK(cudaEventCreateWithFlags(&op->ev_ready, cudaEventDisableTiming));
        op->ev_chunk.resize((size_t) nb);
        for (cudaEvent_t& e : op->ev_chunk)
            SB200_CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    else
        split_column_blocks(op->A, choose_col_blocks(op->A.n), op->stream);
    // Device layout of the column blocks.  Default ("auto"): the sliced layout (SELL-32 in 1024-row windows, lane-per-row kernels, spmv.cu)
    // whenever its padding stays below SB200_SELL_MAX_FILL (1.3) -- measured on a B200 at n = 1e7, 20 nnz/row: 0.884 ms vs 1.051 ms for
    // the CSR-vector kernels on uniformly random columns, 0.505 vs 0.698 ms on band columns (profiles/r2_spmv_variants_n1e7.log).
    // Operands with very uneven rows (fill above the limit) keep the CSR-vector kernels; SB200_SPMV_FORMAT=csr forces them.
    {
        const char* e = std::getenv("SB200_SPMV_FORMAT");
        const bool want_sell = (e == nullptr) || std::strcmp(e, "sell") == 0 || std::strcmp(e, "auto") == 0;
        SB200_REQUIRE(want_sell || std::strcmp(e, "csr") == 0, SB200_INVALID_ARGUMENT, "SB200_SPMV_FORMAT must be auto, sell or csr");
        if (want_sell)
        {
            double max_fill = 1.3;
            if (const char* f = std::getenv("SB200_SELL_MAX_FILL"))
                max_fill = std::max(1.0, std::atof(f));
            build_sell_layout(op->A, max_fill, op->stream);
        }
    }
    op->plan = make_spmv_plan(op->A);
    if (P > 1)
    {
        op->x_full.alloc((size_t) op->slab * P);
        op->x_full.zero(op->stream);
        op->x_stage.alloc((size_t) op->slab);
        op->x_stage.zero(op->stream);
    }
    SB200_CUDA_CHECK(cudaEventCreate(&op->ev0));
    SB200_CUDA_CHECK(cudaEventCreate(&op->ev1));
    SB200_CUDA_CHECK(cudaStreamSynchronize(op->stream));
}

sb200_op* op_create_sparse(int64_t n, const void* outer, int outer_is_64, const int32_t* inner, const double* values, int storage_order, int matrix_mode,
                           sb200_comm* comm)
{
    device_info();
    SB200_REQUIRE(outer != nullptr, SB200_INVALID_ARGUMENT, "null matrix arrays");  // inner / values may be null for an empty matrix
    std::unique_ptr<sb200_op> op(new sb200_op());
    op->comm = comm;
    SB200_CUDA_CHECK(cudaStreamCreateWithFlags(&op->stream, cudaStreamNonBlocking));
    const int P = op->nranks();
    const int64_t slab = (n + P - 1) / P;
    const int64_t row0 = std::min<int64_t>(n, slab * op->rank());
    const int64_t nrows = std::max<int64_t>(0, std::min<int64_t>(slab, n - row0));
    build_device_csr(n, outer, outer_is_64 != 0, inner, values, storage_order, matrix_mode, row0, nrows, op->stream, op->A);
    op->symmetric_hint = (matrix_mode != SB200_GENERAL);
    finish_op(op.get());
    return op.release();
}

// SparseSymShiftSolve(mat) (SparseSymShiftSolve.h:57-68): the matrix is uploaded like SparseSymMatProd's; set_shift() factorises
sb200_op* op_create_shift_solve(int64_t n, const void* outer, int outer_is_64, const int32_t* inner, const double* values, int storage_order, int matrix_mode)
{
    std::unique_ptr<sb200_op> op(op_create_sparse(n, outer, outer_is_64, inner, values, storage_order, matrix_mode, nullptr));
    op->band = band_create(op.get());
    return op.release();
}

sb200_op* op_create_csr_slab(int64_t n, int64_t row0, int64_t nrows, const int64_t* rowptr_local, const int32_t* col, const double* values, sb200_comm* comm)
{
    device_info();
    std::unique_ptr<sb200_op> op(new sb200_op());
    op->comm = comm;
    SB200_CUDA_CHECK(cudaStreamCreateWithFlags(&op->stream, cudaStreamNonBlocking));
    const int P = op->nranks();
    const int64_t slab = (n + P - 1) / P;
    SB200_REQUIRE(row0 == std::min<int64_t>(n, slab * op->rank()) && nrows == std::max<int64_t>(0, std::min<int64_t>(slab, n - row0)), SB200_INVALID_ARGUMENT,
                  "row slab must be rows [rank*ceil(n/P), ...) of the matrix");
    upload_csr_slab(n, row0, nrows, rowptr_local, col, values, op->stream, op->A);
    finish_op(op.get());
    return op.release();
}

// SparseHermMatProd<std::complex<double>, Uplo, Flags>(mat) (MatOp/SparseHermMatProd.h:46-54): values are interleaved (re, im)
sb200_op* op_create_sparse_herm(int64_t n, const void* outer, int outer_is_64, const int32_t* inner, const double* values_ri, int storage_order, int matrix_mode)
{
    device_info();
    SB200_REQUIRE(outer != nullptr, SB200_INVALID_ARGUMENT, "null matrix arrays");
    std::unique_ptr<sb200_op> op(new sb200_op());
    SB200_CUDA_CHECK(cudaStreamCreateWithFlags(&op->stream, cudaStreamNonBlocking));
    build_device_csr_z(n, outer, outer_is_64 != 0, inner, values_ri, storage_order, matrix_mode, op->stream, op->Az);
    op->cplx = true;
    op->A.n = n;
    op->A.row0 = 0;
    op->A.nrows = n;
    op->A.nnz = op->Az.nnz;
    op->slab = n;
    op->plan.grid = 1;
    op->symmetric_hint = (matrix_mode != SB200_GENERAL);
    SB200_CUDA_CHECK(cudaEventCreate(&op->ev0));
    SB200_CUDA_CHECK(cudaEventCreate(&op->ev1));
    return op.release();
}

// complex = true: the user's function works on interleaved complex vectors (2 n doubles), for HermEigsSolver
sb200_op* op_create_callback(int64_t n, void (*fn)(const double*, double*, void*), void* user, bool is_complex)
{
    device_info();
    SB200_REQUIRE(fn != nullptr, SB200_INVALID_ARGUMENT, "null operator callback");
    SB200_REQUIRE(n >= 1 && n < (1LL << (is_complex ? 30 : 31)), SB200_INVALID_ARGUMENT, "matrix order out of range");
    std::unique_ptr<sb200_op> op(new sb200_op());
    SB200_CUDA_CHECK(cudaStreamCreateWithFlags(&op->stream, cudaStreamNonBlocking));
    op->A.n = n;
    op->A.row0 = 0;
    op->A.nrows = n;
    op->A.nnz = 0;
    op->cb = fn;
    op->cb_user = user;
    op->cplx = is_complex;
    op->hx.alloc((size_t) n * (is_complex ? 2 : 1));
    op->hy.alloc((size_t) n * (is_complex ? 2 : 1));
    op->slab = n;
    op->plan.grid = 1;
    SB200_CUDA_CHECK(cudaEventCreate(&op->ev0));
    SB200_CUDA_CHECK(cudaEventCreate(&op->ev1));
    return op.release();
}

// y_dev = A * x_dev through the user's host function (device pointers in, device pointers out)
void op_callback_device(sb200_op* op, const double* x_dev, double* y_dev)
{
    const int64_t n = op->A.n * (op->cplx ? 2 : 1);
    SB200_CUDA_CHECK(cudaMemcpyAsync(op->hx.get(), x_dev, sizeof(double) * n, cudaMemcpyDeviceToHost, op->stream));
    SB200_CUDA_CHECK(cudaStreamSynchronize(op->stream));
    op->cb(op->hx.get(), op->hy.get(), op->cb_user);
    SB200_CUDA_CHECK(cudaMemcpyAsync(y_dev, op->hy.get(), sizeof(double) * n, cudaMemcpyHostToDevice, op->stream));
}

// y_dev (local rows) = A * x_dev (full vector).  Sharded: x_dev must already hold all n entries.
void op_spmv_device(sb200_op* op, const double* x_dev, double* y_dev)
{
    if (op->cb)
        op_callback_device(op, x_dev, y_dev);
    else if (op->cplx)
        launch_spmv_z(op->Az, x_dev, y_dev, op->stream);
    else if (op->band)
        band_solve_device(op, x_dev, y_dev);
    else if (op->A.chunk_len)
    {
        // natural-layout operand -> chunk-major operand of the sharded operator (cold paths only: init, perform_op)
        launch_permute_to_chunks(op->A, x_dev, op->xc, op->stream);
        launch_spmv(op->A, op->plan, op->xc, y_dev, op->stream);
    }
    else
        launch_spmv(op->A, op->plan, x_dev, y_dev, op->stream);
}

// Host-pointer perform_op (SparseSymMatProd.h:83-88): H2D x, kernel, D2H y (local rows).
void op_perform_op_host(sb200_op* op, const double* x_host, double* y_host)
{
    if (op->cb)
    {
        op->cb(x_host, y_host, op->cb_user);
        return;
    }
    const int cw = op->cplx ? 2 : 1;  // doubles per scalar
    const int64_t n = op->A.n * cw, nloc = op->A.nrows * cw;
    if (op->x_full.n < (size_t) n)
        op->x_full.alloc((size_t) n);
    if (op->y_loc.n < (size_t) std::max<int64_t>(nloc, 1))
        op->y_loc.alloc((size_t) std::max<int64_t>(nloc, 1));
    SB200_CUDA_CHECK(cudaMemcpyAsync(op->x_full.get(), x_host, sizeof(double) * n, cudaMemcpyHostToDevice, op->stream));
    op_spmv_device(op, op->x_full.get(), op->y_loc.get());
    if (nloc > 0)
        SB200_CUDA_CHECK(cudaMemcpyAsync(y_host, op->y_loc.get(), sizeof(double) * nloc, cudaMemcpyDeviceToHost, op->stream));
    SB200_CUDA_CHECK(cudaStreamSynchronize(op->stream));
}

}  // namespace sb200

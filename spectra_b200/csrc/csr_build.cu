// Device-side construction of the full CSR operand from the user's compressed arrays.
//
// Semantics follow the reference wrappers:
//   SB200_GENERAL   : y = M x with every stored entry            (MatOp/SparseGenMatProd.h:82-87)
//   SB200_SYM_LOWER : y = M.selfadjointView<Lower>() x — only entries with row >= col are read and
//                     mirrored; the strictly-upper stored entries are ignored
//                                                                  (MatOp/SparseSymMatProd.h:83-88)
//   SB200_SYM_UPPER : same with row <= col.
// ColMajor input (Eigen default, CSC) is transposed on the fly.  The arrays are uploaded once; the
// expansion / transpose (count -> exclusive scan -> scatter -> per-row sort by column) runs on the
// GPU, so the one-time cost is dominated by the H2D copy.  Duplicate entries of an uncompressed
// input are kept as separate terms (they add up in the SpMV), columns ascend inside a row, so the
// SpMV summation order is reproducible.
#include "kernels.h"

namespace sb200 {

namespace {

template <typename OuterT>
__global__ void count_kernel(const OuterT* __restrict__ outer, const int* __restrict__ inner, int64_t n, int order, int mode, int64_t row0, int64_t nrows,
                             int* __restrict__ cnt)
{
    for (int64_t o = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (int64_t) gridDim.x * blockDim.x)
    {
        const int64_t pb = outer[o], pe = outer[o + 1];
        for (int64_t p = pb; p < pe; p++)
        {
            const int64_t in = inner[p];
            const int64_t i = (order == SB200_COL_MAJOR) ? in : o;
            const int64_t j = (order == SB200_COL_MAJOR) ? o : in;
            if (mode == SB200_GENERAL)
            {
                if (i >= row0 && i < row0 + nrows)
                    atomicAdd(cnt + (i - row0), 1);
            }
            else
            {
                const bool used = (mode == SB200_SYM_LOWER || mode == SB200_HERM_LOWER) ? (i >= j) : (i <= j);
                if (!used)
                    continue;
                if (i >= row0 && i < row0 + nrows)
                    atomicAdd(cnt + (i - row0), 1);
                if (i != j && j >= row0 && j < row0 + nrows)
                    atomicAdd(cnt + (j - row0), 1);
            }
        }
    }
}

// Value handling of the mirrored / diagonal entries: real symmetric entries are copied; Hermitian ones (complex values stored as
// double2 = (re, im), selfadjointView<Uplo> of a complex matrix, MatOp/SparseHermMatProd.h:83-88) are conjugated when mirrored and
// their diagonal is taken as real.
__device__ __forceinline__ double mirror_value(double v, bool) { return v; }
__device__ __forceinline__ double2 mirror_value(double2 v, bool herm) { return herm ? make_double2(v.x, -v.y) : v; }
__device__ __forceinline__ double diag_value(double v, bool) { return v; }
__device__ __forceinline__ double2 diag_value(double2 v, bool herm) { return herm ? make_double2(v.x, 0.0) : v; }

template <typename OuterT, typename VT>
__global__ void fill_kernel(const OuterT* __restrict__ outer, const int* __restrict__ inner, const VT* __restrict__ values, int64_t n, int order, int mode,
                            int64_t row0, int64_t nrows, const int* __restrict__ rowptr, int* __restrict__ cursor, int* __restrict__ col, VT* __restrict__ val)
{
    const bool herm = (mode == SB200_HERM_LOWER || mode == SB200_HERM_UPPER);
    const bool lower = (mode == SB200_SYM_LOWER || mode == SB200_HERM_LOWER);
    for (int64_t o = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (int64_t) gridDim.x * blockDim.x)
    {
        const int64_t pb = outer[o], pe = outer[o + 1];
        for (int64_t p = pb; p < pe; p++)
        {
            const int64_t in = inner[p];
            const int64_t i = (order == SB200_COL_MAJOR) ? in : o;
            const int64_t j = (order == SB200_COL_MAJOR) ? o : in;
            const VT v = values[p];
            bool put_ij = false, put_ji = false;
            if (mode == SB200_GENERAL)
                put_ij = true;
            else
            {
                const bool used = lower ? (i >= j) : (i <= j);
                put_ij = used;
                put_ji = used && (i != j);
            }
            if (put_ij && i >= row0 && i < row0 + nrows)
            {
                const int li = (int) (i - row0);
                const int q = rowptr[li] + atomicAdd(cursor + li, 1);
                col[q] = (int) j;
                val[q] = (i == j && mode != SB200_GENERAL) ? diag_value(v, herm) : v;
            }
            if (put_ji && j >= row0 && j < row0 + nrows)
            {
                const int lj = (int) (j - row0);
                const int q = rowptr[lj] + atomicAdd(cursor + lj, 1);
                col[q] = (int) i;
                val[q] = mirror_value(v, herm);
            }
        }
    }
}

// Per-row insertion sort by column id (rows are short; long rows only appear in small test inputs).
template <typename VT>
__global__ void sort_rows_kernel(const int* __restrict__ rowptr, int* __restrict__ col, VT* __restrict__ val, int64_t nrows)
{
    for (int64_t r = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += (int64_t) gridDim.x * blockDim.x)
    {
        const int b = rowptr[r], e = rowptr[r + 1];
        for (int p = b + 1; p < e; p++)
        {
            const int c = col[p];
            const VT v = val[p];
            int q = p - 1;
            while (q >= b && col[q] > c)
            {
                col[q + 1] = col[q];
                val[q + 1] = val[q];
                q--;
            }
            col[q + 1] = c;
            val[q + 1] = v;
        }
    }
}

// ---- exclusive scan of int counts (three-phase) ----
constexpr int kScanBlock = 1024;
constexpr int kScanItems = 4;  // items per thread

__global__ void scan_block_sums_kernel(const int* __restrict__ cnt, int64_t n, long long* __restrict__ block_sums)
{
    __shared__ long long s[kScanBlock / 32];
    const int64_t base = (int64_t) blockIdx.x * kScanBlock * kScanItems;
    long long acc = 0;
    for (int t = 0; t < kScanItems; t++)
    {
        const int64_t idx = base + (int64_t) threadIdx.x * kScanItems + t;
        if (idx < n)
            acc += cnt[idx];
    }
    for (int o = 16; o > 0; o >>= 1)
        acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0)
        s[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0)
    {
        long long t = 0;
        for (int q = 0; q < kScanBlock / 32; q++)
            t += s[q];
        block_sums[blockIdx.x] = t;
    }
}

__global__ void scan_offsets_kernel(long long* block_sums, int nblocks, long long* total)
{
    // single thread: nblocks is a few thousand at most
    long long run = 0;
    for (int b = 0; b < nblocks; b++)
    {
        const long long t = block_sums[b];
        block_sums[b] = run;
        run += t;
    }
    *total = run;
}

__global__ void scan_write_kernel(const int* __restrict__ cnt, int64_t n, const long long* __restrict__ block_offs, int* __restrict__ rowptr)
{
    __shared__ long long s_warp[kScanBlock / 32];
    const int64_t base = (int64_t) blockIdx.x * kScanBlock * kScanItems;
    int v[kScanItems];
    long long mine = 0;
    for (int t = 0; t < kScanItems; t++)
    {
        const int64_t idx = base + (int64_t) threadIdx.x * kScanItems + t;
        v[t] = (idx < n) ? cnt[idx] : 0;
        mine += v[t];
    }
    // inclusive scan of `mine` across the block
    long long incl = mine;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int o = 1; o < 32; o <<= 1)
    {
        const long long t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o)
            incl += t;
    }
    if (lane == 31)
        s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0)
    {
        long long w = s_warp[lane];
        for (int o = 1; o < 32; o <<= 1)
        {
            const long long t = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o)
                w += t;
        }
        s_warp[lane] = w;
    }
    __syncthreads();
    long long excl = incl - mine + (warp > 0 ? s_warp[warp - 1] : 0) + block_offs[blockIdx.x];
    for (int t = 0; t < kScanItems; t++)
    {
        const int64_t idx = base + (int64_t) threadIdx.x * kScanItems + t;
        if (idx < n)
            rowptr[idx] = (int) excl;
        excl += v[t];
    }
}

// rowptr[n] = total, cursor reset happens by the caller
__global__ void set_last_kernel(int* rowptr, int64_t n, const long long* total) { rowptr[n] = (int) *total; }

__global__ void narrow_rowptr_kernel(const long long* __restrict__ src, long long base, int* __restrict__ dst, int64_t n)
{
    for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t) gridDim.x * blockDim.x)
        dst[i] = (int) (src[i] - base);
}

int grid_for(int64_t n, int block)
{
    const int sms = device_info().sm_count;
    const int64_t need = (n + block - 1) / block;
    return (int) std::max<int64_t>(1, std::min<int64_t>(need, (int64_t) sms * 16));
}

// Destination arrays of one build (real DeviceCsr or complex DeviceCsrZ)
template <typename VT>
struct CsrDest
{
    int64_t& nnz;
    DevBuf<int>& rowptr;
    DevBuf<int>& col;
    DevBuf<VT>& val;
};

template <typename OuterT, typename VT>
void build_impl(int64_t n, const OuterT* h_outer, const int32_t* h_inner, const VT* h_values, int order, int mode, int64_t row0, int64_t nrows,
                cudaStream_t stream, CsrDest<VT> out)
{
    const int64_t nnz_in = (int64_t) h_outer[n] - (int64_t) h_outer[0];
    SB200_REQUIRE(h_outer[0] == 0, SB200_INVALID_ARGUMENT, "sparse matrix must be in compressed form (outer[0] == 0)");
    DevBuf<OuterT> d_outer(n + 1);
    DevBuf<int> d_inner(std::max<int64_t>(nnz_in, 1));
    DevBuf<VT> d_values(std::max<int64_t>(nnz_in, 1));
    SB200_CUDA_CHECK(cudaMemcpyAsync(d_outer.get(), h_outer, sizeof(OuterT) * (n + 1), cudaMemcpyHostToDevice, stream));
    if (nnz_in > 0)
    {
        SB200_CUDA_CHECK(cudaMemcpyAsync(d_inner.get(), h_inner, sizeof(int) * nnz_in, cudaMemcpyHostToDevice, stream));
        SB200_CUDA_CHECK(cudaMemcpyAsync(d_values.get(), h_values, sizeof(VT) * nnz_in, cudaMemcpyHostToDevice, stream));
    }

    out.rowptr.alloc(nrows + 1);
    DevBuf<int> cnt(std::max<int64_t>(nrows, 1));
    cnt.zero(stream);
    const int g = grid_for(n, 256);
    count_kernel<OuterT><<<g, 256, 0, stream>>>(d_outer.get(), d_inner.get(), n, order, mode, row0, nrows, cnt.get());
    SB200_CUDA_CHECK(cudaGetLastError());

    const int nblocks = (int) ((nrows + (int64_t) kScanBlock * kScanItems - 1) / ((int64_t) kScanBlock * kScanItems));
    DevBuf<long long> block_sums(std::max(nblocks, 1) + 1);
    long long* d_total = block_sums.get() + std::max(nblocks, 1);
    if (nrows > 0)
    {
        scan_block_sums_kernel<<<nblocks, kScanBlock, 0, stream>>>(cnt.get(), nrows, block_sums.get());
        scan_offsets_kernel<<<1, 1, 0, stream>>>(block_sums.get(), nblocks, d_total);
        scan_write_kernel<<<nblocks, kScanBlock, 0, stream>>>(cnt.get(), nrows, block_sums.get(), out.rowptr.get());
        set_last_kernel<<<1, 1, 0, stream>>>(out.rowptr.get(), nrows, d_total);
        SB200_CUDA_CHECK(cudaGetLastError());
    }
    long long total = 0;
    if (nrows > 0)
        SB200_CUDA_CHECK(cudaMemcpyAsync(&total, d_total, sizeof(long long), cudaMemcpyDeviceToHost, stream));
    SB200_CUDA_CHECK(cudaStreamSynchronize(stream));
    SB200_REQUIRE(total < (1LL << 31), SB200_INVALID_ARGUMENT, "local nnz exceeds the int32 index range");
    out.nnz = total;
    out.col.alloc(std::max<int64_t>(total, 1));
    out.val.alloc(std::max<int64_t>(total, 1));
    if (total > 0)
    {
        cnt.zero(stream);  // reuse as the per-row cursor
        fill_kernel<OuterT, VT><<<g, 256, 0, stream>>>(d_outer.get(), d_inner.get(), d_values.get(), n, order, mode, row0, nrows, out.rowptr.get(), cnt.get(),
                                                       out.col.get(), out.val.get());
        sort_rows_kernel<VT><<<grid_for(nrows, 128), 128, 0, stream>>>(out.rowptr.get(), out.col.get(), out.val.get(), nrows);
        SB200_CUDA_CHECK(cudaGetLastError());
    }
    SB200_CUDA_CHECK(cudaStreamSynchronize(stream));
}

}  // namespace

void build_device_csr(int64_t n, const void* outer, bool outer64, const int32_t* inner, const double* values, int order, int mode, int64_t row0, int64_t nrows,
                      cudaStream_t stream, DeviceCsr& out)
{
    SB200_REQUIRE(n >= 1, SB200_INVALID_ARGUMENT, "matrix order must be positive");
    SB200_REQUIRE(n < (1LL << 31), SB200_INVALID_ARGUMENT, "matrix order exceeds the int32 index range");
    SB200_REQUIRE(order == SB200_COL_MAJOR || order == SB200_ROW_MAJOR, SB200_INVALID_ARGUMENT, "bad storage order");
    SB200_REQUIRE(mode >= SB200_GENERAL && mode <= SB200_SYM_UPPER, SB200_INVALID_ARGUMENT, "bad matrix mode");
    out.n = n;
    out.row0 = row0;
    out.nrows = nrows;
    CsrDest<double> dest{out.nnz, out.rowptr, out.col, out.val};
    if (outer64)
        build_impl<long long, double>(n, static_cast<const long long*>(outer), inner, values, order, mode, row0, nrows, stream, dest);
    else
        build_impl<int, double>(n, static_cast<const int*>(outer), inner, values, order, mode, row0, nrows, stream, dest);
}

// Complex operand (interleaved (re, im) values): SB200_HERM_LOWER / SB200_HERM_UPPER read one triangle and mirror it conjugated
// (selfadjointView<Uplo> of a complex matrix, MatOp/SparseHermMatProd.h:83-88); SB200_GENERAL keeps every stored entry.
void build_device_csr_z(int64_t n, const void* outer, bool outer64, const int32_t* inner, const double* values_ri, int order, int mode, cudaStream_t stream,
                        DeviceCsrZ& out)
{
    SB200_REQUIRE(n >= 1, SB200_INVALID_ARGUMENT, "matrix order must be positive");
    SB200_REQUIRE(n < (1LL << 30), SB200_INVALID_ARGUMENT, "matrix order exceeds the index range of the complex operator");
    SB200_REQUIRE(order == SB200_COL_MAJOR || order == SB200_ROW_MAJOR, SB200_INVALID_ARGUMENT, "bad storage order");
    SB200_REQUIRE(mode == SB200_GENERAL || mode == SB200_HERM_LOWER || mode == SB200_HERM_UPPER, SB200_INVALID_ARGUMENT, "bad matrix mode for a complex operator");
    out.n = n;
    CsrDest<double2> dest{out.nnz, out.rowptr, out.col, out.val};
    const double2* v2 = reinterpret_cast<const double2*>(values_ri);
    if (outer64)
        build_impl<long long, double2>(n, static_cast<const long long*>(outer), inner, v2, order, mode, 0, n, stream, dest);
    else
        build_impl<int, double2>(n, static_cast<const int*>(outer), inner, v2, order, mode, 0, n, stream, dest);
}

// ---------------------------------------------------------------------------------------------
// Column blocking: re-lay the (column-sorted) CSR out as nb sub-matrices over column ranges of width W
// ---------------------------------------------------------------------------------------------
namespace {

struct BlockPtrs
{
    int* rowptr[kMaxColBlocks];
    int* col[kMaxColBlocks];
    double* val[kMaxColBlocks];
};

// Column -> (block, id inside the block's operand slice).  RANGE: contiguous column ranges of width W, ids stay global.
// CHUNK: block = which 1/K-th of its owner rank's slab the column lies in, id = position in that partial all-gather.
struct ColMap
{
    int chunked;
    int64_t W, slab, len;
    int64_t W0 = 0;  // natural layout, two blocks of unequal width: columns below W0 form block 0 (0: equal widths W)
    __device__ __forceinline__ int block(int col) const
    {
        return chunked ? (int) (((int64_t) col % slab) / len) : (W0 > 0 ? (col >= W0 ? 1 : 0) : (int) (col / W));
    }
    __device__ __forceinline__ int remap(int col) const
    {
        return chunked ? (int) (((int64_t) col / slab) * len + ((int64_t) col % slab) % len) : col;
    }
};

// cnt[c * nrows + r] = number of entries of row r that fall into column block c
__global__ void block_count_kernel(const int* __restrict__ rowptr, const int* __restrict__ col, int64_t nrows, ColMap map, int nb, int* __restrict__ cnt)
{
    for (int64_t r = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += (int64_t) gridDim.x * blockDim.x)
    {
        int run[kMaxColBlocks];
#pragma unroll
        for (int c = 0; c < kMaxColBlocks; c++)
            run[c] = 0;
        for (int p = rowptr[r]; p < rowptr[r + 1]; p++)
            run[map.block(col[p])]++;
        for (int c = 0; c < nb; c++)
            cnt[(int64_t) c * nrows + r] = run[c];
    }
}

// Entries keep their order inside each block; for both maps ascending global columns give ascending remapped ids.
__global__ void block_fill_kernel(const int* __restrict__ rowptr, const int* __restrict__ col, const double* __restrict__ val, int64_t nrows, ColMap map, int nb,
                                  BlockPtrs bp)
{
    for (int64_t r = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += (int64_t) gridDim.x * blockDim.x)
    {
        int q[kMaxColBlocks];
        for (int c = 0; c < nb; c++)
            q[c] = bp.rowptr[c][r];
        for (int p = rowptr[r]; p < rowptr[r + 1]; p++)
        {
            const int cj = col[p];
            const int c = map.block(cj);
            bp.col[c][q[c]] = map.remap(cj);
            bp.val[c][q[c]] = val[p];
            q[c]++;
        }
    }
}

__global__ void permute_to_chunks_kernel(const double* __restrict__ x, double* __restrict__ xc, int64_t n, int64_t slab, int64_t len, int ranks, int nchunks)
{
    const int64_t stride = (int64_t) ranks * len, total = stride * nchunks;
    for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t) gridDim.x * blockDim.x)
    {
        const int64_t c = i / stride, rem = i % stride, r = rem / len, ii = rem % len;
        const int64_t local = c * len + ii, colg = r * slab + local;
        xc[i] = (local < slab && colg < n) ? x[colg] : 0.0;
    }
}

void exclusive_scan_to_rowptr(const int* cnt, int64_t nrows, int* rowptr, long long* h_total, cudaStream_t stream)
{
    const int nblocks = (int) ((nrows + (int64_t) kScanBlock * kScanItems - 1) / ((int64_t) kScanBlock * kScanItems));
    DevBuf<long long> block_sums(std::max(nblocks, 1) + 1);
    long long* d_total = block_sums.get() + std::max(nblocks, 1);
    scan_block_sums_kernel<<<nblocks, kScanBlock, 0, stream>>>(cnt, nrows, block_sums.get());
    scan_offsets_kernel<<<1, 1, 0, stream>>>(block_sums.get(), nblocks, d_total);
    scan_write_kernel<<<nblocks, kScanBlock, 0, stream>>>(cnt, nrows, block_sums.get(), rowptr);
    set_last_kernel<<<1, 1, 0, stream>>>(rowptr, nrows, d_total);
    SB200_CUDA_CHECK(cudaGetLastError());
    SB200_CUDA_CHECK(cudaMemcpyAsync(h_total, d_total, sizeof(long long), cudaMemcpyDeviceToHost, stream));
    SB200_CUDA_CHECK(cudaStreamSynchronize(stream));
}

}  // namespace

static void split_by_map(DeviceCsr& A, int nb, const ColMap& map, cudaStream_t stream)
{
    DevBuf<int> cnt((size_t) nb * A.nrows);
    block_count_kernel<<<grid_for(A.nrows, 128), 128, 0, stream>>>(A.rowptr.get(), A.col.get(), A.nrows, map, nb, cnt.get());
    SB200_CUDA_CHECK(cudaGetLastError());
    A.blocks.clear();
    A.blocks.resize(nb);
    BlockPtrs bp;
    for (int c = 0; c < kMaxColBlocks; c++)
    {
        bp.rowptr[c] = nullptr;
        bp.col[c] = nullptr;
        bp.val[c] = nullptr;
    }
    for (int c = 0; c < nb; c++)
    {
        CsrBlock& B = A.blocks[c];
        B.rowptr.alloc(A.nrows + 1);
        long long total = 0;
        exclusive_scan_to_rowptr(cnt.get() + (size_t) c * A.nrows, A.nrows, B.rowptr.get(), &total, stream);
        B.nnz = total;
        B.col.alloc(std::max<int64_t>(total, 1));
        B.val.alloc(std::max<int64_t>(total, 1));
        bp.rowptr[c] = B.rowptr.get();
        bp.col[c] = B.col.get();
        bp.val[c] = B.val.get();
    }
    block_fill_kernel<<<grid_for(A.nrows, 128), 128, 0, stream>>>(A.rowptr.get(), A.col.get(), A.val.get(), A.nrows, map, nb, bp);
    SB200_CUDA_CHECK(cudaGetLastError());
    SB200_CUDA_CHECK(cudaStreamSynchronize(stream));
    // the unblocked copy is no longer needed
    A.rowptr.release();
    A.col.release();
    A.val.release();
}

void split_column_blocks(DeviceCsr& A, int nb, cudaStream_t stream)
{
    if (nb <= 1 || A.nrows == 0 || A.nnz == 0)
        return;
    nb = std::min(nb, kMaxColBlocks);
    const int64_t W = (A.n + nb - 1) / nb;
    // A/B knob SB200_XSPLIT0_MB (two blocks only): size of the first operand slice.  A smaller first slice moves gathers from the head kernel
    // (gather bound, HBM idle) into the fused kernel, where they hide under the V stream -- as long as the larger slice still lives in L2.
    int64_t W0 = 0;
    if (nb == 2)
        if (const char* e = std::getenv("SB200_XSPLIT0_MB"))
        {
            const int64_t w0 = (int64_t) (std::atof(e) * 1024.0 * 1024.0 / 8.0);
            if (w0 >= 1024 && w0 < A.n - 1024)
                W0 = w0 & ~int64_t(15);
        }
    ColMap map{0, W, 0, 0};
    map.W0 = W0;
    split_by_map(A, nb, map, stream);
    A.col_block_width = W0 > 0 ? W0 : W;
}

void split_column_chunks(DeviceCsr& A, int nchunks, int64_t slab, int nranks, cudaStream_t stream)
{
    SB200_REQUIRE(nchunks >= 1 && nchunks <= kMaxColBlocks && slab >= 1 && nranks >= 1, SB200_LOGIC, "bad chunk layout");
    const int64_t len = round_up((slab + nchunks - 1) / nchunks, 16);
    SB200_REQUIRE((int64_t) nranks * len < (1LL << 31), SB200_LOGIC, "chunk ids exceed the int32 range");
    A.chunk_slab = slab;
    A.chunk_len = len;
    A.chunk_ranks = nranks;
    if (A.nrows == 0 || A.nnz == 0)
    {
        // keep nchunks (empty) blocks so that every rank issues the same sequence of collectives and launches
        A.blocks.clear();
        A.blocks.resize(nchunks);
        for (CsrBlock& B : A.blocks)
        {
            B.rowptr.alloc(A.nrows + 1);
            B.rowptr.zero(stream);
            B.col.alloc(1);
            B.val.alloc(1);
        }
        A.rowptr.release();
        A.col.release();
        A.val.release();
        SB200_CUDA_CHECK(cudaStreamSynchronize(stream));
        return;
    }
    split_by_map(A, nchunks, ColMap{1, 0, slab, len}, stream);
}

void launch_permute_to_chunks(const DeviceCsr& A, const double* x_nat, double* x_chunked, cudaStream_t stream)
{
    const int nb = (int) A.blocks.size();
    const int64_t total = A.chunk_stride() * nb;
    permute_to_chunks_kernel<<<grid_for(total, 256), 256, 0, stream>>>(x_nat, x_chunked, A.n, A.chunk_slab, A.chunk_len, A.chunk_ranks, nb);
    SB200_CUDA_CHECK(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------------
// Sliced layout (SellBlock, kernels.h): window sort -> slice widths -> scan -> step-major fill
// ---------------------------------------------------------------------------------------------
namespace {

// One CTA per window of kSellWindow rows, one thread per row: stable rank of the row by descending length.
// perm[win * W + rank] = row inside the window; cnt[slice] = 32 * (longest row of the slice) = stored entries of the slice.
__global__ void __launch_bounds__(kSellWindow) sell_sort_kernel(const int* __restrict__ rowptr, int64_t nrows, unsigned short* __restrict__ perm, int* __restrict__ cnt)
{
    __shared__ __align__(16) int s_len[kSellWindow];
    const int r = threadIdx.x;
    const int64_t row = (int64_t) blockIdx.x * kSellWindow + r;
    const int len = (row < nrows) ? rowptr[row + 1] - rowptr[row] : 0;
    s_len[r] = len;
    __syncthreads();
    int rank = 0;
    const int4* s4 = reinterpret_cast<const int4*>(s_len);
    for (int q = 0; q < kSellWindow / 4; q++)
    {
        const int4 l = s4[q];  // broadcast read
        const int j = 4 * q;
        rank += (l.x > len || (l.x == len && j + 0 < r)) ? 1 : 0;
        rank += (l.y > len || (l.y == len && j + 1 < r)) ? 1 : 0;
        rank += (l.z > len || (l.z == len && j + 2 < r)) ? 1 : 0;
        rank += (l.w > len || (l.w == len && j + 3 < r)) ? 1 : 0;
    }
    perm[(int64_t) blockIdx.x * kSellWindow + rank] = (unsigned short) r;
    if ((rank & (kSellSlice - 1)) == 0)
        cnt[(int64_t) blockIdx.x * (kSellWindow / kSellSlice) + rank / kSellSlice] = (len < (1 << 25)) ? len * kSellSlice : 0x7fffffe0;  // absurd rows: forces rejection
}

// One warp per slice: lane = row of the slice; writes the slice step-major, padding with (col -1, val 0).
__global__ void sell_fill_kernel(const int* __restrict__ rowptr, const int* __restrict__ col, const double* __restrict__ val, int64_t nrows, int64_t nslices,
                                 const unsigned short* __restrict__ perm, const int* __restrict__ slice_ptr, int* __restrict__ scol, double* __restrict__ sval)
{
    const int lane = threadIdx.x & 31;
    const int64_t gw = ((int64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t) gridDim.x * blockDim.x) >> 5;
    constexpr int SPW = kSellWindow / kSellSlice;  // slices per window
    for (int64_t s = gw; s < nslices; s += nw)
    {
        const int64_t win = s / SPW;
        const int pos = (int) (s % SPW) * kSellSlice + lane;
        const int64_t row = win * kSellWindow + perm[win * kSellWindow + pos];
        int b = 0, len = 0;
        if (row < nrows)
        {
            b = rowptr[row];
            len = rowptr[row + 1] - b;
        }
        const int base = slice_ptr[s];
        const int width = (slice_ptr[s + 1] - base) / kSellSlice;
        for (int t = 0; t < width; t++)
        {
            const int dst = base + t * kSellSlice + lane;
            const bool real = t < len;
            scol[dst] = real ? col[b + t] : -1;
            sval[dst] = real ? val[b + t] : 0.0;
        }
    }
}

// Stored entries of the sliced layout of one CSR (device rowptr), without building it.
void sell_prepare(const int* rowptr, int64_t nrows, SellBlock& S, long long* padded, cudaStream_t stream)
{
    const int64_t nwin = (nrows + kSellWindow - 1) / kSellWindow;
    const int64_t nslices = nwin * (kSellWindow / kSellSlice);
    S.nwin = nwin;
    S.perm.alloc((size_t) (nwin * kSellWindow));
    S.slice_ptr.alloc((size_t) (nslices + 1));
    DevBuf<int> cnt((size_t) nslices);
    sell_sort_kernel<<<(unsigned) nwin, kSellWindow, 0, stream>>>(rowptr, nrows, S.perm.get(), cnt.get());
    SB200_CUDA_CHECK(cudaGetLastError());
    exclusive_scan_to_rowptr(cnt.get(), nslices, S.slice_ptr.get(), padded, stream);
    S.padded = *padded;
}

void sell_fill(const int* rowptr, const int* col, const double* val, int64_t nrows, SellBlock& S, cudaStream_t stream)
{
    const int64_t nslices = S.nwin * (kSellWindow / kSellSlice);
    S.col.alloc((size_t) std::max<int64_t>(S.padded, 1));
    S.val.alloc((size_t) std::max<int64_t>(S.padded, 1));
    sell_fill_kernel<<<grid_for(nslices * 32, 256), 256, 0, stream>>>(rowptr, col, val, nrows, nslices, S.perm.get(), S.slice_ptr.get(), S.col.get(), S.val.get());
    SB200_CUDA_CHECK(cudaGetLastError());
}

}  // namespace

bool build_sell_layout(DeviceCsr& A, double max_fill, cudaStream_t stream)
{
    if (A.nrows == 0)
        return false;
    struct Part
    {
        const int* rowptr;
        const int* col;
        const double* val;
        int64_t nnz;
        SellBlock* S;
    };
    std::vector<Part> parts;
    if (A.blocks.empty())
        parts.push_back({A.rowptr.get(), A.col.get(), A.val.get(), A.nnz, &A.sell});
    else
        for (CsrBlock& B : A.blocks)
            parts.push_back({B.rowptr.get(), B.col.get(), B.val.get(), B.nnz, &B.sell});
    bool ok = true;
    long long total_padded = 0, total_nnz = 0;
    for (Part& p : parts)
    {
        long long padded = 0;
        sell_prepare(p.rowptr, A.nrows, *p.S, &padded, stream);
        ok = ok && padded < (1LL << 31) - 64;
        total_padded += padded;
        total_nnz += p.nnz;
    }
    // padding of at most one 32-row step per window is always accepted (tiny operands)
    ok = ok && double(total_padded) <= max_fill * double(total_nnz) + 32.0 * double(kSellWindow) * double(parts.size());
    if (!ok)
    {
        for (Part& p : parts)
            p.S->release();
        return false;
    }
    for (Part& p : parts)
        sell_fill(p.rowptr, p.col, p.val, A.nrows, *p.S, stream);
    SB200_CUDA_CHECK(cudaStreamSynchronize(stream));
    return true;
}

void upload_csr_slab(int64_t n, int64_t row0, int64_t nrows, const int64_t* rowptr_local, const int32_t* col, const double* values, cudaStream_t stream,
                     DeviceCsr& out)
{
    SB200_REQUIRE(n >= 1 && n < (1LL << 31), SB200_INVALID_ARGUMENT, "matrix order out of range");
    SB200_REQUIRE(row0 >= 0 && nrows >= 0 && row0 + nrows <= n, SB200_INVALID_ARGUMENT, "row slab out of range");
    const int64_t base = rowptr_local[0];
    const int64_t nnz = rowptr_local[nrows] - base;
    SB200_REQUIRE(nnz < (1LL << 31), SB200_INVALID_ARGUMENT, "local nnz exceeds the int32 index range");
    out.n = n;
    out.row0 = row0;
    out.nrows = nrows;
    out.nnz = nnz;
    out.rowptr.alloc(nrows + 1);
    out.col.alloc(std::max<int64_t>(nnz, 1));
    out.val.alloc(std::max<int64_t>(nnz, 1));
    DevBuf<long long> tmp(nrows + 1);
    static_assert(sizeof(long long) == sizeof(int64_t), "int64 layout");
    SB200_CUDA_CHECK(cudaMemcpyAsync(tmp.get(), rowptr_local, sizeof(int64_t) * (nrows + 1), cudaMemcpyHostToDevice, stream));
    narrow_rowptr_kernel<<<grid_for(nrows + 1, 256), 256, 0, stream>>>(tmp.get(), base, out.rowptr.get(), nrows + 1);
    SB200_CUDA_CHECK(cudaGetLastError());
    if (nnz > 0)
    {
        SB200_CUDA_CHECK(cudaMemcpyAsync(out.col.get(), col + base, sizeof(int) * nnz, cudaMemcpyHostToDevice, stream));
        SB200_CUDA_CHECK(cudaMemcpyAsync(out.val.get(), values + base, sizeof(double) * nnz, cudaMemcpyHostToDevice, stream));
    }
    SB200_CUDA_CHECK(cudaStreamSynchronize(stream));
}

}  // namespace sb200

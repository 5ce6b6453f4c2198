// CSR SpMV kernels for sm_100a (K1 / K-A of SURVEY.md §2.1).
//
// Replaces SparseSymMatProd::perform_op (MatOp/SparseSymMatProd.h:83-88) and
// SparseGenMatProd::perform_op (MatOp/SparseGenMatProd.h:82-87); the fused variant also covers
// the head of one Lanczos/Arnoldi step: v_i = f/beta (Lanczos.h:106, Arnoldi.h:236),
// w -= H(i,i-1) v_{i-1} (Lanczos.h:139) and the partial <v_i, w> (Lanczos.h:142).
//
// Layout: full CSR in HBM, int32 row pointers / column ids, fp64 values, columns ascending in a
// row.  A sub-warp of L lanes owns one row: the L lanes read consecutive (col, val) pairs, so a
// warp reads one contiguous slab of the CSR arrays per iteration (coalesced), gathers x through
// the read-only path and combines with log2(L) shuffles.  One warp owns 32 consecutive rows and
// routes the row sums to lane (row - row0) so that the epilogue is coalesced.  CSR streams are
// read with L1::no_allocate + L2 evict_first, the gathered vector x with evict_last.
//
// Column blocking: when x (8n bytes) exceeds the L2 budget the operand is stored as several CSR
// sub-matrices over column ranges (csr_build.cu: split_column_blocks) and one launch per block
// accumulates into y, so every pass gathers from an L2-resident slice of x instead of DRAM
// (measured at n = 1e7 unblocked: 7.4 GB of DRAM reads per SpMV for 2.76 GB of algorithmic bytes).
// Algorithmic bytes per row at d nnz/row: 12 d + 4 + 8 (x once) + 8 (y)  (SURVEY §8d); blocking adds
// 4 (row pointer) + 16 (y read-modify-write) bytes per row per extra block.
//
// Sliced layout (SellBlock, kernels.h; SB200_SPMV_FORMAT=sell, experimental): the same column blocks stored as 32-row slices,
// step-major, rows sorted by length inside 1024-row windows.  One lane owns one row: no shuffles, every CSR byte arrives in a
// fully coalesced warp load, the row sum runs in ascending column order.  A CTA owns one window at a time and returns the sums
// to natural row order through 8 KB of shared memory, so the epilogues (accumulate, fused step head) are the coalesced ones of
// the CSR-vector kernels.  Traffic per row at d nnz/row and padding p: 12 d (1 + p) + 2 (perm) + 0.25 (slice_ptr) + x + y.
#include <cstdlib>
#include <cstring>

#include "kernels.h"

namespace sb200 {

namespace {

constexpr int kSpmvBlock = 256;

// All 32 lanes of a warp must call this together (sub-warp shuffle with a full mask); a sub-warp whose
// row is out of range passes start == end.
template <int L>
__device__ __forceinline__ double row_dot(int start, int end, const int* __restrict__ col, const double* __restrict__ val, const double* __restrict__ x,
                                          int lane, uint64_t pol_stream, uint64_t pol_keep)
{
    double acc = 0.0;
    int p = start + lane;
    // two independent gathers in flight per lane
    for (; p + L < end; p += 2 * L)
    {
        const int c0 = ld_stream_s32(col + p, pol_stream);
        const int c1 = ld_stream_s32(col + p + L, pol_stream);
        const double v0 = ld_stream_f64(val + p, pol_stream);
        const double v1 = ld_stream_f64(val + p + L, pol_stream);
        const double x0 = ld_keep_f64(x + c0, pol_keep);
        const double x1 = ld_keep_f64(x + c1, pol_keep);
        acc = fma(v0, x0, acc);
        acc = fma(v1, x1, acc);
    }
    if (p < end)
    {
        const int c0 = ld_stream_s32(col + p, pol_stream);
        const double v0 = ld_stream_f64(val + p, pol_stream);
        acc = fma(v0, ld_keep_f64(x + c0, pol_keep), acc);
    }
    return subwarp_sum<L>(acc);
}

// One warp owns 32 consecutive rows: L iterations of 32/L rows.  The row sums are routed by one
// shuffle per iteration so that lane j ends up with the sum of row (row0 + j); the epilogue (stores,
// the fused Lanczos head) then runs with all 32 lanes on consecutive rows, i.e. coalesced and without
// the 4-of-32-lane divergence a per-sub-warp epilogue would have.
template <int L>
__device__ __forceinline__ double warp_rows_dot(const int* __restrict__ rowptr, const int* __restrict__ col, const double* __restrict__ val,
                                                const double* __restrict__ x, int64_t row0, int64_t nrows, int wlane, uint64_t pol_stream, uint64_t pol_keep)
{
    constexpr int RPW = 32 / L;  // rows per warp iteration
    const int lane = wlane % L, sub = wlane / L;
    // row pointers of the 32-row block: one coalesced load + one extra element
    const int64_t rmine = row0 + wlane;
    const int rp = __ldg(rowptr + (rmine <= nrows ? rmine : nrows));
    const int64_t rlast = row0 + 32;
    const int rp_hi = __ldg(rowptr + (rlast <= nrows ? rlast : nrows));
    double mine = 0.0;
#pragma unroll
    for (int it = 0; it < L; it++)
    {
        const int j = it * RPW + sub;  // row index inside the block handled by this sub-warp
        const int start = __shfl_sync(0xffffffffu, rp, j);
        const int endn = __shfl_sync(0xffffffffu, rp, (j + 1) & 31);
        const int end = (j == 31) ? rp_hi : endn;
        const double sres = row_dot<L>(start, end, col, val, x, lane, pol_stream, pol_keep);
        const double t = __shfl_sync(0xffffffffu, sres, (wlane % RPW) * L);
        if (wlane / RPW == it)
            mine = t;
    }
    return mine;
}

// y = (ACCUM ? y : 0) + A_block x
template <int L, bool ACCUM>
__global__ void __launch_bounds__(kSpmvBlock, 8) spmv_plain_kernel(const int* __restrict__ rowptr, const int* __restrict__ col, const double* __restrict__ val,
                                                                const double* __restrict__ x, double* y, int64_t nrows, const int* abort)
{
    if (abort != nullptr && *abort != 0)
        return;
    const uint64_t pol_stream = l2_policy_evict_first();
    const uint64_t pol_keep = l2_policy_evict_last();
    const int wlane = threadIdx.x & 31;
    const int64_t warp = (int64_t) blockIdx.x * (kSpmvBlock / 32) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t) gridDim.x * (kSpmvBlock / 32);
    for (int64_t row0 = warp * 32; row0 < nrows; row0 += nwarps * 32)
    {
        double s = warp_rows_dot<L>(rowptr, col, val, x, row0, nrows, wlane, pol_stream, pol_keep);
        const int64_t row = row0 + wlane;
        if (row < nrows)
        {
            if (ACCUM)
                s += y[row];
            y[row] = s;
        }
    }
}

// Fused step head on the LAST column block.  x_full: un-normalised residual (all n entries), f_loc: this
// rank's rows of it.  ACCUM: w already holds the partial product of the previous column blocks.
// Resident CTAs per SM requested for the fused kernel: 8 (32 registers, a few spilled loop invariants) keeps the gather
// rate of the plain kernel; SB200_STEP_MINBLOCKS is a build-time knob for A/B runs.
#ifndef SB200_STEP_MINBLOCKS
#define SB200_STEP_MINBLOCKS 8
#endif
template <int L, bool SYM, bool ACCUM>
__global__ void __launch_bounds__(kSpmvBlock, SB200_STEP_MINBLOCKS)
    spmv_step_kernel(const int* __restrict__ rowptr, const int* __restrict__ col, const double* __restrict__ val, const double* __restrict__ x_full,
                     const double* __restrict__ f_loc, double* __restrict__ V, int64_t ldv, double* w, int64_t nrows, FacCtl* ctl, double* H, int m, int i,
                     int restarted, double* partials, unsigned int* ticket)
{
    const uint64_t pol_stream = l2_policy_evict_first();
    const uint64_t pol_keep = l2_policy_evict_last();
    const int wlane = threadIdx.x & 31;
    const int64_t warp = (int64_t) blockIdx.x * (kSpmvBlock / 32) + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t) gridDim.x * (kSpmvBlock / 32);
    if (ctl->abort != 0)
        return;  // sweep mode: an earlier step handed control back to the host
    const double beta = ctl->beta;
    const double hsub = restarted ? 0.0 : beta;
    double* __restrict__ vi = V + (int64_t) i * ldv;
    const double* __restrict__ vp = V + (int64_t) (i - 1) * ldv;

    double part = 0.0;
    for (int64_t row0 = warp * 32; row0 < nrows; row0 += nwarps * 32)
    {
        double s = warp_rows_dot<L>(rowptr, col, val, x_full, row0, nrows, wlane, pol_stream, pol_keep);
        const int64_t row = row0 + wlane;
        if (row < nrows)
        {
            if (ACCUM)
                s += w[row];
            const double v = f_loc[row] / beta;  // v_i = f / ||f||      (Lanczos.h:106)
            vi[row] = v;
            double wr = s / beta;                // w = A v_i, with the scaling applied after the product
            if (SYM)
            {
                wr -= hsub * vp[row];            // w -= H(i,i-1) v_{i-1}  (Lanczos.h:139)
                part = fma(v, wr, part);         // <v_i, w>               (Lanczos.h:142)
            }
            w[row] = wr;
        }
    }

    if (blockIdx.x == 0 && threadIdx.x == 0)
    {
        // step bookkeeping (Lanczos.h:127-128, Arnoldi.h:239)
        ctl->i = i;
        ctl->count = 0;
        ctl->hsub = hsub;
        ctl->need_corr = 0;
        ctl->f_zeroed = 0;
        ctl->dgks_skip = 0;
        H[i + (int64_t) (i - 1) * m] = hsub;
        if (SYM)
            H[(i - 1) + (int64_t) i * m] = hsub;
    }

    if (SYM)
    {
        // CTA partial in a fixed order: warp shuffle tree, then warps 0..7 sequentially
        __shared__ double s_w[kSpmvBlock / 32];
        part = warp_sum(part);
        if ((threadIdx.x & 31) == 0)
            s_w[threadIdx.x >> 5] = part;
        __syncthreads();
        double cta = 0.0;
        if (threadIdx.x == 0)
        {
#pragma unroll
            for (int q = 0; q < kSpmvBlock / 32; q++)
                cta += s_w[q];
        }
        grid_reduce_fixed_order<kSpmvBlock>(cta, 1, partials, ticket, ctl->red_a);
    }
}

// ---------------------------------------------------------------------------------------------
// Sliced-layout kernels: one lane per row, one CTA per 1024-row window at a time
// ---------------------------------------------------------------------------------------------
constexpr int sell_min_blocks(int threads) { return 1536 / threads; }  // 48 resident warps per SM (32 at 1024 threads)

// Row sums of window `win` into s_y (natural row order inside the window).  The slices of a window are sorted by length (longest
// first); warp w takes them in serpentine order -- w, 2 WARPS - 1 - w, 2 WARPS + w, ... -- so that every warp of the CTA gets about
// the same number of steps before the barrier (w, w + WARPS, ... would give warp 0 the longest slice of every group).
template <int THREADS, int UN = 4>
__device__ __forceinline__ void sell_window_dot(const int* __restrict__ slice_ptr, const int* __restrict__ scol, const double* __restrict__ sval,
                                                const unsigned short* __restrict__ perm, const double* __restrict__ x, int64_t win, double* s_y,
                                                uint64_t pol_stream, uint64_t pol_keep)
{
    constexpr int WARPS = THREADS / 32;
    constexpr int SPW = kSellWindow / kSellSlice;
    static_assert(SPW % WARPS == 0, "warps must divide the slices of a window");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll 1
    for (int q = 0; q < SPW / WARPS; q++)
    {
        const int sl = (q & 1) ? (q + 1) * WARPS - 1 - warp : q * WARPS + warp;
        const int64_t s = win * SPW + sl;
        const int base = __ldg(slice_ptr + s);
        const int steps = (__ldg(slice_ptr + s + 1) - base) >> 5;
        const int* cp = scol + base + lane;
        const double* vp = sval + base + lane;
        double acc = 0.0;
        int t = 0;
        // UN steps (= UN independent gathers per lane) in flight; padding carries col -1 / val 0
        for (; t + UN <= steps; t += UN)
        {
            int c[UN];
            double v[UN], xv[UN];
#pragma unroll
            for (int u = 0; u < UN; u++)
                c[u] = ld_stream_s32(cp + (t + u) * 32, pol_stream);
#pragma unroll
            for (int u = 0; u < UN; u++)
                v[u] = ld_stream_f64(vp + (t + u) * 32, pol_stream);
#pragma unroll
            for (int u = 0; u < UN; u++)
                xv[u] = (c[u] >= 0) ? ld_keep_f64(x + c[u], pol_keep) : 0.0;
#pragma unroll
            for (int u = 0; u < UN; u++)
                acc = fma(v[u], xv[u], acc);
        }
        for (; t < steps; t++)
        {
            const int c0 = ld_stream_s32(cp + t * 32, pol_stream);
            const double v0 = ld_stream_f64(vp + t * 32, pol_stream);
            const double x0 = (c0 >= 0) ? ld_keep_f64(x + c0, pol_keep) : 0.0;
            acc = fma(v0, x0, acc);
        }
        s_y[__ldg(perm + win * kSellWindow + sl * kSellSlice + lane)] = acc;
    }
}

// y = (ACCUM ? y : 0) + A_block x
template <int THREADS, bool ACCUM>
__global__ void __launch_bounds__(THREADS, sell_min_blocks(THREADS))
    sell_plain_kernel(const int* __restrict__ slice_ptr, const int* __restrict__ scol, const double* __restrict__ sval, const unsigned short* __restrict__ perm,
                      const double* __restrict__ x, double* y, int64_t nrows, int64_t nwin, const int* abort)
{
    __shared__ double s_y[kSellWindow];
    if (abort != nullptr && *abort != 0)
        return;
    const uint64_t pol_stream = l2_policy_evict_first();
    const uint64_t pol_keep = l2_policy_evict_last();
    for (int64_t win = blockIdx.x; win < nwin; win += gridDim.x)
    {
        sell_window_dot<THREADS>(slice_ptr, scol, sval, perm, x, win, s_y, pol_stream, pol_keep);
        __syncthreads();
        for (int r = threadIdx.x; r < kSellWindow; r += THREADS)
        {
            const int64_t row = win * kSellWindow + r;
            if (row < nrows)
            {
                double s = s_y[r];
                if (ACCUM)
                    s += y[row];
                y[row] = s;
            }
        }
        __syncthreads();
    }
}

// Fused step head on the last column block; same contract as spmv_step_kernel.
template <int THREADS, bool SYM, bool ACCUM>
__global__ void __launch_bounds__(THREADS, sell_min_blocks(THREADS))
    sell_step_kernel(const int* __restrict__ slice_ptr, const int* __restrict__ scol, const double* __restrict__ sval, const unsigned short* __restrict__ perm,
                     const double* __restrict__ x_full, const double* __restrict__ f_loc, double* __restrict__ V, int64_t ldv, double* w, int64_t nrows,
                     int64_t nwin, FacCtl* ctl, double* H, int m, int i, int restarted, double* partials, unsigned int* ticket)
{
    __shared__ double s_y[kSellWindow];
    const uint64_t pol_stream = l2_policy_evict_first();
    const uint64_t pol_keep = l2_policy_evict_last();
    if (ctl->abort != 0)
        return;  // sweep mode: an earlier step handed control back to the host
    const double beta = ctl->beta;
    const double hsub = restarted ? 0.0 : beta;
    double* __restrict__ vi = V + (int64_t) i * ldv;
    const double* __restrict__ vp = V + (int64_t) (i - 1) * ldv;

    double part = 0.0;
    for (int64_t win = blockIdx.x; win < nwin; win += gridDim.x)
    {
        sell_window_dot<THREADS>(slice_ptr, scol, sval, perm, x_full, win, s_y, pol_stream, pol_keep);
        __syncthreads();
        for (int r = threadIdx.x; r < kSellWindow; r += THREADS)
        {
            const int64_t row = win * kSellWindow + r;
            if (row < nrows)
            {
                double s = s_y[r];
                if (ACCUM)
                    s += w[row];
                const double v = f_loc[row] / beta;  // v_i = f / ||f||      (Lanczos.h:106)
                vi[row] = v;
                double wr = s / beta;                // w = A v_i, with the scaling applied after the product
                if (SYM)
                {
                    wr -= hsub * vp[row];            // w -= H(i,i-1) v_{i-1}  (Lanczos.h:139)
                    part = fma(v, wr, part);         // <v_i, w>               (Lanczos.h:142)
                }
                w[row] = wr;
            }
        }
        __syncthreads();
    }

    if (blockIdx.x == 0 && threadIdx.x == 0)
    {
        // step bookkeeping (Lanczos.h:127-128, Arnoldi.h:239)
        ctl->i = i;
        ctl->count = 0;
        ctl->hsub = hsub;
        ctl->need_corr = 0;
        ctl->f_zeroed = 0;
        ctl->dgks_skip = 0;
        H[i + (int64_t) (i - 1) * m] = hsub;
        if (SYM)
            H[(i - 1) + (int64_t) i * m] = hsub;
    }

    if (SYM)
    {
        // CTA partial in a fixed order: warp shuffle tree, then the warps sequentially
        __shared__ double s_w[THREADS / 32];
        part = warp_sum(part);
        if ((threadIdx.x & 31) == 0)
            s_w[threadIdx.x >> 5] = part;
        __syncthreads();
        double cta = 0.0;
        if (threadIdx.x == 0)
        {
#pragma unroll
            for (int q = 0; q < THREADS / 32; q++)
                cta += s_w[q];
        }
        grid_reduce_fixed_order<THREADS>(cta, 1, partials, ticket, ctl->red_a);
    }
}

// ---------------------------------------------------------------------------------------------
// Fused step head + first panel pass ("K-A+B"): the last column block of the operator application, the step head
//   v_i = f/beta (Lanczos.h:106), w = A v_i - H(i,i-1) v_{i-1} (:131-139)
// and the Gram-Schmidt coefficients  c = V[:, :i+1]^T w  (the adjoint_product of Lanczos.h:152 / Arnoldi.h:251) in ONE kernel.
// Why: on uniformly random columns the gather phase is bound by the L1TEX wavefront / L2 sector rate and leaves HBM more than half
// idle, while the panel pass is a pure HBM stream.  A CTA alternates, window by window, between the two (gather the window's row sums ->
// step head -> stream the window's 1024 x i tile of V against the fresh w values held in shared memory), and the two resident CTAs of
// an SM are in different phases most of the time, so the V stream fills the memory pipe the gathers cannot use.
// Phase C is fed by TMA: warp q owns the columns k = q (mod 16) of the panel and walks them in tasks of 256 rows (2 KB, contiguous in
// the column-major basis); one lane brings each task in with a 1-D bulk copy (cp.async.bulk + mbarrier complete_tx, L2 evict_first) into
// the warp's private ring of kFusedSlots slots in shared memory, so the bytes in flight per SM (2 CTAs x 16 warps x 3 x 2 KB = 192 KB)
// do not cost registers; the lanes then read 16-byte row pairs from the slot and from the w values of the window (conflict free) and
// keep one accumulator per owned column across all windows of the CTA.
// The coefficient c_i = <v_i, w> is formed from the registers of the step head (column i of V is being written by this kernel).
// Reduction: warp shuffle tree, fixed order inside the CTA, grid_reduce_fixed_order across CTAs (bit-reproducible).
// Output: red[0..i] = V[:, :i+1]^T w.  Algorithmic bytes: those of the SpMV block + 24 B/row (step head) + 8 * nrows * i (V).
// ---------------------------------------------------------------------------------------------
// Configurations of the fused kernel (A/B knob SB200_FUSED_CFG = "<threads>x<NB>[u<UN>]"; two CTAs are resident per SM in every one):
//   THREADS  CTA size: 512 threads at <= 64 registers, or 256 threads at <= 128 registers;
//   NB       16-byte loads of V in flight per lane in phase C (the bytes in flight per SM are 2 CTAs x THREADS x NB x 16 B: 64 KB at
//            512x4, 128 KB at 256x16 -- what a pure register pipeline can afford; rings in shared memory fed by cp.async.bulk (TMA) or
//            cp.async were measured and are SLOWER: 96 KB of ring per CTA leave the L1 ~10 KB, which throttles the gather phase
//            (357 vs 477 SpMV-iters/s at n = 1e7, profiles/r2d_quick_*_n1e7.log));
//   UN       row-sum steps (gathers per lane) in flight in phase A.
template <bool SYM, bool ACCUM, int THREADS, int NB, int UN, int RES = 2>
__global__ void __launch_bounds__(THREADS, RES)
    sell_step_dot_kernel(const int* __restrict__ slice_ptr, const int* __restrict__ scol, const double* __restrict__ sval, const unsigned short* __restrict__ perm,
                         const double* __restrict__ x_full, const double* __restrict__ f_loc, double* V, int64_t ldv, double* w, int64_t nrows, int64_t nwin,
                         FacCtl* ctl, double* H, int m, int i, int restarted, double* red_out, double* partials, unsigned int* ticket)
{
    constexpr int WARPS = THREADS / 32;
    constexpr int CPW = kPanelMaxCols / WARPS;      // owned columns per warp (4 or 8)
    constexpr int LPC = kSellWindow / 64;           // 16-byte loads per lane and column of a window (16)
    static_assert(LPC % NB == 0, "NB must divide 16");
    __shared__ __align__(16) double s_y[kSellWindow];
    __shared__ double s_col[kPanelMaxCols];
    __shared__ double s_w[WARPS];
    __shared__ double s_alpha;
    const uint64_t pol_stream = l2_policy_evict_first();
    const uint64_t pol_keep = l2_policy_evict_last();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (ctl->abort != 0)
        return;  // sweep mode: an earlier step handed control back to the host
    const double beta = ctl->beta;
    const double hsub = restarted ? 0.0 : beta;
    double* vi = V + (int64_t) i * ldv;
    const double* vp = V + (int64_t) (i - 1) * ldv;
    const int ncols_mine = (i > warp) ? (i - warp + WARPS - 1) / WARPS : 0;  // columns warp, warp + WARPS, ... below i

    double acc[CPW];
#pragma unroll
    for (int q = 0; q < CPW; q++)
        acc[q] = 0.0;
    double part = 0.0;

    for (int64_t win = blockIdx.x; win < nwin; win += gridDim.x)
    {
        const int64_t wrow0 = win * kSellWindow;
        // ---- phase A: row sums of the window (gather bound) ----
        sell_window_dot<THREADS, UN>(slice_ptr, scol, sval, perm, x_full, win, s_y, pol_stream, pol_keep);
        __syncthreads();
        // ---- phase B: step head on natural-order rows; the finished w values replace the row sums in shared memory ----
        for (int r = threadIdx.x; r < kSellWindow; r += THREADS)
        {
            const int64_t row = wrow0 + r;
            double wr = 0.0;
            if (row < nrows)
            {
                double sum = s_y[r];
                if (ACCUM)
                    sum += w[row];
                const double v = f_loc[row] / beta;  // v_i = f / ||f||      (Lanczos.h:106)
                vi[row] = v;
                wr = sum / beta;                     // w = A v_i, with the scaling applied after the product
                if (SYM)
                    wr -= hsub * vp[row];            // w -= H(i,i-1) v_{i-1}  (Lanczos.h:139)
                part = fma(v, wr, part);             // <v_i, w>               (Lanczos.h:142, Arnoldi.h:251 entry i)
                w[row] = wr;
            }
            s_y[r] = wr;
        }
        __syncthreads();
        // ---- phase C: c_k += V[window rows, k]^T w for the owned columns k < i (HBM stream) ----
        const double2* wsm = reinterpret_cast<const double2*>(s_y) + lane;
        const int64_t rlane = wrow0 + lane * 2;
#pragma unroll
        for (int cl = 0; cl < CPW; cl++)
        {
            if (cl < ncols_mine)
            {
                const double* colp = V + (int64_t) (warp + WARPS * cl) * ldv + rlane;
                double sacc = 0.0;
#pragma unroll
                for (int b = 0; b < LPC; b += NB)
                {
                    double2 a[NB];
#pragma unroll
                    for (int u = 0; u < NB; u++)
                        a[u] = (rlane + (b + u) * 64 < ldv) ? ld_stream_f64x2(colp + (b + u) * 64, pol_stream) : make_double2(0.0, 0.0);
#pragma unroll
                    for (int u = 0; u < NB; u++)
                    {
                        const double2 ww = wsm[(b + u) * 32];
                        sacc = fma(a[u].x, ww.x, sacc);
                        sacc = fma(a[u].y, ww.y, sacc);
                    }
                }
                acc[cl] += sacc;
            }
        }
        __syncthreads();  // s_y is overwritten by the next window
    }

    if (blockIdx.x == 0 && threadIdx.x == 0)
    {
        // step bookkeeping (Lanczos.h:127-128, Arnoldi.h:239)
        ctl->i = i;
        ctl->count = 0;
        ctl->hsub = hsub;
        ctl->need_corr = 0;
        ctl->f_zeroed = 0;
        ctl->dgks_skip = 0;
        H[i + (int64_t) (i - 1) * m] = hsub;
        if (SYM)
            H[(i - 1) + (int64_t) i * m] = hsub;
    }

    // ---- CTA-level combine in a fixed order ----
#pragma unroll
    for (int q = 0; q < CPW; q++)
    {
        const double sum = warp_sum(acc[q]);
        if (lane == 0)
            s_col[warp + WARPS * q] = sum;
    }
    part = warp_sum(part);
    if (lane == 0)
        s_w[warp] = part;
    __syncthreads();
    if (threadIdx.x == 0)
    {
        double a = 0.0;
#pragma unroll
        for (int q = 0; q < WARPS; q++)
            a += s_w[q];
        s_alpha = a;
    }
    __syncthreads();
    double cta = 0.0;
    if ((int) threadIdx.x < i)
        cta = s_col[threadIdx.x];
    else if ((int) threadIdx.x == i)
        cta = s_alpha;
    grid_reduce_fixed_order<THREADS>(cta, i + 1, partials, ticket, red_out);
}

// ---------------------------------------------------------------------------------------------
// Complex Hermitian operand (SparseHermMatProd::perform_op, MatOp/SparseHermMatProd.h:83-88): the full CSR carries
// interleaved complex values; x and y are interleaved complex vectors.  Sub-warp of L lanes per row as in the real kernel;
// one 16-byte gather per entry (a complex operand entry is half a sector, so the gather is twice as sector-efficient as the
// real one).  Algorithmic bytes per row at d nnz/row: 20 d + 4 + 16 (x) + 16 (y).
// ---------------------------------------------------------------------------------------------
template <int L>
__global__ void __launch_bounds__(kSpmvBlock, 4) spmv_z_kernel(const int* __restrict__ rowptr, const int* __restrict__ col, const double2* __restrict__ val,
                                                                const double* __restrict__ x, double* __restrict__ y, int64_t nrows)
{
    const uint64_t pol_stream = l2_policy_evict_first();
    const uint64_t pol_keep = l2_policy_evict_last();
    const int lane = threadIdx.x % L;
    const int64_t sub = ((int64_t) blockIdx.x * kSpmvBlock + threadIdx.x) / L, nsub = (int64_t) gridDim.x * kSpmvBlock / L;
    // every lane of a warp runs the same number of rounds: subwarp_sum shuffles with the full mask
    const int64_t rounds = (nrows + nsub - 1) / nsub;
    for (int64_t it = 0; it < rounds; it++)
    {
        const int64_t row = sub + it * nsub;
        int start = 0, end = 0;
        if (row < nrows)
        {
            start = __ldg(rowptr + row);
            end = __ldg(rowptr + row + 1);
        }
        double ar = 0.0, ai = 0.0;
        for (int p = start + lane; p < end; p += L)
        {
            const int c = ld_stream_s32(col + p, pol_stream);
            const double2 a = ld_stream_f64x2(reinterpret_cast<const double*>(val + p), pol_stream);
            const double2 xv = ld_keep_f64x2(x + 2 * (int64_t) c, pol_keep);
            ar = fma(a.x, xv.x, ar);
            ar = fma(-a.y, xv.y, ar);
            ai = fma(a.x, xv.y, ai);
            ai = fma(a.y, xv.x, ai);
        }
        ar = subwarp_sum<L>(ar);
        ai = subwarp_sum<L>(ai);
        if (lane == 0 && row < nrows)
            *reinterpret_cast<double2*>(y + 2 * row) = make_double2(ar, ai);
    }
}

struct BlockView
{
    const int* rowptr;
    const int* col;
    const double* val;
    int64_t nnz;
    const SellBlock* sell;  // non-null when the sliced layout was built
};

int nblocks_of(const DeviceCsr& A) { return A.blocks.empty() ? 1 : (int) A.blocks.size(); }

BlockView view_of(const DeviceCsr& A, int c)
{
    if (A.blocks.empty())
        return {A.rowptr.get(), A.col.get(), A.val.get(), A.nnz, A.sell.built() ? &A.sell : nullptr};
    const CsrBlock& B = A.blocks[c];
    return {B.rowptr.get(), B.col.get(), B.val.get(), B.nnz, B.sell.built() ? &B.sell : nullptr};
}

int lanes_for(double avg)
{
    if (avg <= 3.0)
        return 2;
    if (avg <= 6.0)
        return 4;
    if (avg <= 28.0)
        return 8;
    if (avg <= 64.0)
        return 16;
    return 32;
}

template <int L>
void launch_plain_t(const BlockView& b, int grid, int64_t nrows, const double* x, double* y, bool accum, cudaStream_t stream, const int* abort)
{
    if (accum)
        spmv_plain_kernel<L, true><<<grid, kSpmvBlock, 0, stream>>>(b.rowptr, b.col, b.val, x, y, nrows, abort);
    else
        spmv_plain_kernel<L, false><<<grid, kSpmvBlock, 0, stream>>>(b.rowptr, b.col, b.val, x, y, nrows, abort);
}

template <int THREADS>
void launch_sell_plain_t(const SellBlock& S, int grid, int64_t nrows, const double* x, double* y, bool accum, cudaStream_t stream, const int* abort)
{
    if (accum)
        sell_plain_kernel<THREADS, true><<<grid, THREADS, 0, stream>>>(S.slice_ptr.get(), S.col.get(), S.val.get(), S.perm.get(), x, y, nrows, S.nwin, abort);
    else
        sell_plain_kernel<THREADS, false><<<grid, THREADS, 0, stream>>>(S.slice_ptr.get(), S.col.get(), S.val.get(), S.perm.get(), x, y, nrows, S.nwin, abort);
}

void launch_plain_block(const SpmvPlan& plan, const BlockView& b, int64_t nrows, const double* x, double* y, bool accum, cudaStream_t stream,
                        const int* abort = nullptr)
{
    if (b.sell && plan.sell_threads)
    {
        switch (plan.sell_threads)
        {
            case 256: launch_sell_plain_t<256>(*b.sell, plan.sell_grid_plain, nrows, x, y, accum, stream, abort); break;
            case 1024: launch_sell_plain_t<1024>(*b.sell, plan.sell_grid_plain, nrows, x, y, accum, stream, abort); break;
            default: launch_sell_plain_t<512>(*b.sell, plan.sell_grid_plain, nrows, x, y, accum, stream, abort); break;
        }
        return;
    }
    const int grid = plan.grid;
    switch (plan.lanes)
    {
        case 2: launch_plain_t<2>(b, grid, nrows, x, y, accum, stream, abort); break;
        case 4: launch_plain_t<4>(b, grid, nrows, x, y, accum, stream, abort); break;
        case 8: launch_plain_t<8>(b, grid, nrows, x, y, accum, stream, abort); break;
        case 16: launch_plain_t<16>(b, grid, nrows, x, y, accum, stream, abort); break;
        default: launch_plain_t<32>(b, grid, nrows, x, y, accum, stream, abort); break;
    }
}

template <int THREADS>
void launch_sell_step_t(const SellBlock& S, int grid, int64_t nrows, const double* x_full, const double* f_loc, double* V, int64_t ldv, double* w, FacCtl* ctl,
                        double* H, int m, int i, int restarted, bool symmetric, bool accum, const RedScratch& rs, cudaStream_t stream)
{
#define SB200_SELL_STEP(SYM, ACC)                                                                                                                              \
    sell_step_kernel<THREADS, SYM, ACC><<<grid, THREADS, 0, stream>>>(S.slice_ptr.get(), S.col.get(), S.val.get(), S.perm.get(), x_full, f_loc, V, ldv, w, nrows, \
                                                                      S.nwin, ctl, H, m, i, restarted, rs.partials, rs.ticket)
    if (symmetric && accum)
        SB200_SELL_STEP(true, true);
    else if (symmetric)
        SB200_SELL_STEP(true, false);
    else if (accum)
        SB200_SELL_STEP(false, true);
    else
        SB200_SELL_STEP(false, false);
#undef SB200_SELL_STEP
}

template <int THREADS, int NB, int UN, int RES = 2>
void launch_sell_step_dot_t(const SellBlock& S, int grid, int64_t nrows, const double* x_full, const double* f_loc, double* V, int64_t ldv, double* w, FacCtl* ctl,
                            double* H, int m, int i, int restarted, bool symmetric, bool accum, double* red_out, const RedScratch& rs, cudaStream_t stream)
{
#define SB200_SELL_STEP_DOT(SYM, ACC)                                                                                                                          \
    sell_step_dot_kernel<SYM, ACC, THREADS, NB, UN, RES><<<grid, THREADS, 0, stream>>>(S.slice_ptr.get(), S.col.get(), S.val.get(), S.perm.get(), x_full, f_loc, V, ldv, w, \
                                                                                  nrows, S.nwin, ctl, H, m, i, restarted, red_out, rs.partials, rs.ticket)
    if (symmetric && accum)
        SB200_SELL_STEP_DOT(true, true);
    else if (symmetric)
        SB200_SELL_STEP_DOT(true, false);
    else if (accum)
        SB200_SELL_STEP_DOT(false, true);
    else
        SB200_SELL_STEP_DOT(false, false);
#undef SB200_SELL_STEP_DOT
}

// 0: 512 threads x 4 loads, 1: 512x8 (default), 2: 256x8, 3: 256x16, 4: 256x16 with 8 gathers in flight.  Measured at n = 1e7 on the first 12
// restarts (profiles/r2e_quick_*_n1e7.log): 486 / 505 / 439 / 457 / 453 SpMV-iters/s -- the 256-thread CTAs lose more in the gather phase than
// their deeper V pipeline wins.
int fused_config()
{
    static const int cfg = [] {
        const char* e = std::getenv("SB200_FUSED_CFG");
        if (!e)
            return 1;
        if (std::strcmp(e, "512x4") == 0)
            return 0;
        if (std::strcmp(e, "512x8") == 0)
            return 1;
        if (std::strcmp(e, "256x8") == 0)
            return 2;
        if (std::strcmp(e, "256x16") == 0)
            return 3;
        if (std::strcmp(e, "256x16u8") == 0)
            return 4;
        if (std::strcmp(e, "512x4r3") == 0)
            return 5;  // three resident CTAs per SM (<= 42 registers): finer wave granularity for small row counts
        if (std::strcmp(e, "512x8u8") == 0)
            return 6;  // eight gathers per lane in flight in phase A
        return 0;
    }();
    return cfg;
}

void launch_sell_step_dot(const SellBlock& S, int grid, int64_t nrows, const double* x_full, const double* f_loc, double* V, int64_t ldv, double* w, FacCtl* ctl,
                          double* H, int m, int i, int restarted, bool symmetric, bool accum, double* red_out, const RedScratch& rs, cudaStream_t stream)
{
#define SB200_FUSED_ARGS S, grid, nrows, x_full, f_loc, V, ldv, w, ctl, H, m, i, restarted, symmetric, accum, red_out, rs, stream
    switch (fused_config())
    {
        case 1: launch_sell_step_dot_t<512, 8, 4>(SB200_FUSED_ARGS); break;
        case 2: launch_sell_step_dot_t<256, 8, 4>(SB200_FUSED_ARGS); break;
        case 3: launch_sell_step_dot_t<256, 16, 4>(SB200_FUSED_ARGS); break;
        case 4: launch_sell_step_dot_t<256, 16, 8>(SB200_FUSED_ARGS); break;
        case 5: launch_sell_step_dot_t<512, 4, 4, 3>(SB200_FUSED_ARGS); break;
        case 6: launch_sell_step_dot_t<512, 8, 8>(SB200_FUSED_ARGS); break;
        default: launch_sell_step_dot_t<512, 4, 4>(SB200_FUSED_ARGS); break;
    }
#undef SB200_FUSED_ARGS
}

template <int L>
void launch_step_t(const BlockView& b, int grid, int64_t nrows, const double* x_full, const double* f_loc, double* V, int64_t ldv, double* w, FacCtl* ctl, double* H,
                   int m, int i, int restarted, bool symmetric, bool accum, const RedScratch& rs, cudaStream_t stream)
{
#define SB200_STEP(SYM, ACC)                                                                                                                                     \
    spmv_step_kernel<L, SYM, ACC><<<grid, kSpmvBlock, 0, stream>>>(b.rowptr, b.col, b.val, x_full, f_loc, V, ldv, w, nrows, ctl, H, m, i, restarted, rs.partials, \
                                                                   rs.ticket)
    if (symmetric && accum)
        SB200_STEP(true, true);
    else if (symmetric)
        SB200_STEP(true, false);
    else if (accum)
        SB200_STEP(false, true);
    else
        SB200_STEP(false, false);
#undef SB200_STEP
}

}  // namespace

SpmvPlan make_spmv_plan(const DeviceCsr& A)
{
    SpmvPlan p;
    const int nb = nblocks_of(A);
    const double avg = A.nrows > 0 ? double(A.nnz) / double(A.nrows) / nb : 1.0;
    p.lanes = lanes_for(avg);
    if (const char* e = std::getenv("SB200_SPMV_LANES"))
    {
        const int v = std::atoi(e);
        if (v == 2 || v == 4 || v == 8 || v == 16 || v == 32)
            p.lanes = v;
    }
    const int sms = device_info().sm_count;
    const int64_t rpb = kSpmvBlock;  // 8 warps x 32 rows per CTA iteration
    const int64_t need = (A.nrows + rpb - 1) / rpb;
    // persistent grid: up to 8 resident CTAs of 256 threads per SM
    p.grid = (int) std::max<int64_t>(1, std::min<int64_t>(need, (int64_t) sms * 8));
    // A/B knob for the next round (profiles/README.md, "grid quantisation"): SB200_SPMV_GRID=blocks launches one CTA per 256-row block
    // when that fits the reduction scratch, so that the hardware balances a small number of rounds instead of leaving part of the last
    // persistent round idle.  Default: unchanged (the GPU-verified persistent grid).
    if (const char* e = std::getenv("SB200_SPMV_GRID"))
        if (std::strcmp(e, "blocks") == 0 && need <= (int64_t) reduction_max_grid(sms))
            p.grid = (int) std::max<int64_t>(1, need);
    // sliced layout present (build_sell_layout): lane-per-row kernels, persistent grid of one window per CTA iteration
    const SellBlock& S0 = A.blocks.empty() ? A.sell : A.blocks[0].sell;
    if (S0.built())
    {
        p.sell_threads = 512;
        if (const char* e = std::getenv("SB200_SELL_THREADS"))
        {
            const int v = std::atoi(e);
            if (v == 256 || v == 512 || v == 1024)
                p.sell_threads = v;
        }
        // One CTA per window while the grid fits the reduction scratch of the fused step kernel (reduction_max_grid, FacBase::alloc_common):
        // the hardware then balances the windows over the SMs as CTAs retire.  Larger operands use a persistent grid of resident
        // CTAs that stride over the windows (>= 20 windows per CTA at that size, so the last round costs little).
        const int64_t resident = (int64_t) sms * sell_min_blocks(p.sell_threads);
        const char* force = std::getenv("SB200_SELL_PERSISTENT");  // test / A-B knob: always use the persistent strided grid
        const bool persistent = (force && force[0] == '1') || S0.nwin > (int64_t) reduction_max_grid(sms);
        p.sell_grid = (int) std::max<int64_t>(1, persistent ? std::min<int64_t>(S0.nwin, resident) : S0.nwin);
        p.sell_grid_plain = persistent && force ? p.sell_grid : (int) std::max<int64_t>(1, std::min<int64_t>(S0.nwin, 1 << 30));  // no reduction: one CTA per window
        // fused step head + panel pass (sell_step_dot_kernel): persistent, two resident CTAs of 512 threads per SM (per-lane accumulators
        // live across the windows of a CTA).  SB200_FUSE_DOT=0 keeps the separate panel pass (A/B knob).
        const char* fd = std::getenv("SB200_FUSE_DOT");
        const char* fc = std::getenv("SB200_FUSED_CFG");
        const int fused_res = (fc && std::strcmp(fc, "512x4r3") == 0) ? 3 : 2;
        p.sell_grid_fused = (fd && fd[0] == '0') ? 0 : (int) std::max<int64_t>(1, std::min<int64_t>(S0.nwin, (int64_t) sms * fused_res));
    }
    return p;
}

// Number of column blocks for an operand of order n: slices of at most SB200_XSLICE_MB (default 40 MB, measured best at n = 1e7) of x.
int choose_col_blocks(int64_t n)
{
    double slice_mb = 40.0;
    if (const char* e = std::getenv("SB200_XSLICE_MB"))
        slice_mb = std::max(1e-3, std::atof(e));
    const double x_mb = 8.0 * double(n) / (1024.0 * 1024.0);
    const int nb = (int) std::ceil(x_mb / slice_mb);
    return std::max(1, std::min(nb, kMaxColBlocks));
}

void launch_spmv(const DeviceCsr& A, const SpmvPlan& plan, const double* x, double* y, cudaStream_t stream)
{
    if (A.nrows == 0)
        return;
    const int nb = nblocks_of(A);
    for (int c = 0; c < nb; c++)
        launch_plain_block(plan, view_of(A, c), A.nrows, A.x_of_block(x, c), y, c > 0, stream);
    SB200_CUDA_CHECK(cudaGetLastError());
}

int spmv_num_blocks(const DeviceCsr& A) { return nblocks_of(A); }

void launch_spmv_z(const DeviceCsrZ& A, const double* x_ri, double* y_ri, cudaStream_t stream)
{
    if (A.n == 0)
        return;
    const int lanes = lanes_for(double(A.nnz) / double(A.n));
    const int sms = device_info().sm_count;
    const int64_t need = (A.n * lanes + kSpmvBlock - 1) / kSpmvBlock;
    const int grid = (int) std::max<int64_t>(1, std::min<int64_t>(need, (int64_t) sms * 4));
    switch (lanes)
    {
        case 2: spmv_z_kernel<2><<<grid, kSpmvBlock, 0, stream>>>(A.rowptr.get(), A.col.get(), A.val.get(), x_ri, y_ri, A.n); break;
        case 4: spmv_z_kernel<4><<<grid, kSpmvBlock, 0, stream>>>(A.rowptr.get(), A.col.get(), A.val.get(), x_ri, y_ri, A.n); break;
        case 8: spmv_z_kernel<8><<<grid, kSpmvBlock, 0, stream>>>(A.rowptr.get(), A.col.get(), A.val.get(), x_ri, y_ri, A.n); break;
        case 16: spmv_z_kernel<16><<<grid, kSpmvBlock, 0, stream>>>(A.rowptr.get(), A.col.get(), A.val.get(), x_ri, y_ri, A.n); break;
        default: spmv_z_kernel<32><<<grid, kSpmvBlock, 0, stream>>>(A.rowptr.get(), A.col.get(), A.val.get(), x_ri, y_ri, A.n); break;
    }
    SB200_CUDA_CHECK(cudaGetLastError());
}

bool launch_spmv_step_block(const DeviceCsr& A, const SpmvPlan& plan, int c, const double* x_block, const double* f_loc, double* V, int64_t ldv, double* w,
                            FacCtl* ctl, double* H, int m, int i, int restarted, bool symmetric, const RedScratch& rs, cudaStream_t stream, double* dot_out)
{
    SB200_REQUIRE(plan.grid <= rs.max_grid, SB200_LOGIC, "spmv: reduction scratch too small");
    const int nb = nblocks_of(A);
    if (c + 1 < nb)
    {
        // all but the last column block accumulate the raw product into w
        launch_plain_block(plan, view_of(A, c), A.nrows, x_block, w, c > 0, stream, &ctl->abort);
        SB200_CUDA_CHECK(cudaGetLastError());
        return false;
    }
    const BlockView b = view_of(A, nb - 1);
    const bool accum = nb > 1;
    if (b.sell && plan.sell_threads && dot_out && plan.sell_grid_fused > 0 && i >= 1 && i < kPanelMaxCols)
    {
        SB200_REQUIRE(plan.sell_grid_fused <= rs.max_grid, SB200_LOGIC, "spmv: reduction scratch too small");
        launch_sell_step_dot(*b.sell, plan.sell_grid_fused, A.nrows, x_block, f_loc, V, ldv, w, ctl, H, m, i, restarted, symmetric, accum, dot_out, rs, stream);
        SB200_CUDA_CHECK(cudaGetLastError());
        return true;
    }
    if (b.sell && plan.sell_threads)
    {
        SB200_REQUIRE(plan.sell_grid <= rs.max_grid, SB200_LOGIC, "spmv: reduction scratch too small");
        switch (plan.sell_threads)
        {
            case 256: launch_sell_step_t<256>(*b.sell, plan.sell_grid, A.nrows, x_block, f_loc, V, ldv, w, ctl, H, m, i, restarted, symmetric, accum, rs, stream); break;
            case 1024: launch_sell_step_t<1024>(*b.sell, plan.sell_grid, A.nrows, x_block, f_loc, V, ldv, w, ctl, H, m, i, restarted, symmetric, accum, rs, stream); break;
            default: launch_sell_step_t<512>(*b.sell, plan.sell_grid, A.nrows, x_block, f_loc, V, ldv, w, ctl, H, m, i, restarted, symmetric, accum, rs, stream); break;
        }
        SB200_CUDA_CHECK(cudaGetLastError());
        return false;
    }
    switch (plan.lanes)
    {
        case 2: launch_step_t<2>(b, plan.grid, A.nrows, x_block, f_loc, V, ldv, w, ctl, H, m, i, restarted, symmetric, accum, rs, stream); break;
        case 4: launch_step_t<4>(b, plan.grid, A.nrows, x_block, f_loc, V, ldv, w, ctl, H, m, i, restarted, symmetric, accum, rs, stream); break;
        case 8: launch_step_t<8>(b, plan.grid, A.nrows, x_block, f_loc, V, ldv, w, ctl, H, m, i, restarted, symmetric, accum, rs, stream); break;
        case 16: launch_step_t<16>(b, plan.grid, A.nrows, x_block, f_loc, V, ldv, w, ctl, H, m, i, restarted, symmetric, accum, rs, stream); break;
        default: launch_step_t<32>(b, plan.grid, A.nrows, x_block, f_loc, V, ldv, w, ctl, H, m, i, restarted, symmetric, accum, rs, stream); break;
    }
    SB200_CUDA_CHECK(cudaGetLastError());
    return false;
}

bool launch_spmv_step(const DeviceCsr& A, const SpmvPlan& plan, const double* x_full, const double* f_loc, double* V, int64_t ldv, double* w, FacCtl* ctl,
                      double* H, int m, int i, int restarted, bool symmetric, const RedScratch& rs, cudaStream_t stream, double* dot_out)
{
    const int nb = nblocks_of(A);
    bool fused = false;
    for (int c = 0; c < nb; c++)
        fused = launch_spmv_step_block(A, plan, c, A.x_of_block(x_full, c), f_loc, V, ldv, w, ctl, H, m, i, restarted, symmetric, rs, stream, dot_out);
    return fused;
}

}  // namespace sb200

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
// the read-only path and combines with log2(L) shuffles.  CSR streams are read with
// L1::no_allocate + L2 evict_first so that the gathered vector x (8n bytes) stays L2-resident.
// Algorithmic bytes per row at d nnz/row: 12 d + 4 + 8 (x once) + 8 (y)  (SURVEY §8d).
#include "kernels.h"

namespace sb200 {

namespace {

constexpr int kSpmvBlock = 256;

template <int L>
// All 32 lanes of a warp must call this together (sub-warp shuffle with a full mask); a sub-warp whose
// row is out of range passes start == end.
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

template <int L>
__global__ void __launch_bounds__(kSpmvBlock) spmv_plain_kernel(const int* __restrict__ rowptr, const int* __restrict__ col, const double* __restrict__ val,
                                                                const double* __restrict__ x, double* __restrict__ y, int64_t nrows)
{
    const uint64_t pol_stream = l2_policy_evict_first();
    const uint64_t pol_keep = l2_policy_evict_last();
    constexpr int RPB = kSpmvBlock / L;  // rows per CTA per iteration
    const int lane = threadIdx.x % L;
    const int sub = threadIdx.x / L;
    // the loop bound is CTA-uniform so that every warp executes the shuffles with all 32 lanes
    for (int64_t rbase = (int64_t) blockIdx.x * RPB; rbase < nrows; rbase += (int64_t) gridDim.x * RPB)
    {
        const int64_t row = rbase + sub;
        const bool active = row < nrows;
        const int start = active ? __ldg(rowptr + row) : 0;
        const int end = active ? __ldg(rowptr + row + 1) : 0;
        const double s = row_dot<L>(start, end, col, val, x, lane, pol_stream, pol_keep);
        if (active && lane == 0)
            y[row] = s;
    }
}

// Fused step head.  x_full: un-normalised residual (all n entries), f_loc: this rank's rows of it.
template <int L, bool SYM>
__global__ void __launch_bounds__(kSpmvBlock)
    spmv_step_kernel(const int* __restrict__ rowptr, const int* __restrict__ col, const double* __restrict__ val, const double* __restrict__ x_full,
                     const double* __restrict__ f_loc, double* __restrict__ V, int64_t ldv, double* __restrict__ w, int64_t nrows, FacCtl* ctl, double* H, int m,
                     int i, int restarted, double* partials, unsigned int* ticket)
{
    const uint64_t pol_stream = l2_policy_evict_first();
    const uint64_t pol_keep = l2_policy_evict_last();
    constexpr int RPB = kSpmvBlock / L;
    const int lane = threadIdx.x % L;
    const int sub = threadIdx.x / L;
    const double beta = ctl->beta;
    const double hsub = restarted ? 0.0 : beta;
    double* __restrict__ vi = V + (int64_t) i * ldv;
    const double* __restrict__ vp = V + (int64_t) (i - 1) * ldv;

    double part = 0.0;
    for (int64_t rbase = (int64_t) blockIdx.x * RPB; rbase < nrows; rbase += (int64_t) gridDim.x * RPB)
    {
        const int64_t row = rbase + sub;
        const bool active = row < nrows;
        const int start = active ? __ldg(rowptr + row) : 0;
        const int end = active ? __ldg(rowptr + row + 1) : 0;
        const double s = row_dot<L>(start, end, col, val, x_full, lane, pol_stream, pol_keep);
        if (active && lane == 0)
        {
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

template <int L>
void launch_plain_t(const DeviceCsr& A, const SpmvPlan& plan, const double* x, double* y, cudaStream_t stream)
{
    spmv_plain_kernel<L><<<plan.grid, kSpmvBlock, 0, stream>>>(A.rowptr.get(), A.col.get(), A.val.get(), x, y, A.nrows);
}

template <int L>
void launch_step_t(const DeviceCsr& A, const SpmvPlan& plan, const double* x_full, const double* f_loc, double* V, int64_t ldv, double* w, FacCtl* ctl,
                   double* H, int m, int i, int restarted, bool symmetric, const RedScratch& rs, cudaStream_t stream)
{
    if (symmetric)
        spmv_step_kernel<L, true><<<plan.grid, kSpmvBlock, 0, stream>>>(A.rowptr.get(), A.col.get(), A.val.get(), x_full, f_loc, V, ldv, w, A.nrows, ctl, H, m, i,
                                                                        restarted, rs.partials, rs.ticket);
    else
        spmv_step_kernel<L, false><<<plan.grid, kSpmvBlock, 0, stream>>>(A.rowptr.get(), A.col.get(), A.val.get(), x_full, f_loc, V, ldv, w, A.nrows, ctl, H, m,
                                                                         i, restarted, rs.partials, rs.ticket);
}

}  // namespace

SpmvPlan make_spmv_plan(const DeviceCsr& A)
{
    SpmvPlan p;
    const double avg = A.nrows > 0 ? double(A.nnz) / double(A.nrows) : 1.0;
    if (avg <= 3.0)
        p.lanes = 2;
    else if (avg <= 6.0)
        p.lanes = 4;
    else if (avg <= 28.0)
        p.lanes = 8;
    else if (avg <= 64.0)
        p.lanes = 16;
    else
        p.lanes = 32;
    const int sms = device_info().sm_count;
    const int64_t rpb = kSpmvBlock / p.lanes;
    const int64_t need = (A.nrows + rpb - 1) / rpb;
    // persistent grid: 8 resident CTAs of 256 threads per SM (2048 threads/SM)
    p.grid = (int) std::max<int64_t>(1, std::min<int64_t>(need, (int64_t) sms * 8));
    return p;
}

void launch_spmv(const DeviceCsr& A, const SpmvPlan& plan, const double* x, double* y, cudaStream_t stream)
{
    if (A.nrows == 0)
        return;
    switch (plan.lanes)
    {
        case 2: launch_plain_t<2>(A, plan, x, y, stream); break;
        case 4: launch_plain_t<4>(A, plan, x, y, stream); break;
        case 8: launch_plain_t<8>(A, plan, x, y, stream); break;
        case 16: launch_plain_t<16>(A, plan, x, y, stream); break;
        default: launch_plain_t<32>(A, plan, x, y, stream); break;
    }
    SB200_CUDA_CHECK(cudaGetLastError());
}

void launch_spmv_step(const DeviceCsr& A, const SpmvPlan& plan, const double* x_full, const double* f_loc, double* V, int64_t ldv, double* w, FacCtl* ctl,
                      double* H, int m, int i, int restarted, bool symmetric, const RedScratch& rs, cudaStream_t stream)
{
    SB200_REQUIRE(plan.grid <= rs.max_grid, SB200_LOGIC, "spmv: reduction scratch too small");
    switch (plan.lanes)
    {
        case 2: launch_step_t<2>(A, plan, x_full, f_loc, V, ldv, w, ctl, H, m, i, restarted, symmetric, rs, stream); break;
        case 4: launch_step_t<4>(A, plan, x_full, f_loc, V, ldv, w, ctl, H, m, i, restarted, symmetric, rs, stream); break;
        case 8: launch_step_t<8>(A, plan, x_full, f_loc, V, ldv, w, ctl, H, m, i, restarted, symmetric, rs, stream); break;
        case 16: launch_step_t<16>(A, plan, x_full, f_loc, V, ldv, w, ctl, H, m, i, restarted, symmetric, rs, stream); break;
        default: launch_step_t<32>(A, plan, x_full, f_loc, V, ldv, w, ctl, H, m, i, restarted, symmetric, rs, stream); break;
    }
    SB200_CUDA_CHECK(cudaGetLastError());
}

}  // namespace sb200

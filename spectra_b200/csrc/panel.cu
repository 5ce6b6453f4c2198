// Fused Krylov-panel kernels for sm_100a (K3/K4/K5/K6 of SURVEY.md §2.1).
//
// One pass over the tall-skinny basis V (nrows x j, column-major, ld = ldv) does what the
// reference does in two or three BLAS-2 calls:
//   PANEL_FORM : f = w - alpha v_i; c = V^T f; ||f||^2          Lanczos.h:145-152
//   PANEL_CORR : f = x - V c;       c' = V^T f; ||f||^2         Lanczos.h:171-179, Arnoldi.h:254-262,281-287
//   PANEL_DOT  : c = V^T x; ||x||^2                             ArnoldiOp.h:144-148 (adjoint_product)
// f_new[r] only needs row r of V, so the pass is row-local and V is read exactly once:
// algorithmic bytes = 8*nrows*j (V) + 8*nrows (x) + 8*nrows (f)   (SURVEY §8d "fused K4+K3").
//
// Mapping: a CTA of 8 warps is split into NG column groups (16 columns each) x RS = 8/NG row
// slices.  A lane owns 2 consecutive rows (one 128-bit load per column), keeps its 16x2 tile of V
// in registers, and (CORR) exchanges the 16-column partial dot products through shared memory so
// that every group sees the complete f_new of its rows.  Per-thread accumulators are combined by a
// shuffle tree, then across row slices and CTAs in a fixed order (bit-reproducible run to run).
// V is streamed with L1::no_allocate / L2 evict_first; x, f stay L2-resident between kernels.
//
// The restart GEMM  V[:, :kk] <- V Q  (Arnoldi.h:320-340) and  X = V S  (HermEigsBase.h:467) use
// one thread per row with the whole row of V in registers and Q staged in shared memory.
#include <cstdlib>

#include "kernels.h"
#include "peer_dev.cuh"

namespace sb200 {

namespace {

constexpr int kPanelBlock = 256;
constexpr int kColsPerGroup = 16;

// CPLX (Hermitian path): V, x, f hold interleaved complex numbers and are addressed as real arrays of twice the
// length (nrows, ldv in doubles), so the double2 a lane loads per column IS one complex entry; the products become
// c = V^H f (conjugated, ArnoldiOp.h:144-148 with Scalar = complex) and f -= V c with complex c.  A column group then holds 8
// columns (16 accumulators: Re and Im).  Reduction layout: red[k] = Re c_k, red[kRedNrm] = ||f||^2, red[kRedNrm + 1 + k] = Im c_k
// (the same layout is read from `coef` by the CORR mode), which limits the complex panel to 63 columns.
template <int NG, int MODE, bool CPLX>
__global__ void __launch_bounds__(kPanelBlock, 2)
    panel_kernel(const double* __restrict__ V, int64_t ldv, int64_t nrows, int j, const double* x, double* f_out, const double* __restrict__ coef,
                 double* red_out, double* partials, unsigned int* ticket, const int* pred, const PeerX push, const int* abort, int64_t row_limit)
{
    // speculatively enqueued pass: skip when the device-side flag says no correction is needed; sweep mode: skip after an abort
    if (pred != nullptr && *pred == 0)
        return;
    if (abort != nullptr && *abort != 0)
        return;
    constexpr int CPG = CPLX ? 8 : kColsPerGroup;  // columns per group
    constexpr int NCOEF = CPLX ? 2 * kPanelMaxCols : kPanelMaxCols;
    constexpr int RS = 8 / NG;        // row slices
    constexpr int RPI = RS * 64;      // rows per CTA iteration
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = warp % NG, rs = warp / NG;
    const uint64_t pol = l2_policy_evict_first();

    __shared__ double s_c[NCOEF];               // CPLX: Re c_k at [k], Im c_k at [kPanelMaxCols + k]
    __shared__ double2 s_p[2][NG][RS * 32];     // CORR: per-group partial  sum_k c_k V[r,k]
    __shared__ double s_part[RS][NCOEF + 1];    // CPLX: Re at [k], Im at [kPanelMaxCols + k], ||f||^2 at [NCOEF]
    __shared__ double s_out[kRedStride];

    if (MODE == PANEL_CORR)
    {
        if (threadIdx.x < kPanelMaxCols)
        {
            s_c[threadIdx.x] = (threadIdx.x < j) ? coef[threadIdx.x] : 0.0;
            if (CPLX)
                s_c[kPanelMaxCols + threadIdx.x] = (threadIdx.x < j) ? coef[kRedNrm + 1 + threadIdx.x] : 0.0;
        }
        __syncthreads();
    }
    const double alpha = (MODE == PANEL_FORM) ? coef[0] : 0.0;

    double acc[CPG];
    double acci[CPLX ? CPG : 1];
#pragma unroll
    for (int kk = 0; kk < CPG; kk++)
        acc[kk] = 0.0;
#pragma unroll
    for (int kk = 0; kk < (CPLX ? CPG : 1); kk++)
        acci[kk] = 0.0;
    double nrm = 0.0;
    const double* cg = s_c + g * CPG;  // CORR coefficients of this group (shared-memory broadcast)

    const double* __restrict__ vi = V + (int64_t) (j - 1) * ldv;  // FORM: v_i is the last panel column
    int buf = 0;
    for (int64_t base = (int64_t) blockIdx.x * RPI; base < nrows; base += (int64_t) gridDim.x * RPI)
    {
        const int64_t r0 = base + rs * 64 + lane * 2;
        const bool valid = r0 < row_limit;  // full pass: ldv (padding rows [nrows, ldv) hold zeros); row-range pass: the range length
        double2 v[CPG];
#pragma unroll
        for (int kk = 0; kk < CPG; kk++)
        {
            const int k = g * CPG + kk;
            v[kk] = (valid && k < j) ? ld_stream_f64x2(V + r0 + (int64_t) k * ldv, pol) : make_double2(0.0, 0.0);
        }
        double2 xv = valid ? *reinterpret_cast<const double2*>(x + r0) : make_double2(0.0, 0.0);
        double2 fn;
        if (MODE == PANEL_FORM)
        {
            const double2 vv = valid ? *reinterpret_cast<const double2*>(vi + r0) : make_double2(0.0, 0.0);
            fn.x = xv.x - alpha * vv.x;  // f = w - H(i,i) v_i   (Lanczos.h:145)
            fn.y = xv.y - alpha * vv.y;
        }
        else if (MODE == PANEL_CORR)
        {
            double2 p = make_double2(0.0, 0.0);
#pragma unroll
            for (int kk = 0; kk < CPG; kk++)
            {
                p.x = fma(cg[kk], v[kk].x, p.x);
                p.y = fma(cg[kk], v[kk].y, p.y);
                if (CPLX)
                {
                    // (cr + i ci) (vx + i vy)
                    p.x = fma(-cg[kPanelMaxCols + kk], v[kk].y, p.x);
                    p.y = fma(cg[kPanelMaxCols + kk], v[kk].x, p.y);
                }
            }
            if (NG > 1)
            {
                s_p[buf][g][rs * 32 + lane] = p;
                __syncthreads();
                p = s_p[buf][0][rs * 32 + lane];
#pragma unroll
                for (int q = 1; q < NG; q++)
                {
                    const double2 t = s_p[buf][q][rs * 32 + lane];
                    p.x += t.x;
                    p.y += t.y;
                }
                buf ^= 1;
            }
            fn.x = xv.x - p.x;  // f -= V * Vf   (Lanczos.h:171)
            fn.y = xv.y - p.y;
        }
        else
        {
            fn = xv;
        }
        if (MODE != PANEL_DOT && g == 0 && valid)
            *reinterpret_cast<double2*>(f_out + r0) = fn;
        if (MODE == PANEL_CORR && !CPLX && push.np > 0 && valid && r0 + push.row0 < push.rows)
        {
            // row-sharded runs: the new residual is the next SpMV operand of EVERY rank -- store this rank's rows straight into all
            // operand buffers (peer memory over NVLink, chunk-major layout) instead of all-gathering them afterwards.  Every column group
            // holds the complete f_new of its rows, so the NG groups share the destinations (group g serves ranks g, g + NG, ...) and the
            // remote stores are spread over all warps of the CTA.
            const int64_t rl = r0 + push.row0;  // local row
            const int64_t c = rl / push.len;
            const int64_t dst = c * push.stride + (int64_t) push.rank * push.len + (rl - c * push.len);
            for (int p = g; p < push.np; p += NG)
                st_peer_f64x2(push.dst[p] + dst, fn);
        }
#pragma unroll
        for (int kk = 0; kk < CPG; kk++)
        {
            acc[kk] = fma(v[kk].x, fn.x, acc[kk]);
            acc[kk] = fma(v[kk].y, fn.y, acc[kk]);
            if (CPLX)
            {
                // conj(v) f = (vx fx + vy fy) + i (vx fy - vy fx)
                acci[kk] = fma(v[kk].x, fn.y, acci[kk]);
                acci[kk] = fma(-v[kk].y, fn.x, acci[kk]);
            }
        }
        if (g == 0)
        {
            nrm = fma(fn.x, fn.x, nrm);
            nrm = fma(fn.y, fn.y, nrm);
        }
    }

    // ---- CTA-level combine (fixed order) ----
#pragma unroll
    for (int kk = 0; kk < CPG; kk++)
    {
        const double s = warp_sum(acc[kk]);
        if (lane == 0)
            s_part[rs][g * CPG + kk] = s;
        if (CPLX)
        {
            const double si = warp_sum(acci[kk]);
            if (lane == 0)
                s_part[rs][kPanelMaxCols + g * CPG + kk] = si;
        }
    }
    if (g == 0)
    {
        const double s = warp_sum(nrm);
        if (lane == 0)
            s_part[rs][NCOEF] = s;
    }
    if (NG * CPG < kPanelMaxCols)
    {
        // columns of the unused groups
        for (int t = threadIdx.x; t < RS * NCOEF; t += kPanelBlock)
        {
            const int c = t % NCOEF, r = t / NCOEF;
            if ((c % kPanelMaxCols) >= NG * CPG)
                s_part[r][c] = 0.0;
        }
    }
    __syncthreads();
    double cta = 0.0;
    const int t = threadIdx.x;
    const int K = CPLX ? 2 * j + 1 : j + 1;  // reduced values: [0, j) Re c (or c), slot j = ||f||^2, (j, 2j] Im c
    if (t < K)
    {
        const int src = CPLX ? ((t < j) ? t : ((t == j) ? NCOEF : kPanelMaxCols + (t - j - 1))) : ((t == j) ? kPanelMaxCols : t);  // slot j carries ||f||^2
#pragma unroll
        for (int q = 0; q < RS; q++)
            cta += s_part[q][src];
    }
    if (grid_reduce_fixed_order<kPanelBlock>(cta, K, partials, ticket, s_out))
    {
        if (t < j)
            red_out[t] = s_out[t];
        if (t == 0)
            red_out[kRedNrm] = s_out[j];
        if (CPLX && t > j && t < K)
            red_out[kRedNrm + 1 + (t - j - 1)] = s_out[t];
    }
}

template <int MODE>
void launch_panel_mode(const double* V, int64_t ldv, int64_t nrows, int j, const double* x, double* f_out, const double* coef, double* red_out, int grid,
                       const RedScratch& rs, const int* pred, cudaStream_t stream, const PeerX& push, const int* abort, int64_t row_limit)
{
    if (j <= 16)
        panel_kernel<1, MODE, false><<<grid, kPanelBlock, 0, stream>>>(V, ldv, nrows, j, x, f_out, coef, red_out, rs.partials, rs.ticket, pred, push, abort, row_limit);
    else if (j <= 32)
        panel_kernel<2, MODE, false><<<grid, kPanelBlock, 0, stream>>>(V, ldv, nrows, j, x, f_out, coef, red_out, rs.partials, rs.ticket, pred, push, abort, row_limit);
    else
        panel_kernel<4, MODE, false><<<grid, kPanelBlock, 0, stream>>>(V, ldv, nrows, j, x, f_out, coef, red_out, rs.partials, rs.ticket, pred, push, abort, row_limit);
}

// complex panel: 8 columns per group
template <int MODE>
void launch_panel_mode_z(const double* V, int64_t ldv, int64_t nrows, int j, const double* x, double* f_out, const double* coef, double* red_out, int grid,
                         const RedScratch& rs, const int* pred, cudaStream_t stream)
{
    const PeerX none{};
    if (j <= 8)
        panel_kernel<1, MODE, true><<<grid, kPanelBlock, 0, stream>>>(V, ldv, nrows, j, x, f_out, coef, red_out, rs.partials, rs.ticket, pred, none, nullptr, ldv);
    else if (j <= 16)
        panel_kernel<2, MODE, true><<<grid, kPanelBlock, 0, stream>>>(V, ldv, nrows, j, x, f_out, coef, red_out, rs.partials, rs.ticket, pred, none, nullptr, ldv);
    else if (j <= 32)
        panel_kernel<4, MODE, true><<<grid, kPanelBlock, 0, stream>>>(V, ldv, nrows, j, x, f_out, coef, red_out, rs.partials, rs.ticket, pred, none, nullptr, ldv);
    else
        panel_kernel<8, MODE, true><<<grid, kPanelBlock, 0, stream>>>(V, ldv, nrows, j, x, f_out, coef, red_out, rs.partials, rs.ticket, pred, none, nullptr, ldv);
}

// ---------------------------------------------------------------------------------------------
// decide kernels (single warp)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_max(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__device__ __forceinline__ double panel_max_abs(const double* red, int j, int lane)
{
    double mx = 0.0;
    for (int k = lane; k < j; k += 32)
        mx = fmax(mx, fabs(red[k]));
    return warp_max(mx);
}

// Lanczos step, decisions after a panel pass.  The device schedule is classical Gram-Schmidt with re-orthogonalisation checks:
//   stage 0, after  c = V[:, :j]^T w  (fused into the operator kernel or PANEL_DOT):
//            H(i,i) = c_i (= <v_i, w>, Lanczos.h:142), H(i-1,i) = H(i,i-1) = hsub + c_{i-1}; c becomes the coefficient vector of the first
//            pass  f = w - V c.  In exact arithmetic this is the reference's  f = w - H(i,i) v_i  followed by its first correction
//            f -= V (V^T f), h += Vf  (Lanczos.h:145-175: Vf = c - H(i,i) e_i up to |H(i,i)| O(eps) terms) -- the correction the
//            reference applies on all but a fraction of a percent of the steps -- done in one pass over V instead of two.
//   stage 1, after that pass (red = V^T f, ||f||^2): count = 1, beta = ||f||, the test of Lanczos.h:156 on the new f.
//   stage 2, after a further correction pass  f -= V c  (Lanczos.h:171-179): h += c with the coefficients just applied, count += 1, test.
__device__ __forceinline__ void lanczos_decide_body(FacCtl* ctl, double* H, int m, double beta_thresh, int stage, int sweep, int two_part, int lane)
{
    const int i = ctl->i, j = i + 1;
    if (stage == 0)
    {
        for (int k = lane; k < j; k += 32)
            ctl->c[k] = ctl->red[k];
        if (lane == 0)
        {
            H[i + (int64_t) i * m] = ctl->red[i];
            const double hu = ctl->hsub + ctl->red[i - 1];
            H[(i - 1) + (int64_t) i * m] = hu;
            H[i + (int64_t) (i - 1) * m] = hu;
            ctl->count = 0;
            ctl->need_corr = 1;
        }
        return;
    }
    if (two_part)
    {
        // the pass ran in two row ranges (first range -> red2, second -> red): their sum, in that order, is the result of the pass
        for (int k = lane; k < j; k += 32)
            ctl->red[k] = ctl->red2[k] + ctl->red[k];
        if (lane == 0)
            ctl->red[kRedNrm] = ctl->red2[kRedNrm] + ctl->red[kRedNrm];
        __syncwarp();
    }
    int count = ctl->count;
    if (lane == 0 && stage == 2)
    {
        // h <- h + Vf   (Lanczos.h:172-175) with the coefficients the pass just applied
        const double hu = H[(i - 1) + (int64_t) i * m] + ctl->c[i - 1];
        H[(i - 1) + (int64_t) i * m] = hu;
        H[i + (int64_t) (i - 1) * m] = hu;
        H[i + (int64_t) i * m] += ctl->c[i];
    }
    count += 1;
    const double ortho_err = panel_max_abs(ctl->red, j, lane);
    double beta = sqrt(ctl->red[kRedNrm]);  // ||f||   (Lanczos.h:146,177)
    __syncwarp();
    for (int k = lane; k < j; k += 32)
        ctl->c[k] = ctl->red[k];
    int need = (count < 5) && (ortho_err > kEps * beta);  // Lanczos.h:156
    int zeroed = 0;
    if (need && beta < beta_thresh)  // Lanczos.h:163-168
    {
        zeroed = 1;
        beta = 0.0;
        need = 0;
    }
    if (lane == 0)
    {
        ctl->beta = beta;
        ctl->ortho_err = ortho_err;
        ctl->count = count;
        ctl->need_corr = need;
        ctl->f_zeroed = zeroed;
        // sweep mode: the next step may only follow on the device when this one is complete and the host-side tests at the head of the
        // next step (beta < near_0, beta < sqrt(eps): Lanczos.h:99-113) cannot fire
        if (sweep && (need || zeroed || !(beta >= 1.4901161193847656e-08)))
            ctl->abort = 1;
    }
}

__global__ void lanczos_decide_kernel(FacCtl* ctl, double* H, int m, double beta_thresh, int stage, int predicated, int sweep, int two_part)
{
    if (predicated && ctl->need_corr == 0)
        return;  // the speculative correction pass was skipped
    if (ctl->abort != 0)
        return;  // sweep mode: an earlier step handed control back to the host
    lanczos_decide_body(ctl, H, m, beta_thresh, stage, sweep, two_part, threadIdx.x);
}

// Peer-memory runs: the all-reduce of the panel result (count values of ctl->red, one-shot mailbox protocol of peer.cu) and the decisions in
// ONE launch of 128 threads -- a sharded Lanczos step then is K-A, K-A+B, this kernel, K-C, this kernel.
__global__ void __launch_bounds__(128) lanczos_decide_peer_kernel(PeerCtl pc, int count, FacCtl* ctl, double* H, int m, double beta_thresh, int stage, int sweep,
                                                                  int two_part)
{
    if (ctl->abort != 0)
        return;  // identical on every rank: all ranks skip the same rounds
    peer_allreduce_body(pc, ctl->red, count, 0);
    if (threadIdx.x < 32)
        lanczos_decide_body(ctl, H, m, beta_thresh, stage, sweep, two_part, threadIdx.x);
}

// Complex (Hermitian) flavour of lanczos_decide_kernel: ctl->red / ctl->c carry Re at [k] and Im at [kRedNrm + 1 + k]; the
// orthogonality error is the largest complex modulus (Vf.cwiseAbs().maxCoeff(), Lanczos.h:153); H is kept real -- the restart
// reads m_fac_H.real() (HermEigsBase.h:131, 207) -- so only the real parts of the corrections enter it (Lanczos.h:172-175).
__global__ void lanczos_decide_z_kernel(FacCtl* ctl, double* H, int m, double beta_thresh, int stage, int predicated)
{
    if (predicated && ctl->need_corr == 0)
        return;
    const int lane = threadIdx.x;
    const int i = ctl->i, j = i + 1;
    if (stage == 0)
    {
        for (int k = lane; k < j; k += 32)
        {
            ctl->c[k] = ctl->red[k];
            ctl->c[kRedNrm + 1 + k] = ctl->red[kRedNrm + 1 + k];
        }
        if (lane == 0)
        {
            H[i + (int64_t) i * m] = ctl->red[i];  // H(i,i) = Re <v, w>   (Lanczos.h:142)
            const double hu = ctl->hsub + ctl->red[i - 1];
            H[(i - 1) + (int64_t) i * m] = hu;
            H[i + (int64_t) (i - 1) * m] = hu;
            ctl->count = 0;
            ctl->need_corr = 1;
        }
        return;
    }
    int count = ctl->count;
    if (lane == 0 && stage == 2)
    {
        const double hu = H[(i - 1) + (int64_t) i * m] + ctl->c[i - 1];
        H[(i - 1) + (int64_t) i * m] = hu;
        H[i + (int64_t) (i - 1) * m] = hu;
        H[i + (int64_t) i * m] += ctl->c[i];
    }
    count += 1;
    double mx = 0.0;
    for (int k = lane; k < j; k += 32)
        mx = fmax(mx, hypot(ctl->red[k], ctl->red[kRedNrm + 1 + k]));
    const double ortho_err = warp_max(mx);
    double beta = sqrt(ctl->red[kRedNrm]);
    __syncwarp();
    for (int k = lane; k < j; k += 32)
    {
        ctl->c[k] = ctl->red[k];
        ctl->c[kRedNrm + 1 + k] = ctl->red[kRedNrm + 1 + k];
    }
    int need = (count < 5) && (ortho_err > kEps * beta);  // Lanczos.h:156
    int zeroed = 0;
    if (need && beta < beta_thresh)  // Lanczos.h:163-168
    {
        zeroed = 1;
        beta = 0.0;
        need = 0;
    }
    if (lane == 0)
    {
        ctl->beta = beta;
        ctl->ortho_err = ortho_err;
        ctl->count = count;
        ctl->need_corr = need;
        ctl->f_zeroed = zeroed;
    }
}

__global__ void arnoldi_decide_kernel(FacCtl* ctl, double* H, int m, double beta_thresh, int stage, int predicated, int sweep)
{
    if (ctl->abort != 0)
        return;  // sweep mode: an earlier step handed control back to the host
    const int lane = threadIdx.x;
    const int i = ctl->i, j = i + 1;
    if (predicated && ctl->need_corr == 0)
    {
        // the speculative correction pass was skipped (it returned in its prologue)
        if (sweep && lane == 0)
        {
            ctl->acc_skipped += 1;
            ctl->acc_skipped_cols += j;
        }
        return;
    }
    if (stage == 0)
    {
        // h = V^T w -> H(0:i, i)  (Arnoldi.h:249-251); it is also the coefficient vector of f = w - V h
        for (int k = lane; k < j; k += 32)
        {
            const double h = ctl->red[k];
            H[k + (int64_t) i * m] = h;
            ctl->c[k] = h;
        }
        if (lane == 0)
        {
            double s = 0.0;
            for (int k = 0; k < j; k++)
                s += ctl->red[k] * ctl->red[k];
            ctl->hnorm = sqrt(s);  // ||h||   (Arnoldi.h:257)
        }
        return;
    }
    int count = ctl->count;
    if (stage == 2)
    {
        for (int k = lane; k < j; k += 32)
            H[k + (int64_t) i * m] += ctl->c[k];  // h += Vf   (Arnoldi.h:283)
        count += 1;
    }
    double beta = sqrt(ctl->red[kRedNrm]);
    const double ortho_err = panel_max_abs(ctl->red, j, lane);
    __syncwarp();
    int need, skip = 0, zeroed = 0;
    if (stage == 1 && beta > 0.717 * ctl->hnorm)  // DGKS   (Arnoldi.h:257)
    {
        skip = 1;
        need = 0;
    }
    else
    {
        for (int k = lane; k < j; k += 32)
            ctl->c[k] = ctl->red[k];
        need = (count < 5) && (ortho_err > kEps * beta);  // Arnoldi.h:266
        if (need && beta < beta_thresh)                    // Arnoldi.h:273-278
        {
            zeroed = 1;
            beta = 0.0;
            need = 0;
        }
    }
    if (lane == 0)
    {
        ctl->beta = beta;
        ctl->ortho_err = ortho_err;
        ctl->count = count;
        ctl->need_corr = need;
        ctl->f_zeroed = zeroed;
        ctl->dgks_skip = skip;
        if (sweep)
        {
            // stage 1 with need set: the speculatively enqueued pass and its stage-2 decision follow on the device.  Otherwise the step is
            // complete, unless the host is needed: a third pass, a residual to zero, or beta below near_0 (expand_basis at the next step).
            const bool host = (stage == 2 && need) || zeroed || !(beta >= kNear0);
            if (host)
                ctl->abort = 1;
            else if (!need)
                ctl->acc_count += count;
        }
    }
}

// Complex flavour of arnoldi_decide_kernel (Arnoldi.h:242-290 with Scalar = std::complex<double>): ctl->red / ctl->c carry Re at [k]
// and Im at [kRedNrm + 1 + k]; column i of the Hessenberg matrix is kept as H (real parts) and Hi (imaginary parts).
__global__ void arnoldi_decide_z_kernel(FacCtl* ctl, double* H, double* Hi, int m, double beta_thresh, int stage, int predicated)
{
    if (predicated && ctl->need_corr == 0)
        return;
    const int lane = threadIdx.x;
    const int i = ctl->i, j = i + 1;
    if (stage == 0)
    {
        for (int k = lane; k < j; k += 32)
        {
            const double hr = ctl->red[k], hi = ctl->red[kRedNrm + 1 + k];
            H[k + (int64_t) i * m] = hr;
            Hi[k + (int64_t) i * m] = hi;
            ctl->c[k] = hr;
            ctl->c[kRedNrm + 1 + k] = hi;
        }
        if (lane == 0)
        {
            double s = 0.0;
            for (int k = 0; k < j; k++)
                s += ctl->red[k] * ctl->red[k] + ctl->red[kRedNrm + 1 + k] * ctl->red[kRedNrm + 1 + k];
            ctl->hnorm = sqrt(s);  // ||h||   (Arnoldi.h:257)
        }
        return;
    }
    int count = ctl->count;
    if (stage == 2)
    {
        for (int k = lane; k < j; k += 32)
        {
            H[k + (int64_t) i * m] += ctl->c[k];  // h += Vf   (Arnoldi.h:283)
            Hi[k + (int64_t) i * m] += ctl->c[kRedNrm + 1 + k];
        }
        count += 1;
    }
    double beta = sqrt(ctl->red[kRedNrm]);
    double mx = 0.0;
    for (int k = lane; k < j; k += 32)
        mx = fmax(mx, hypot(ctl->red[k], ctl->red[kRedNrm + 1 + k]));
    const double ortho_err = warp_max(mx);
    __syncwarp();
    int need, skip = 0, zeroed = 0;
    if (stage == 1 && beta > 0.717 * ctl->hnorm)  // DGKS   (Arnoldi.h:257)
    {
        skip = 1;
        need = 0;
    }
    else
    {
        for (int k = lane; k < j; k += 32)
        {
            ctl->c[k] = ctl->red[k];
            ctl->c[kRedNrm + 1 + k] = ctl->red[kRedNrm + 1 + k];
        }
        need = (count < 5) && (ortho_err > kEps * beta);  // Arnoldi.h:266
        if (need && beta < beta_thresh)                    // Arnoldi.h:273-278
        {
            zeroed = 1;
            beta = 0.0;
            need = 0;
        }
    }
    if (lane == 0)
    {
        ctl->beta = beta;
        ctl->ortho_err = ortho_err;
        ctl->count = count;
        ctl->need_corr = need;
        ctl->f_zeroed = zeroed;
        ctl->dgks_skip = skip;
    }
}

// ---------------------------------------------------------------------------------------------
// Restart GEMM: one thread per row, row of V in registers (MP = padded panel width), Q in smem.
// ---------------------------------------------------------------------------------------------
constexpr int kGemmBlock = 128;

template <int MP>
__global__ void __launch_bounds__(kGemmBlock)
    compress_kernel(const double* V, int64_t ldv, int64_t nrows, int m, const double* __restrict__ Q, int kk, double* Vout, int64_t ldo, double* f,
                    const double* __restrict__ H, double* red_out, double* partials, unsigned int* ticket)
{
    extern __shared__ double s_q[];  // MP x kk, column-major, rows >= m zero
    for (int t = threadIdx.x; t < MP * kk; t += kGemmBlock)
    {
        const int r = t % MP, c = t / MP;
        s_q[t] = (r < m) ? Q[r + (int64_t) c * m] : 0.0;
    }
    __syncthreads();
    const uint64_t pol = l2_policy_evict_first();
    double fq = 0.0, fh = 0.0;
    if (f)
    {
        fq = Q[(m - 1) + (int64_t) (kk - 2) * m];   // Q(m-1, k-1)
        fh = H[(kk - 1) + (int64_t) (kk - 2) * m];  // H(k, k-1)
    }
    double nrm = 0.0;
    for (int64_t r = (int64_t) blockIdx.x * kGemmBlock + threadIdx.x; r < nrows; r += (int64_t) gridDim.x * kGemmBlock)
    {
        double v[MP];
#pragma unroll
        for (int jx = 0; jx < MP; jx++)
            v[jx] = (jx < m) ? ld_stream_f64(V + r + (int64_t) jx * ldv, pol) : 0.0;
        double last = 0.0;
        for (int c = 0; c < kk; c++)
        {
            const double* q = s_q + c * MP;
            double s0 = 0.0, s1 = 0.0;
#pragma unroll
            for (int jx = 0; jx < MP; jx += 2)
            {
                const double2 qq = *reinterpret_cast<const double2*>(q + jx);
                s0 = fma(v[jx], qq.x, s0);
                s1 = fma(v[jx + 1], qq.y, s1);
            }
            last = s0 + s1;
            Vout[r + (int64_t) c * ldo] = last;
        }
        if (f)
        {
            const double fn = f[r] * fq + last * fh;  // Arnoldi.h:337
            f[r] = fn;
            nrm = fma(fn, fn, nrm);
        }
    }
    if (f)
    {
        __shared__ double s_w[kGemmBlock / 32];
        nrm = warp_sum(nrm);
        if ((threadIdx.x & 31) == 0)
            s_w[threadIdx.x >> 5] = nrm;
        __syncthreads();
        double cta = 0.0;
        if (threadIdx.x == 0)
            for (int q = 0; q < kGemmBlock / 32; q++)
                cta += s_w[q];
        grid_reduce_fixed_order<kGemmBlock>(cta, 1, partials, ticket, red_out);
    }
}

}  // namespace

void launch_panel_pass(int mode, const double* V, int64_t ldv, int64_t nrows, int j, const double* x, double* f_out, const double* coef, double* red_out,
                       const RedScratch& rs, cudaStream_t stream, const int* pred, bool cplx, const PeerX* push_or_null, const int* abort, int64_t row_limit)
{
    if (row_limit < 0)
        row_limit = ldv;
    const PeerX none{};
    const PeerX& push = (push_or_null && mode == PANEL_CORR && !cplx) ? *push_or_null : none;
    SB200_REQUIRE(j >= 1 && j <= kPanelMaxCols, SB200_INVALID_ARGUMENT, "panel width must be in [1, 64]");
    const int sms = device_info().sm_count;
    if (cplx)
    {
        // nrows / ldv count doubles (two per complex entry); see panel_kernel
        SB200_REQUIRE(j < kPanelMaxCols, SB200_INVALID_ARGUMENT, "complex panel width must be in [1, 63]");
        const int ngz = j <= 8 ? 1 : (j <= 16 ? 2 : (j <= 32 ? 4 : 8));
        const int64_t rpiz = (8 / ngz) * 64;
        const int gridz = (int) std::max<int64_t>(1, std::min<int64_t>((nrows + rpiz - 1) / rpiz, (int64_t) sms * 2));
        SB200_REQUIRE(gridz <= rs.max_grid, SB200_LOGIC, "panel: reduction scratch too small");
        switch (mode)
        {
            case PANEL_DOT: launch_panel_mode_z<PANEL_DOT>(V, ldv, nrows, j, x, f_out, coef, red_out, gridz, rs, pred, stream); break;
            case PANEL_FORM: launch_panel_mode_z<PANEL_FORM>(V, ldv, nrows, j, x, f_out, coef, red_out, gridz, rs, pred, stream); break;
            case PANEL_CORR: launch_panel_mode_z<PANEL_CORR>(V, ldv, nrows, j, x, f_out, coef, red_out, gridz, rs, pred, stream); break;
            default: throw Error(SB200_LOGIC, "bad panel mode");
        }
        SB200_CUDA_CHECK(cudaGetLastError());
        return;
    }
    const int ng = j <= 16 ? 1 : (j <= 32 ? 2 : 4);
    const int64_t rpi = (8 / ng) * 64;
    const int64_t need = (nrows + rpi - 1) / rpi;
    const int grid = (int) std::max<int64_t>(1, std::min<int64_t>(need, (int64_t) sms * 2));
    SB200_REQUIRE(grid <= rs.max_grid, SB200_LOGIC, "panel: reduction scratch too small");
    switch (mode)
    {
        case PANEL_DOT: launch_panel_mode<PANEL_DOT>(V, ldv, nrows, j, x, f_out, coef, red_out, grid, rs, pred, stream, none, abort, row_limit); break;
        case PANEL_FORM: launch_panel_mode<PANEL_FORM>(V, ldv, nrows, j, x, f_out, coef, red_out, grid, rs, pred, stream, none, abort, row_limit); break;
        case PANEL_CORR: launch_panel_mode<PANEL_CORR>(V, ldv, nrows, j, x, f_out, coef, red_out, grid, rs, pred, stream, push, abort, row_limit); break;
        default: throw Error(SB200_LOGIC, "bad panel mode");
    }
    SB200_CUDA_CHECK(cudaGetLastError());
}

void launch_lanczos_decide(FacCtl* ctl, double* H, int m, double beta_thresh, int stage, cudaStream_t stream, int predicated, bool cplx, int sweep, int two_part)
{
    if (cplx)
        lanczos_decide_z_kernel<<<1, 32, 0, stream>>>(ctl, H, m, beta_thresh, stage, predicated);
    else
        lanczos_decide_kernel<<<1, 32, 0, stream>>>(ctl, H, m, beta_thresh, stage, predicated, sweep, two_part);
    SB200_CUDA_CHECK(cudaGetLastError());
}

void launch_lanczos_decide_peer(const PeerCtl& pc, int count, FacCtl* ctl, double* H, int m, double beta_thresh, int stage, cudaStream_t stream, int sweep, int two_part)
{
#ifdef SB200_EMU
    // kernel-logic emulator: a kernel that spins on a peer cannot run under the serialised launches (see launch_peer_allreduce)
    launch_peer_allreduce(pc, ctl->red, count, 0, stream, &ctl->abort);
    launch_lanczos_decide(ctl, H, m, beta_thresh, stage, stream, 0, false, sweep, two_part);
#else
    lanczos_decide_peer_kernel<<<1, 128, 0, stream>>>(pc, count, ctl, H, m, beta_thresh, stage, sweep, two_part);
    SB200_CUDA_CHECK(cudaGetLastError());
#endif
}

void launch_arnoldi_decide(FacCtl* ctl, double* H, int m, double beta_thresh, int stage, cudaStream_t stream, int predicated, double* Hi, int sweep)
{
    if (Hi)
        arnoldi_decide_z_kernel<<<1, 32, 0, stream>>>(ctl, H, Hi, m, beta_thresh, stage, predicated);
    else
        arnoldi_decide_kernel<<<1, 32, 0, stream>>>(ctl, H, m, beta_thresh, stage, predicated, sweep);
    SB200_CUDA_CHECK(cudaGetLastError());
}

void launch_compress(const double* V, int64_t ldv, int64_t nrows, int m, const double* Q, int kk, double* Vout, int64_t ldo, double* f, const double* H,
                     double* red_out, const RedScratch& rs, cudaStream_t stream)
{
    SB200_REQUIRE(m >= 1 && m <= kPanelMaxCols && kk >= 1 && kk <= m, SB200_INVALID_ARGUMENT, "compress: bad dimensions");
    SB200_REQUIRE(!f || kk >= 2, SB200_INVALID_ARGUMENT, "compress: residual update needs k >= 1");
    // default: DMMA + TMA kernel (gemm_dmma.cu); SB200_COMPRESS_FMA=1 selects the plain FMA kernel below (A/B comparison)
    static const bool force_fma = [] { const char* e = std::getenv("SB200_COMPRESS_FMA"); return e && e[0] == '1'; }();
    if (!force_fma && ldv % 64 == 0 && ldo % 64 == 0)
    {
        launch_compress_dmma(V, ldv, nrows, m, Q, kk, Vout, ldo, f, H, red_out, rs, stream);
        return;
    }
    launch_compress_fma(V, ldv, nrows, m, Q, kk, Vout, ldo, f, H, red_out, rs, stream);
}

void launch_compress_fma(const double* V, int64_t ldv, int64_t nrows, int m, const double* Q, int kk, double* Vout, int64_t ldo, double* f, const double* H,
                         double* red_out, const RedScratch& rs, cudaStream_t stream)
{
    const int sms = device_info().sm_count;
    const int64_t need = (nrows + kGemmBlock - 1) / kGemmBlock;
    const int grid = (int) std::max<int64_t>(1, std::min<int64_t>(need, (int64_t) sms * 3));
    SB200_REQUIRE(grid <= rs.max_grid, SB200_LOGIC, "compress: reduction scratch too small");
    const int mp = m <= 16 ? 16 : (m <= 32 ? 32 : (m <= 48 ? 48 : 64));
    const size_t smem = (size_t) mp * kk * sizeof(double);
    switch (mp)
    {
        case 16: compress_kernel<16><<<grid, kGemmBlock, smem, stream>>>(V, ldv, nrows, m, Q, kk, Vout, ldo, f, H, red_out, rs.partials, rs.ticket); break;
        case 32: compress_kernel<32><<<grid, kGemmBlock, smem, stream>>>(V, ldv, nrows, m, Q, kk, Vout, ldo, f, H, red_out, rs.partials, rs.ticket); break;
        case 48: compress_kernel<48><<<grid, kGemmBlock, smem, stream>>>(V, ldv, nrows, m, Q, kk, Vout, ldo, f, H, red_out, rs.partials, rs.ticket); break;
        default: compress_kernel<64><<<grid, kGemmBlock, smem, stream>>>(V, ldv, nrows, m, Q, kk, Vout, ldo, f, H, red_out, rs.partials, rs.ticket); break;
    }
    SB200_CUDA_CHECK(cudaGetLastError());
}

// test hook: host buffers in / out, one launch of the selected product kernel
void dense_compress_host(int64_t n, int m, int kk, const double* V, const double* Q, const double* H, double* Vout, double* f, double* fnorm2, int impl)
{
    const int64_t ld = round_up(std::max<int64_t>(n, 2), 64);
    DevBuf<double> dV((size_t) ld * m), dQ((size_t) m * m), dH((size_t) m * m), dO((size_t) ld * kk), df((size_t) ld), dred(8);
    const int max_grid = device_info().sm_count * 8;
    DevBuf<double> partials((size_t) max_grid * 128);
    DevBuf<unsigned int> ticket(1);
    SB200_CUDA_CHECK(cudaMemset(dV.get(), 0, sizeof(double) * (size_t) ld * m));
    SB200_CUDA_CHECK(cudaMemset(dO.get(), 0, sizeof(double) * (size_t) ld * kk));
    SB200_CUDA_CHECK(cudaMemset(df.get(), 0, sizeof(double) * (size_t) ld));
    SB200_CUDA_CHECK(cudaMemset(ticket.get(), 0, sizeof(unsigned int)));
    SB200_CUDA_CHECK(cudaMemcpy2D(dV.get(), sizeof(double) * ld, V, sizeof(double) * n, sizeof(double) * n, m, cudaMemcpyHostToDevice));
    SB200_CUDA_CHECK(cudaMemcpy(dQ.get(), Q, sizeof(double) * (size_t) m * m, cudaMemcpyHostToDevice));
    if (f)
    {
        SB200_CUDA_CHECK(cudaMemcpy(dH.get(), H, sizeof(double) * (size_t) m * m, cudaMemcpyHostToDevice));
        SB200_CUDA_CHECK(cudaMemcpy(df.get(), f, sizeof(double) * n, cudaMemcpyHostToDevice));
    }
    RedScratch rs{partials.get(), ticket.get(), max_grid};
    if (impl == 0)
        launch_compress_dmma(dV.get(), ld, n, m, dQ.get(), kk, dO.get(), ld, f ? df.get() : nullptr, dH.get(), dred.get(), rs, nullptr);
    else
        launch_compress_fma(dV.get(), ld, n, m, dQ.get(), kk, dO.get(), ld, f ? df.get() : nullptr, dH.get(), dred.get(), rs, nullptr);
    SB200_CUDA_CHECK(cudaDeviceSynchronize());
    SB200_CUDA_CHECK(cudaMemcpy2D(Vout, sizeof(double) * n, dO.get(), sizeof(double) * ld, sizeof(double) * n, kk, cudaMemcpyDeviceToHost));
    if (f)
    {
        SB200_CUDA_CHECK(cudaMemcpy(f, df.get(), sizeof(double) * n, cudaMemcpyDeviceToHost));
        SB200_CUDA_CHECK(cudaMemcpy(fnorm2, dred.get(), sizeof(double), cudaMemcpyDeviceToHost));
    }
}

}  // namespace sb200

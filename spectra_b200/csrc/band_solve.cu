// Device shift-solve operator  y = (A - sigma I)^{-1} x  for banded A          
// Replaces SparseSymShiftSolve (MatOp/SparseSymShiftSolve.h:85-109: set_shift = Eigen::SparseLU factorisation,
// perform_op = solve) for the matrix class of BASELINE config 5 (symmetric banded, half-bandwidth <= 32).
//
// B200 design.  A sequential band LU (what a CPU does, ~n dependent steps) would leave 147 SMs idle, so the
// factorisation is block cyclic reduction (BCR) on the block-tridiagonal form of A - sigma I with B x B blocks,
// B >= half-bandwidth: log2(n/B) levels, every level embarrassingly parallel over block rows, one warp per block row.
//
//   level l (stride s = 2^l), active rows i = k s - 1 (k = 1..floor(N/s)); odd k are eliminated, even k kept:
//     eliminated j:  Dinv_j = D_j^{-1} (Gauss-Jordan, partial pivoting inside the block), GL_j = Dinv_j L_j, GU_j = Dinv_j U_j
//     kept i (p = i - s, q = i + s):  ML_i = L_i Dinv_p, MU_i = U_i Dinv_q,
//                                     D_i <- D_i - L_i GU_p - U_i GL_q,  L_i <- -L_i GL_p,  U_i <- -U_i GU_q
//   solve:  forward  l = 0..L-1 : f_i -= ML_i f_p + MU_i f_q            (kept rows)
//           top               : x_t = Dinv_t f_t
//           backward l = L-1..0 : x_j = Dinv_j f_j - GL_j x_{j-s} - GU_j x_{j+s}   (eliminated rows)
//   A solve is ~5 N B^2 doubles of streamed factors (120 MB at n = 2e5, B = 15: L2 resident on B200) in 2 L + 1 parallel
//   sweeps; the levels with <= 64 active rows run inside one CTA.  One step of iterative refinement with the CSR
//   operator (r = x - (A - sigma I) y) removes the growth BCR can show on an indefinite shifted matrix.
//
// Wide bands (half-bandwidth b > 32, e.g. 2-D / 3-D stencil matrices in their natural ordering: b = nx, nx*ny): with B = b the
// block-tridiagonal form has few (N = n / b) but large block rows, and the parallelism is INSIDE the block operations, so the
// factorisation is the sequential block-tridiagonal ("block Thomas") elimination with grid-wide kernels per block row:
//     S_i = D_i - ML_i U_{i-1},  ML_i = L_i S_{i-1}^{-1},  Sinv_i = S_i^{-1} (Gauss-Jordan, partial pivoting inside the block),  GU_i = Sinv_i U_i
//     solve:  g_i = f_i - ML_i g_{i-1}  (forward),   x_i = Sinv_i g_i - GU_i x_{i+1}  (backward)
//   The off-diagonal blocks are never densified: L_i and U_i = L_{i+1}^T are read from the CSR, so the three products cost
//   O(B^2 nnz/row) and only the inverse is O(B^3).  Stored factors: ML, Sinv, GU = 3 N B^2 doubles (16.6 GB for a 27-point stencil on
//   58^3 points, n = 2e5, b = 3423), streamed once per solve.  Same refinement step, same verification solve, same failure rule.
//
// Pivoting is confined to the diagonal blocks, so a shift for which some reduced diagonal block is singular is
// rejected by set_shift (std::invalid_argument in the shim, like the reference's "factorization failed with the given
// shift"); set_shift also verifies the factorisation with one random solve.
#include <cmath>
#include <cstdlib>
#include <string>

#include "host.h"

namespace sb200 {

namespace {

constexpr int kTopRows = 64;  // levels with at most this many active rows are fused into one CTA
constexpr int kTopThreads = 512;

__global__ void band_width_kernel(const int* __restrict__ rowptr, const int* __restrict__ col, int64_t nrows, int64_t row0, int* out_bw)
{
    int bw = 0;
    for (int64_t r = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += (int64_t) gridDim.x * blockDim.x)
    {
        const int64_t i = row0 + r;
        for (int p = rowptr[r]; p < rowptr[r + 1]; p++)
        {
            const int64_t d = i - col[p];
            bw = max(bw, (int) (d < 0 ? -d : d));
        }
    }
    if (bw > 0)
        atomicMax(out_bw, bw);
}

// blocks of A - sigma I from the CSR: D (diagonal), Lo (coupling to block row i-1), Up (to i+1); column-major B x B each
__global__ void band_scatter_kernel(const int* __restrict__ rowptr, const int* __restrict__ col, const double* __restrict__ val, int64_t n, int B, double* D,
                                    double* Lo, double* Up, int* flag)
{
    const int64_t bb = (int64_t) B * B;
    for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t) gridDim.x * blockDim.x)
    {
        const int64_t bi = i / B;
        const int ri = (int) (i % B);
        for (int p = rowptr[i]; p < rowptr[i + 1]; p++)
        {
            const int64_t j = col[p];
            const int64_t bj = j / B;
            const int cj = (int) (j % B);
            double* dst = bj == bi ? D : (bj == bi - 1 ? Lo : (bj == bi + 1 ? Up : nullptr));
            if (!dst)
            {
                *flag = 2;
                continue;
            }
            atomicAdd(dst + bi * bb + ri + (int64_t) cj * B, val[p]);  // column blocks of the operator may split a row
        }
    }
}

__global__ void band_diag_kernel(double* D, int64_t n, int64_t N, int B, double sigma)
{
    const int64_t bb = (int64_t) B * B;
    for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < N * B; i += (int64_t) gridDim.x * blockDim.x)
    {
        const int64_t bi = i / B;
        const int ri = (int) (i % B);
        double* d = D + bi * bb + ri + (int64_t) ri * B;
        if (i < n)
            *d -= sigma;
        else
            *d = 1.0;  // identity padding of the last block
    }
}

// element (r, c) of A * Bm, both B x B column-major in shared memory
__device__ __forceinline__ double row_dot(const double* A, const double* Bm, int B, int r, int c)
{
    double acc = 0.0;
    for (int k = 0; k < B; k++)
        acc = fma(A[r + k * B], Bm[k + c * B], acc);
    return acc;
}

__device__ __forceinline__ void load_block(double* dst_smem, const double* src, int bb, int lane)
{
    for (int e = lane; e < bb; e += 32)
        dst_smem[e] = src[e];
    __syncwarp();
}

// ---- factor, eliminated rows of one level: one warp (= one CTA) per row ----
__global__ void __launch_bounds__(32) bcr_eliminate_kernel(const double* __restrict__ D, const double* __restrict__ Lo, const double* __restrict__ Up, double* Dinv,
                                                           double* GL, double* GU, int64_t N, int B, int level, int64_t count, int* flag)
{
    extern __shared__ double sm[];
    const int bb = B * B;
    double* M = sm;            // working copy of D_j
    double* Inv = sm + bb;     // becomes D_j^{-1}
    double* T = sm + 2 * bb;   // operand
    const int lane = threadIdx.x;
    const int64_t s = (int64_t) 1 << level;
    const int64_t k = 2 * (int64_t) blockIdx.x + 1;  // odd k
    if (k > count)
        return;
    const int64_t j = k * s - 1;
    (void) N;
    load_block(M, D + j * bb, bb, lane);
    for (int e = lane; e < bb; e += 32)
        Inv[e] = (e % B == e / B) ? 1.0 : 0.0;
    __syncwarp();
    for (int kk = 0; kk < B; kk++)
    {
        // partial pivoting inside the block
        double best = (lane >= kk && lane < B) ? fabs(M[lane + kk * B]) : -1.0;
        int arg = lane;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
        {
            const double ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
            if (ob > best || (ob == best && oa < arg))
            {
                best = ob;
                arg = oa;
            }
        }
        if (!(best > 0.0))
        {
            if (lane == 0)
                *flag = 1;  // singular (or NaN) reduced diagonal block
            return;
        }
        if (arg != kk && lane < B)
        {
            double t = M[kk + lane * B];
            M[kk + lane * B] = M[arg + lane * B];
            M[arg + lane * B] = t;
            t = Inv[kk + lane * B];
            Inv[kk + lane * B] = Inv[arg + lane * B];
            Inv[arg + lane * B] = t;
        }
        __syncwarp();
        const double piv = M[kk + kk * B];
        __syncwarp();
        if (lane < B)
        {
            M[kk + lane * B] /= piv;
            Inv[kk + lane * B] /= piv;
        }
        __syncwarp();
        if (lane < B && lane != kk)
        {
            const double fct = M[lane + kk * B];
            if (fct != 0.0)
                for (int c = 0; c < B; c++)
                {
                    M[lane + c * B] = fma(-fct, M[kk + c * B], M[lane + c * B]);
                    Inv[lane + c * B] = fma(-fct, Inv[kk + c * B], Inv[lane + c * B]);
                }
        }
        __syncwarp();
    }
    for (int e = lane; e < bb; e += 32)
        Dinv[j * bb + e] = Inv[e];
    // GL = Dinv * L_j, GU = Dinv * U_j
    load_block(T, Lo + j * bb, bb, lane);
    if (lane < B)
        for (int c = 0; c < B; c++)
            GL[j * bb + lane + c * B] = row_dot(Inv, T, B, lane, c);
    __syncwarp();
    load_block(T, Up + j * bb, bb, lane);
    if (lane < B)
        for (int c = 0; c < B; c++)
            GU[j * bb + lane + c * B] = row_dot(Inv, T, B, lane, c);
}

// ---- factor, kept rows of one level ----
__global__ void __launch_bounds__(32) bcr_update_kernel(double* D, double* Lo, double* Up, const double* __restrict__ Dinv, const double* __restrict__ GL,
                                                        const double* __restrict__ GU, double* ML, double* MU, int64_t N, int B, int level, int64_t count)
{
    extern __shared__ double sm[];
    const int bb = B * B;
    double* Ls = sm;
    double* Us = sm + bb;
    double* T = sm + 2 * bb;
    const int lane = threadIdx.x;
    const int64_t s = (int64_t) 1 << level;
    const int64_t k = 2 * ((int64_t) blockIdx.x + 1);  // even k
    if (k > count)
        return;
    const int64_t i = k * s - 1, p = i - s, q = i + s;
    const bool has_q = q < N;
    double* ml = ML + (int64_t) blockIdx.x * bb;
    double* mu = MU + (int64_t) blockIdx.x * bb;
    load_block(Ls, Lo + i * bb, bb, lane);
    load_block(Us, Up + i * bb, bb, lane);
    double* Di = D + i * bb;
    // --- left neighbour p ---
    load_block(T, Dinv + p * bb, bb, lane);
    if (lane < B)
        for (int c = 0; c < B; c++)
            ml[lane + c * B] = row_dot(Ls, T, B, lane, c);
    __syncwarp();
    load_block(T, GU + p * bb, bb, lane);
    if (lane < B)
        for (int c = 0; c < B; c++)
            Di[lane + c * B] -= row_dot(Ls, T, B, lane, c);
    __syncwarp();
    load_block(T, GL + p * bb, bb, lane);
    if (lane < B)
        for (int c = 0; c < B; c++)
            Lo[i * bb + lane + c * B] = -row_dot(Ls, T, B, lane, c);
    __syncwarp();
    // --- right neighbour q ---
    if (has_q)
    {
        load_block(T, Dinv + q * bb, bb, lane);
        if (lane < B)
            for (int c = 0; c < B; c++)
                mu[lane + c * B] = row_dot(Us, T, B, lane, c);
        __syncwarp();
        load_block(T, GL + q * bb, bb, lane);
        if (lane < B)
            for (int c = 0; c < B; c++)
                Di[lane + c * B] -= row_dot(Us, T, B, lane, c);
        __syncwarp();
        load_block(T, GU + q * bb, bb, lane);
        if (lane < B)
            for (int c = 0; c < B; c++)
                Up[i * bb + lane + c * B] = -row_dot(Us, T, B, lane, c);
    }
    else
    {
        for (int e = lane; e < bb; e += 32)
        {
            mu[e] = 0.0;
            Up[i * bb + e] = 0.0;
        }
    }
}

// y_r = sum_c M[r + c B] * v[c], v held one entry per lane.  All (<= BMAX) loads of the lane's row are issued before
// the first use so that one block row costs one memory latency, not B of them.
template <int BMAX>
__device__ __forceinline__ double mat_vec_lane(const double* __restrict__ M, int B, int lane, double v_lane)
{
    double m[BMAX];
#pragma unroll
    for (int c = 0; c < BMAX; c++)
        m[c] = (c < B && lane < B) ? __ldg(M + lane + c * B) : 0.0;
    double acc = 0.0;
#pragma unroll
    for (int c = 0; c < BMAX; c++)
        acc = fma(m[c], __shfl_sync(0xffffffffu, v_lane, c), acc);
    return acc;
}

struct BcrView
{
    const double* Dinv;
    const double* GL;
    const double* GU;
    const double* ML;  // level-major: rows kept at level l start at ml_off[l] blocks
    const double* MU;
    int64_t N;
    int B;
    int levels;             // L: number of elimination levels; the top row is 2^L - 1
    int64_t ml_off[40];
};

template <int BMAX>
__device__ __forceinline__ void forward_row(const BcrView& v, double* f, int level, int64_t idx, int lane)
{
    const int bb = v.B * v.B;
    const int64_t s = (int64_t) 1 << level;
    const int64_t k = 2 * (idx + 1);
    const int64_t i = k * s - 1, p = i - s, q = i + s;
    const double* ml = v.ML + (v.ml_off[level] + idx) * bb;
    const double* mu = v.MU + (v.ml_off[level] + idx) * bb;
    const double fp = lane < v.B ? f[p * v.B + lane] : 0.0;
    double acc = mat_vec_lane<BMAX>(ml, v.B, lane, fp);
    if (q < v.N)
    {
        const double fq = lane < v.B ? f[q * v.B + lane] : 0.0;
        acc += mat_vec_lane<BMAX>(mu, v.B, lane, fq);
    }
    if (lane < v.B)
        f[i * v.B + lane] -= acc;
}

template <int BMAX>
__device__ __forceinline__ void backward_row(const BcrView& v, double* f, int level, int64_t idx, int lane)
{
    const int bb = v.B * v.B;
    const int64_t s = (int64_t) 1 << level;
    const int64_t k = 2 * idx + 1;
    const int64_t j = k * s - 1, p = j - s, q = j + s;
    const double fj = lane < v.B ? f[j * v.B + lane] : 0.0;
    double acc = mat_vec_lane<BMAX>(v.Dinv + j * bb, v.B, lane, fj);
    if (p >= 0)
    {
        const double xp = lane < v.B ? f[p * v.B + lane] : 0.0;
        acc -= mat_vec_lane<BMAX>(v.GL + j * bb, v.B, lane, xp);
    }
    if (q < v.N)
    {
        const double xq = lane < v.B ? f[q * v.B + lane] : 0.0;
        acc -= mat_vec_lane<BMAX>(v.GU + j * bb, v.B, lane, xq);
    }
    if (lane < v.B)
        f[j * v.B + lane] = acc;
}

template <int BMAX>
__global__ void __launch_bounds__(128) bcr_forward_kernel(BcrView v, double* f, int level, int64_t nkept)
{
    const int lane = threadIdx.x & 31;
    const int64_t w = (int64_t) blockIdx.x * 4 + (threadIdx.x >> 5);
    if (w < nkept)
        forward_row<BMAX>(v, f, level, w, lane);
}

template <int BMAX>
__global__ void __launch_bounds__(128) bcr_backward_kernel(BcrView v, double* f, int level, int64_t nelim)
{
    const int lane = threadIdx.x & 31;
    const int64_t w = (int64_t) blockIdx.x * 4 + (threadIdx.x >> 5);
    if (w < nelim)
        backward_row<BMAX>(v, f, level, w, lane);
}

// levels [first, L) forward, the top row, then backward L-1 .. first, inside one CTA
template <int BMAX>
__global__ void __launch_bounds__(kTopThreads) bcr_top_kernel(BcrView v, double* f, int first)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = kTopThreads / 32;
    for (int l = first; l < v.levels; l++)
    {
        const int64_t cnt = v.N >> l, nkept = cnt / 2;
        for (int64_t w = warp; w < nkept; w += nw)
            forward_row<BMAX>(v, f, l, w, lane);
        __syncthreads();
    }
    if (warp == 0)
    {
        const int64_t t = ((int64_t) 1 << v.levels) - 1;
        const double ft = lane < v.B ? f[t * v.B + lane] : 0.0;
        const double xt = mat_vec_lane<BMAX>(v.Dinv + t * v.B * v.B, v.B, lane, ft);
        if (lane < v.B)
            f[t * v.B + lane] = xt;
    }
    __syncthreads();
    for (int l = v.levels - 1; l >= first; l--)
    {
        const int64_t cnt = v.N >> l, nelim = (cnt + 1) / 2;
        for (int64_t w = warp; w < nelim; w += nw)
            backward_row<BMAX>(v, f, l, w, lane);
        __syncthreads();
    }
}

__global__ void pad_copy_kernel(const double* __restrict__ x, double* __restrict__ xb, int64_t n, int64_t npad)
{
    for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < npad; i += (int64_t) gridDim.x * blockDim.x)
        xb[i] = i < n ? x[i] : 0.0;
}

// r = x - (t - sigma y)   with t = A y
__global__ void refine_residual_kernel(const double* __restrict__ x, const double* __restrict__ t, const double* __restrict__ y, double sigma, double* __restrict__ r,
                                       int64_t n, int64_t npad)
{
    for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < npad; i += (int64_t) gridDim.x * blockDim.x)
        r[i] = i < n ? x[i] - (t[i] - sigma * y[i]) : 0.0;
}

__global__ void add_out_kernel(const double* __restrict__ y0, const double* __restrict__ dy, double* __restrict__ y, int64_t n)
{
    for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t) gridDim.x * blockDim.x)
        y[i] = dy ? y0[i] + dy[i] : y0[i];
}

// ---- small matrices that are not banded (the reference's own fixtures, test/SymEigsShift.cpp:148-186: random sparse
//      n <= 1000): explicit inverse by Gauss-Jordan with partial pivoting, one grid-wide rank-1 update per pivot;
//      the solve is a dense GEMV.  n <= kDenseMax only.
constexpr int kDenseMax = 2048;

__global__ void dense_scatter_kernel(const int* __restrict__ rowptr, const int* __restrict__ col, const double* __restrict__ val, int64_t n, double* M)
{
    for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t) gridDim.x * blockDim.x)
        for (int p = rowptr[i]; p < rowptr[i + 1]; p++)
            atomicAdd(M + i + (int64_t) col[p] * n, val[p]);
}

__global__ void dense_init_kernel(double* M, double* Inv, int64_t n, double sigma)
{
    for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t) gridDim.x * blockDim.x)
    {
        M[i + i * n] -= sigma;
        Inv[i + i * n] = 1.0;
    }
}

// pivot row of column k among rows >= k (largest magnitude, lowest index on ties)
__global__ void __launch_bounds__(1024) gj_pivot_kernel(const double* __restrict__ M, int64_t n, int k, int* piv_row, double* piv_val, int* flag)
{
    __shared__ double s_best[32];
    __shared__ int s_arg[32];
    double best = -1.0;
    int arg = k;
    for (int r = k + threadIdx.x; r < n; r += blockDim.x)
    {
        const double a = fabs(M[r + (int64_t) k * n]);
        if (a > best)
        {
            best = a;
            arg = r;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
        if (ob > best || (ob == best && oa < arg))
        {
            best = ob;
            arg = oa;
        }
    }
    if ((threadIdx.x & 31) == 0)
    {
        s_best[threadIdx.x >> 5] = best;
        s_arg[threadIdx.x >> 5] = arg;
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        for (int w = 1; w < (int) (blockDim.x >> 5); w++)
            if (s_best[w] > best || (s_best[w] == best && s_arg[w] < arg))
            {
                best = s_best[w];
                arg = s_arg[w];
            }
        *piv_row = arg;
        *piv_val = M[arg + (int64_t) k * n];
        if (!(best > 0.0))
            *flag = 1;
    }
}

// swap rows k and p, scale the new row k by 1/pivot, save column k (after the swap) for the update
__global__ void gj_swap_scale_kernel(double* M, double* Inv, int64_t n, int k, const int* piv_row, const double* piv_val, double* colk)
{
    const int p = *piv_row;
    const double piv = *piv_val;
    const double inv_piv = piv != 0.0 ? 1.0 / piv : 0.0;
    for (int64_t c = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; c < 2 * n; c += (int64_t) gridDim.x * blockDim.x)
    {
        double* X = c < n ? M : Inv;
        const int64_t cc = c < n ? c : c - n;
        const double a = X[k + cc * n], b = X[p + cc * n];
        X[p + cc * n] = a;
        X[k + cc * n] = b * inv_piv;
    }
    for (int64_t r = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t) gridDim.x * blockDim.x)
    {
        // column k after the row swap, read from the pre-swap values (this kernel only rewrites rows k and p)
        double v;
        if (r == k)
            v = 0.0;  // row k itself is not updated
        else if (r == p)
            v = M[k + (int64_t) k * n];  // old row k lands in row p; may race with the swap above: see gj_colk_fix
        else
            v = M[r + (int64_t) k * n];
        colk[r] = v;
    }
}

// colk[p] must be the pre-swap M(k,k); it is saved by the pivot step instead of being read during the swap
__global__ void gj_save_kk_kernel(const double* __restrict__ M, int64_t n, int k, double* save) { *save = M[k + (int64_t) k * n]; }

__global__ void gj_fix_colk_kernel(double* colk, int k, const int* piv_row, const double* save)
{
    const int p = *piv_row;
    if (p != k)
        colk[p] = *save;
}

__global__ void gj_eliminate_kernel(double* M, double* Inv, int64_t n, int k, const double* __restrict__ colk)
{
    const int64_t total = 2 * n * n;
    for (int64_t e = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t) gridDim.x * blockDim.x)
    {
        double* X = e < n * n ? M : Inv;
        const int64_t ee = e < n * n ? e : e - n * n;
        const int64_t r = ee % n, c = ee / n;
        const double f = colk[r];
        if (f != 0.0)
            X[ee] = fma(-f, X[k + c * n], X[ee]);
    }
}

// y = Inv * x  (column-major n x n): one warp per 32 rows, lanes along rows
__global__ void __launch_bounds__(256) dense_gemv_kernel(const double* __restrict__ Inv, const double* __restrict__ x, double* __restrict__ y, int64_t n)
{
    __shared__ double s_part[8][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t r = (int64_t) blockIdx.x * 32 + lane;
    double acc = 0.0;
    if (r < n)
        for (int64_t c = warp; c < n; c += 8)
            acc = fma(Inv[r + c * n], x[c], acc);
    s_part[warp][lane] = acc;
    __syncthreads();
    if (warp == 0 && r < n)
    {
        double t = 0.0;
        for (int w = 0; w < 8; w++)
            t += s_part[w][lane];
        y[r] = t;
    }
}

// Gauss-Jordan step kernels of the wide-band route: the dense route's five launches per pivot folded into three (the pivot kernel also
// saves M(k,k); the swap kernel takes colk[p] from that saved value and leaves column k of M alone; the update writes it as e_k)
// orig[r] = original index of the row now at position r; active[j] = 1 once original row j has served as a pivot row -- column j of the
// Inv half of the tableau is then dense, before that it is a unit vector whose row k entry is zero (nothing to update)
__global__ void __launch_bounds__(1024) gj3_pivot_kernel(const double* __restrict__ M, int64_t n, int k, int* piv_row, double* piv_val, int* flag, int* orig,
                                                         int* active)
{
    __shared__ double s_best[32];
    __shared__ int s_arg[32];
    double best = -1.0;
    int arg = k;
    for (int r = k + threadIdx.x; r < n; r += blockDim.x)
    {
        const double a = fabs(M[r + (int64_t) k * n]);
        if (a > best)
        {
            best = a;
            arg = r;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
        if (ob > best || (ob == best && oa < arg))
        {
            best = ob;
            arg = oa;
        }
    }
    if ((threadIdx.x & 31) == 0)
    {
        s_best[threadIdx.x >> 5] = best;
        s_arg[threadIdx.x >> 5] = arg;
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        for (int w = 1; w < (int) (blockDim.x >> 5); w++)
            if (s_best[w] > best || (s_best[w] == best && s_arg[w] < arg))
            {
                best = s_best[w];
                arg = s_arg[w];
            }
        *piv_row = arg;
        piv_val[0] = M[arg + (int64_t) k * n];
        piv_val[1] = M[k + (int64_t) k * n];  // M(k,k) before the row swap
        if (!(best > 0.0))
            *flag = 1;
        const int o = orig[arg];
        orig[arg] = orig[k];
        orig[k] = o;
        active[o] = 1;
    }
}

// swap rows k and p, scale the new row k by 1/pivot, save column k (as it is after the swap) for the update
__global__ void gj3_swap_scale_kernel(double* M, double* Inv, int64_t n, int k, const int* piv_row, const double* piv_val, double* colk)
{
    const int p = *piv_row;
    const double piv = piv_val[0], old_kk = piv_val[1];
    const double inv_piv = piv != 0.0 ? 1.0 / piv : 0.0;
    for (int64_t r = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t) gridDim.x * blockDim.x)
    {
        // rows other than k and p are not written by this kernel; row k is not updated; the old row k lands in row p
        colk[r] = (r == k) ? 0.0 : ((r == p) ? old_kk : M[r + (int64_t) k * n]);
    }
    for (int64_t c = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; c < 2 * n; c += (int64_t) gridDim.x * blockDim.x)
    {
        if (c == k)
            continue;  // column k of M is read by the loop above (other threads) and rewritten as e_k by gj3_eliminate_kernel
        double* X = c < n ? M : Inv;
        const int64_t cc = c < n ? c : c - n;
        const double a = X[k + cc * n], b = X[p + cc * n];
        X[p + cc * n] = a;
        X[k + cc * n] = b * inv_piv;
    }
}
// rank-1 update of the (M | Inv) tableau; column k of M is written directly (it becomes e_k), so nothing reads it after the pivot step
__global__ void gj3_eliminate_kernel(double* M, double* Inv, int64_t n, int k, const double* __restrict__ colk, const int* __restrict__ active)
{
    const int64_t total = 2 * n * n;
    for (int64_t e = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t) gridDim.x * blockDim.x)
    {
        const bool in_m = e < n * n;
        double* X = in_m ? M : Inv;
        const int64_t ee = in_m ? e : e - n * n;
        const int64_t r = ee % n, c = ee / n;
        // structurally finished parts of the tableau are neither read nor written: columns < k of M are unit vectors with M(k, c) = 0,
        // and so are the columns of Inv whose original row has not been a pivot row yet -- about half of the tableau at any pivot
        if (in_m ? c < k : active[c] == 0)
            continue;
        if (in_m && c == k)
        {
            X[ee] = (r == k) ? 1.0 : 0.0;
            continue;
        }
        const double f = colk[r];
        if (f != 0.0)
            X[ee] = fma(-f, X[k + c * n], X[ee]);
    }
}

// ---- blocked Gauss-Jordan inverse of a large diagonal block (wide-band route, B >= 128): pivots are taken in panels of kGjP columns.
// Per panel: (1) partial-pivoting LU of a copy of the panel columns by ONE CTA fixes the kGjP pivot rows; (2) the row swaps are applied to
// the (M | Inv) tableau; (3) the kGjP x kGjP pivot block A_KK is inverted in shared memory; (4) the pivot rows become A_KK^{-1} T(K, :);
// (5) every other row gets T(r, :) -= A_RK(r, :) T(K, :) -- a rank-kGjP update done in 64 x 64 tiles through shared memory, so the tableau is
// streamed once per PANEL instead of once per pivot (1/32 of the traffic of the rank-1 form, ~6 launches per 32 pivots instead of 96).
constexpr int kGjP = 32;

__global__ void __launch_bounds__(1024) gjb_panel_kernel(const double* __restrict__ M, int64_t B, int k0, int pw, double* work, int* piv, int* flag)
{
    __shared__ double s_best[32];
    __shared__ int s_arg[32];
    __shared__ int s_p;
    __shared__ double s_row[kGjP];
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int64_t e = tid; e < (B - k0) * pw; e += nt)
    {
        const int64_t r = k0 + e % (B - k0);
        const int j = (int) (e / (B - k0));
        work[r + (int64_t) j * B] = M[r + (int64_t) (k0 + j) * B];
    }
    __syncthreads();
    for (int k = 0; k < pw; k++)
    {
        double best = -1.0;
        int arg = k0 + k;
        for (int64_t r = k0 + k + tid; r < B; r += nt)
        {
            const double a = fabs(work[r + (int64_t) k * B]);
            if (a > best)
            {
                best = a;
                arg = (int) r;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
        {
            const double ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
            if (ob > best || (ob == best && oa < arg))
            {
                best = ob;
                arg = oa;
            }
        }
        if ((tid & 31) == 0)
        {
            s_best[tid >> 5] = best;
            s_arg[tid >> 5] = arg;
        }
        __syncthreads();
        if (tid == 0)
        {
            for (int w = 1; w < (nt >> 5); w++)
                if (s_best[w] > best || (s_best[w] == best && s_arg[w] < arg))
                {
                    best = s_best[w];
                    arg = s_arg[w];
                }
            s_p = arg;
            piv[k] = arg;
            if (!(best > 0.0))
                *flag = 1;
        }
        __syncthreads();
        const int p = s_p;
        if (tid < pw)
        {
            const double a = work[(k0 + k) + (int64_t) tid * B], b = work[p + (int64_t) tid * B];
            work[p + (int64_t) tid * B] = a;
            work[(k0 + k) + (int64_t) tid * B] = b;
            s_row[tid] = b;  // pivot row k of the panel after the swap
        }
        __syncthreads();
        const double pv = s_row[k];
        const double inv_pv = pv != 0.0 ? 1.0 / pv : 0.0;
        for (int64_t r = k0 + k + 1 + tid; r < B; r += nt)
        {
            const double l = work[r + (int64_t) k * B] * inv_pv;
            if (l != 0.0)
                for (int c = k + 1; c < pw; c++)
                    work[r + (int64_t) c * B] = fma(-l, s_row[c], work[r + (int64_t) c * B]);
        }
        __syncthreads();
    }
}

// the panel's row swaps on every tableau column (in pivot order); block 0 / thread 0 keeps the row bookkeeping of the Inv half
__global__ void gjb_swap_kernel(double* M, double* Inv, int64_t B, int k0, int pw, const int* __restrict__ piv, int* orig, int* active)
{
    for (int64_t c = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; c < 2 * B; c += (int64_t) gridDim.x * blockDim.x)
    {
        double* X = c < B ? M + c * B : Inv + (c - B) * B;
        for (int k = 0; k < pw; k++)
        {
            const int p = piv[k];
            if (p != k0 + k)
            {
                const double a = X[k0 + k];
                X[k0 + k] = X[p];
                X[p] = a;
            }
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (int k = 0; k < pw; k++)
        {
            const int p = piv[k];
            const int o = orig[p];
            orig[p] = orig[k0 + k];
            orig[k0 + k] = o;
            active[o] = 1;
        }
}

// A_KK^{-1} (pw x pw, column-major with leading dimension kGjP) by Gauss-Jordan with partial pivoting in shared memory: one warp
__global__ void __launch_bounds__(32) gjb_kk_kernel(const double* __restrict__ M, int64_t B, int k0, int pw, double* Akk_inv, int* flag)
{
    __shared__ double a[kGjP][kGjP + 1], v[kGjP][kGjP + 1];
    const int lane = threadIdx.x;
    for (int i = 0; i < pw; i++)
    {
        a[i][lane] = lane < pw ? M[(k0 + i) + (int64_t) (k0 + lane) * B] : 0.0;
        v[i][lane] = (i == lane) ? 1.0 : 0.0;
    }
    __syncwarp();
    for (int k = 0; k < pw; k++)
    {
        double best = (lane >= k && lane < pw) ? fabs(a[lane][k]) : -1.0;
        int arg = lane;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
        {
            const double ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
            if (ob > best || (ob == best && oa < arg))
            {
                best = ob;
                arg = oa;
            }
        }
        if (!(best > 0.0))
        {
            if (lane == 0)
                *flag = 1;
            break;
        }
        // lane = column index: swap rows k and arg, scale row k
        {
            const double ak = a[k][lane], ap = a[arg][lane], vk = v[k][lane], vp = v[arg][lane];
            __syncwarp();
            a[arg][lane] = ak;
            v[arg][lane] = vk;
            const double pvt = __shfl_sync(0xffffffffu, ap, k);  // a(arg, k) before the swap = the pivot
            a[k][lane] = ap / pvt;
            v[k][lane] = vp / pvt;
        }
        __syncwarp();
        for (int r = 0; r < pw; r++)
        {
            if (r == k)
                continue;
            const double f = a[r][k];
            __syncwarp();
            if (f != 0.0)
            {
                a[r][lane] = fma(-f, a[k][lane], a[r][lane]);
                v[r][lane] = fma(-f, v[k][lane], v[r][lane]);
            }
            __syncwarp();
        }
    }
    for (int i = 0; i < kGjP; i++)  // the whole kGjP x kGjP array is defined: zero outside the pw x pw block
        Akk_inv[i + lane * kGjP] = (i < pw && lane < pw) ? v[i][lane] : 0.0;
}

// tableau column c takes part in the panel's update: M columns right of the panel, Inv columns whose original row has been a pivot row
__device__ __forceinline__ bool gjb_col_live(int64_t c, int64_t B, int k0, int pw, const int* __restrict__ active)
{
    return c < B ? c >= k0 + pw : (c < 2 * B && active[c - B] != 0);
}

// ARK <- the panel columns of M (all rows; only rows outside the pivot rows are used); pivot rows: T(K, c) <- A_KK^{-1} T(K, c), also kept
// in TK (kGjP x 2B); panel columns of the pivot rows become the identity
__global__ void __launch_bounds__(256) gjb_rows_kernel(double* M, double* Inv, int64_t B, int k0, int pw, const double* __restrict__ Akk_inv, double* ARK, double* TK,
                                                       const int* __restrict__ active)
{
    __shared__ double s_inv[kGjP * kGjP];
    for (int e = threadIdx.x; e < kGjP * kGjP; e += blockDim.x)
        s_inv[e] = Akk_inv[e];
    __syncthreads();
    for (int64_t e = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; e < B * pw; e += (int64_t) gridDim.x * blockDim.x)
    {
        const int64_t r = e % B;
        if (r >= k0 && r < k0 + pw)
            continue;  // pivot rows are rewritten below by other threads; their ARK entries are never read
        ARK[e] = M[r + (int64_t) (k0 + e / B) * B];
    }
    for (int64_t c = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; c < 2 * B; c += (int64_t) gridDim.x * blockDim.x)
    {
        double* X = c < B ? M + c * B : Inv + (c - B) * B;
        if (c >= k0 && c < k0 + pw)
        {
            for (int i = 0; i < pw; i++)
                X[k0 + i] = (c - k0 == i) ? 1.0 : 0.0;
            continue;
        }
        if (!gjb_col_live(c, B, k0, pw, active))
            continue;
        double t[kGjP];
#pragma unroll
        for (int j = 0; j < kGjP; j++)
            t[j] = j < pw ? X[k0 + j] : 0.0;
        for (int i = 0; i < pw; i++)
        {
            double acc = 0.0;
#pragma unroll
            for (int j = 0; j < kGjP; j++)
                acc = fma(s_inv[i + j * kGjP], t[j], acc);
            X[k0 + i] = acc;
            TK[i + c * kGjP] = acc;
        }
    }
}

// T(r, c) -= sum_j ARK(r, j) TK(j, c) for rows outside the pivot rows and live columns: 64 x 64 tiles, 256 threads, 4 x 4 outputs per thread
__global__ void __launch_bounds__(256) gjb_update_kernel(double* M, double* Inv, int64_t B, int k0, int pw, const double* __restrict__ ARK, const double* __restrict__ TK,
                                                         const int* __restrict__ active)
{
    __shared__ double As[kGjP][64 + 1];   // As[j][i] = ARK(r0 + i, j)
    __shared__ double Ts[kGjP][64 + 1];   // Ts[j][cc] = TK(j, c0 + cc), zero for columns that are not live
    __shared__ int s_any;
    const int64_t r0 = (int64_t) blockIdx.x * 64, c0 = (int64_t) blockIdx.y * 64;
    const int tid = threadIdx.x;
    if (tid == 0)
        s_any = 0;
    __syncthreads();
    if (tid < 64 && gjb_col_live(c0 + tid, B, k0, pw, active))
        s_any = 1;
    __syncthreads();
    if (!s_any)
        return;
    for (int e = tid; e < kGjP * 64; e += 256)
    {
        const int i = e % 64, j = e / 64;
        const int64_t r = r0 + i;
        As[j][i] = (j < pw && r < B && (r < k0 || r >= k0 + pw)) ? ARK[r + (int64_t) j * B] : 0.0;
    }
    for (int e = tid; e < kGjP * 64; e += 256)
    {
        const int j = e % kGjP, cc = e / kGjP;
        const int64_t c = c0 + cc;
        Ts[j][cc] = (j < pw && gjb_col_live(c, B, k0, pw, active)) ? TK[j + c * kGjP] : 0.0;
    }
    __syncthreads();
    const int tx = tid & 15, ty = tid >> 4;
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++)
            acc[a][b] = 0.0;
    for (int j = 0; j < kGjP; j++)
    {
        double av[4], tv[4];
#pragma unroll
        for (int a = 0; a < 4; a++)
            av[a] = As[j][tx + 16 * a];
#pragma unroll
        for (int b = 0; b < 4; b++)
            tv[b] = Ts[j][ty + 16 * b];
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int b = 0; b < 4; b++)
                acc[a][b] = fma(av[a], tv[b], acc[a][b]);
    }
#pragma unroll
    for (int b = 0; b < 4; b++)
    {
        const int64_t c = c0 + ty + 16 * b;
        if (!gjb_col_live(c, B, k0, pw, active))
            continue;
        double* X = c < B ? M + c * B : Inv + (c - B) * B;
#pragma unroll
        for (int a = 0; a < 4; a++)
        {
            const int64_t r = r0 + tx + 16 * a;
            if (r < B && (r < k0 || r >= k0 + pw) && acc[a][b] != 0.0)
                X[r] -= acc[a][b];
        }
    }
}

// panel columns of the rows outside the pivot rows are eliminated exactly
__global__ void gjb_clear_panel_kernel(double* M, int64_t B, int k0, int pw)
{
    for (int64_t e = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; e < B * pw; e += (int64_t) gridDim.x * blockDim.x)
    {
        const int64_t r = e % B;
        if (r < k0 || r >= k0 + pw)
            M[r + (int64_t) (k0 + e / B) * B] = 0.0;
    }
}

// ---- wide bands: block-tridiagonal elimination with grid-wide block kernels (see the header) ----
// S (B x B, column-major) += entries of block row `bi` of the CSR whose column lies in block `bi`
__global__ void thomas_diag_scatter_kernel(const int* __restrict__ rowptr, const int* __restrict__ col, const double* __restrict__ val, int64_t n, int B, int64_t bi,
                                           double* S)
{
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < B; r += gridDim.x * blockDim.x)
    {
        const int64_t row = bi * B + r;
        if (row >= n)
            continue;
        for (int p = rowptr[row]; p < rowptr[row + 1]; p++)
        {
            const int64_t c = (int64_t) col[p] - bi * B;
            if (c >= 0 && c < B)
                atomicAdd(S + r + c * B, val[p]);  // column blocks of the operator may split a row
        }
    }
}

// diagonal of block row bi: -= sigma on real rows, identity on the padding rows of the last block; Inv := I
__global__ void thomas_diag_shift_kernel(double* S, double* Inv, int64_t n, int B, int64_t bi, double sigma, int* orig, int* active)
{
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < B; r += gridDim.x * blockDim.x)
    {
        orig[r] = r;
        active[r] = 0;
        const int64_t row = bi * B + r;
        if (row < n)
            S[r + (int64_t) r * B] -= sigma;
        else
            S[r + (int64_t) r * B] = 1.0;
        Inv[r + (int64_t) r * B] = 1.0;
    }
}

// Out += sign * A(rb, cb) * Dm          (sparse block from the CSR times a dense B x B block)
__global__ void thomas_sparse_dense_kernel(const int* __restrict__ rowptr, const int* __restrict__ col, const double* __restrict__ val, int64_t n, int B, int64_t rb,
                                           int64_t cb, const double* __restrict__ Dm, double* Out, double sign)
{
    const int64_t total = (int64_t) B * B;
    for (int64_t e = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t) gridDim.x * blockDim.x)
    {
        const int r = (int) (e % B);
        const int64_t c = e / B;
        const int64_t row = rb * B + r;
        if (row >= n)
            continue;
        double acc = 0.0;
        for (int p = rowptr[row]; p < rowptr[row + 1]; p++)
        {
            const int64_t k = (int64_t) col[p] - cb * B;
            if (k >= 0 && k < B)
                acc = fma(val[p], Dm[k + c * B], acc);
        }
        if (acc != 0.0)
            Out[e] += sign * acc;
    }
}

// Out += sign * Dm * A(rb, cb)^T        (dense B x B block times the transpose of a sparse block; A symmetric: A(rb, cb)^T = A(cb, rb))
__global__ void thomas_dense_sparse_t_kernel(const int* __restrict__ rowptr, const int* __restrict__ col, const double* __restrict__ val, int64_t n, int B,
                                             int64_t rb, int64_t cb, const double* __restrict__ Dm, double* Out, double sign)
{
    const int64_t total = (int64_t) B * B;
    for (int64_t e = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t) gridDim.x * blockDim.x)
    {
        const int r = (int) (e % B);
        const int64_t c = e / B;
        const int64_t row = rb * B + c;  // row c of the sparse block = column c of its transpose
        if (row >= n)
            continue;
        double acc = 0.0;
        for (int p = rowptr[row]; p < rowptr[row + 1]; p++)
        {
            const int64_t k = (int64_t) col[p] - cb * B;
            if (k >= 0 && k < B)
                acc = fma(val[p], Dm[r + k * B], acc);
        }
        if (acc != 0.0)
            Out[e] += sign * acc;
    }
}

// y = beta * y + alpha * M x   (M column-major B x B): one warp per 32 rows, lanes along rows, 8 warps split the columns
__global__ void __launch_bounds__(256) block_gemv_kernel(const double* __restrict__ M, const double* __restrict__ x, double* __restrict__ y, int64_t B, double alpha,
                                                         double beta)
{
    __shared__ double s_part[8][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t r = (int64_t) blockIdx.x * 32 + lane;
    double acc = 0.0;
    if (r < B)
        for (int64_t c = warp; c < B; c += 8)
            acc = fma(M[r + c * B], x[c], acc);
    s_part[warp][lane] = acc;
    __syncthreads();
    if (warp == 0 && r < B)
    {
        double t = 0.0;
        for (int w = 0; w < 8; w++)
            t += s_part[w][lane];
        y[r] = (beta == 0.0 ? 0.0 : beta * y[r]) + alpha * t;
    }
}

// The same product with the columns split over gridDim.y CTAs per 32-row group (enough loads in flight to stream a large block at HBM
// speed): every CTA writes the partial sums of its column range to `part`, the CTA that takes the last ticket of its row group adds them
// in split order (fixed => bit-reproducible) and applies alpha / beta.  tickets[] must be zero on entry and is left zero.
__global__ void __launch_bounds__(256) block_gemv_split_kernel(const double* __restrict__ M, const double* __restrict__ x, double* __restrict__ y, int64_t B, double alpha,
                                                               double beta, double* part, unsigned int* tickets)
{
    __shared__ double s_part[8][32];
    __shared__ unsigned int s_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nsplit = gridDim.y, split = blockIdx.y;
    const int64_t r = (int64_t) blockIdx.x * 32 + lane;
    const int64_t per = (B + nsplit - 1) / nsplit;
    const int64_t c0 = (int64_t) split * per, c1 = c0 + per < B ? c0 + per : B;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    if (r < B)
    {
        int64_t c = c0 + warp;
        for (; c + 24 < c1; c += 32)
        {
            const double m0 = M[r + c * B], m1 = M[r + (c + 8) * B], m2 = M[r + (c + 16) * B], m3 = M[r + (c + 24) * B];
            a0 = fma(m0, x[c], a0);
            a1 = fma(m1, x[c + 8], a1);
            a2 = fma(m2, x[c + 16], a2);
            a3 = fma(m3, x[c + 24], a3);
        }
        for (; c < c1; c += 8)
            a0 = fma(M[r + c * B], x[c], a0);
    }
    s_part[warp][lane] = (a0 + a1) + (a2 + a3);
    __syncthreads();
    if (warp == 0 && r < B)
    {
        double t = 0.0;
        for (int w = 0; w < 8; w++)
            t += s_part[w][lane];
        part[(int64_t) split * B + r] = t;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0)
        s_last = atomicAdd(&tickets[blockIdx.x], 1u) == (unsigned int) (nsplit - 1) ? 1u : 0u;
    __syncthreads();
    if (s_last == 0u)
        return;
    __threadfence();
    if (warp == 0 && r < B)
    {
        double t = 0.0;
        for (int q = 0; q < nsplit; q++)
            t += part[(int64_t) q * B + r];
        y[r] = (beta == 0.0 ? 0.0 : beta * y[r]) + alpha * t;
    }
    if (threadIdx.x == 0)
        tickets[blockIdx.x] = 0u;
}

int grid_for(int64_t n, int block = 256)
{
    const int sms = device_info().sm_count;
    return (int) std::max<int64_t>(1, std::min<int64_t>((n + block - 1) / block, (int64_t) sms * 8));
}

}  // namespace

struct BandSolve
{
    int64_t n = 0, N = 0;
    int B = 0, levels = 0, first_top = 0, bw = 0;
    bool factored = false;
    double sigma = 0.0;
    int refine = 1;        // refinement steps the solve performs (0 / 1)
    int refine_mode = -1;  // -1: decided by set_shift (the unrefined verification residual), 0 / 1: fixed by band_set_refine()
    double verify_rel = 0.0, verify_rel_unrefined = -1.0;
    DevBuf<double> D, Lo, Up, Dinv, GL, GU, ML, MU;
    DevBuf<double> xb, rb, t;
    DevBuf<int> flag;
    BcrView view;
    int64_t solves = 0, launches = 0;
    // dense fallback (small non-banded matrices)
    bool dense = false;
    DevBuf<double> Mden, Iden, colk, scal;
    DevBuf<int> ipiv;
    // wide bands: block-tridiagonal elimination (B = half-bandwidth > 32); factors ML / Sinv / GU, N blocks of B x B each
    bool thomas = false;
    DevBuf<double> tML, tSinv, tGU, tS, tX, tpart;
    DevBuf<unsigned int> ttick;
    DevBuf<int> torig;  // [0, B): original row index per position, [B, 2B): active flags of the Inv columns (gj3_* / gjb_* kernels)
    DevBuf<double> twork, tark, ttk, tkk;  // blocked Gauss-Jordan: panel copy, A_RK, T(K, :), A_KK^{-1}
    DevBuf<int> tpiv;
    int tsplit = 1;
#ifndef SB200_EMU
    // the solve is a fixed sequence of 3 N launches on fixed buffers: captured once per factorisation, replayed per solve (SB200_SHIFT_GRAPH=0 turns it off)
    cudaGraphExec_t tgraph[2] = {nullptr, nullptr};
    ~BandSolve()
    {
        for (cudaGraphExec_t g : tgraph)
            if (g)
                cudaGraphExecDestroy(g);
    }
#endif
};

void band_destroy(BandSolve* b) { delete b; }

// half-bandwidth of the operator's CSR and the block layout; no factorisation yet
BandSolve* band_create(sb200_op* op)
{
    const DeviceCsr& A = op->A;
    SB200_REQUIRE(A.row0 == 0 && A.nrows == A.n, SB200_INVALID_ARGUMENT, "the shift-solve operator is single-GPU (whole matrix on one device)");
    std::unique_ptr<BandSolve> b(new BandSolve());
    b->n = A.n;
    b->flag.alloc(2);
    SB200_CUDA_CHECK(cudaMemsetAsync(b->flag.get(), 0, sizeof(int) * 2, op->stream));
    auto scan = [&](const int* rp, const int* ci) { band_width_kernel<<<grid_for(A.n), 256, 0, op->stream>>>(rp, ci, A.nrows, A.row0, b->flag.get() + 1); };
    if (A.blocks.empty())
        scan(A.rowptr.get(), A.col.get());
    else
        for (const CsrBlock& blk : A.blocks)
            scan(blk.rowptr.get(), blk.col.get());
    SB200_CUDA_CHECK(cudaGetLastError());
    int h[2];
    SB200_CUDA_CHECK(cudaMemcpyAsync(h, b->flag.get(), sizeof(h), cudaMemcpyDeviceToHost, op->stream));
    SB200_CUDA_CHECK(cudaStreamSynchronize(op->stream));
    b->bw = h[1];
    // route: 0 = by size (default), 1 = block cyclic reduction, 2 = dense inverse, 3 = wide-band block elimination (tests force 2 / 3 on small inputs)
    const int forced_route = [] {  // read at every construction: the tests switch it per case
        const char* e = std::getenv("SB200_SHIFT_ROUTE");
        return !e ? 0 : (std::string(e) == "bcr" ? 1 : (std::string(e) == "dense" ? 2 : (std::string(e) == "thomas" ? 3 : 0)));
    }();
    const bool want_thomas = forced_route == 3 ? b->bw >= 1 : (forced_route == 0 && b->bw > 32 && A.n > kDenseMax);
    if (want_thomas)
    {
        b->thomas = true;
        b->levels = -1;  // marks the sequential block elimination in band_info()
        size_t free_b = 0, total_b = 0;
        SB200_CUDA_CHECK(cudaMemGetInfo(&free_b, &total_b));
        auto bytes_for = [&](int64_t Bq) {
            const size_t Nq = (size_t) ((A.n + Bq - 1) / Bq), bbq = (size_t) Bq * Bq;
            return sizeof(double) * (3 * Nq * bbq + 2 * bbq + 5 * Nq * (size_t) Bq + 17 * (size_t) Bq);
        };
        // Block size: any B >= half-bandwidth keeps the block-tridiagonal form.  A solve is 3 N sequential block products, so narrow bands with
        // many block rows would be launch bound; merging m = floor(Bmin / b) band-widths per block (B ~ 1000) trades a larger factor stream
        // (3 n B doubles) and inverse (n B^2 flops) for m times fewer steps.  SB200_SHIFT_BMIN overrides the target (0: B = b).
        const int64_t b0 = std::max(b->bw, 4);
        int64_t bmin = 1024;
        if (const char* e = std::getenv("SB200_SHIFT_BMIN"))
            bmin = std::max(0, std::atoi(e));
        int64_t m = std::max<int64_t>(1, bmin / b0);
        while (m > 1 && (b0 * m > A.n || bytes_for(b0 * m) > free_b / 2))
            m--;
        b->B = (int) (b0 * m);
        b->N = (A.n + b->B - 1) / b->B;
        const size_t bb = (size_t) b->B * b->B;
        const size_t need = bytes_for(b->B);
        SB200_REQUIRE(need <= free_b - free_b / 8, SB200_INVALID_ARGUMENT,
                      "SparseSymShiftSolve: half-bandwidth " + std::to_string(b->bw) + " at n = " + std::to_string(A.n) + " needs " + std::to_string(need >> 20) +
                          " MiB of block factors, more than the free device memory; the device shift-solve handles banded / mesh-like patterns only");
        b->tML.alloc((size_t) b->N * bb);
        b->tSinv.alloc((size_t) b->N * bb);
        b->tGU.alloc((size_t) b->N * bb);
        b->tS.alloc(bb);
        {
            // column splits of the solve's block products: about four CTAs per SM in total
            const int64_t groups = (b->B + 31) / 32;
            const int sms = device_info().sm_count;
            b->tsplit = (int) std::max<int64_t>(1, std::min<int64_t>(16, (4 * (int64_t) sms + groups - 1) / groups));
            if (b->B < 256)
                b->tsplit = 1;
            if (const char* e = std::getenv("SB200_SHIFT_SPLIT"))  // tests: force the split kernel on small blocks
                b->tsplit = std::max(1, std::min(16, std::atoi(e)));
            b->tpart.alloc((size_t) b->tsplit * b->B);
            b->ttick.alloc((size_t) groups);
            SB200_CUDA_CHECK(cudaMemsetAsync(b->ttick.get(), 0, sizeof(unsigned int) * (size_t) groups, op->stream));
        }
        b->tX.alloc((size_t) b->N * b->B);
        b->torig.alloc(2 * (size_t) b->B);
        b->colk.alloc((size_t) b->B);
        b->scal.alloc(2);
        b->ipiv.alloc(1);
        const size_t npad = (size_t) b->N * b->B;
        b->xb.alloc(npad);
        b->rb.alloc(npad);
        b->t.alloc(npad);
        return b.release();
    }
    if (b->bw > 32 || forced_route == 2)
    {
        SB200_REQUIRE(A.n <= kDenseMax, SB200_INVALID_ARGUMENT,
                      "SparseSymShiftSolve: half-bandwidth " + std::to_string(b->bw) + " exceeds 32 and n exceeds " + std::to_string(kDenseMax) +
                          "; the device shift-solve handles banded matrices (and small general ones) only");
        b->dense = true;
        b->B = 0;
        b->N = 0;
        const size_t nn = (size_t) A.n * A.n;
        b->Mden.alloc(nn);
        b->Iden.alloc(nn);
        b->colk.alloc((size_t) A.n);
        b->scal.alloc(2);
        b->ipiv.alloc(1);
        b->xb.alloc((size_t) A.n);
        b->rb.alloc((size_t) A.n);
        b->t.alloc((size_t) A.n);
        return b.release();
    }
    b->B = std::max(b->bw, 4);
    b->N = (A.n + b->B - 1) / b->B;
    b->levels = 0;
    while ((b->N >> (b->levels + 1)) >= 1)
        b->levels++;
    SB200_REQUIRE(b->levels < 40, SB200_LOGIC, "too many BCR levels");
    b->first_top = 0;
    while (b->first_top < b->levels && (b->N >> b->first_top) > kTopRows)
        b->first_top++;
    const size_t bb = (size_t) b->B * b->B;
    const size_t tot = (size_t) b->N * bb;
    b->D.alloc(tot);
    b->Lo.alloc(tot);
    b->Up.alloc(tot);
    b->Dinv.alloc(tot);
    b->GL.alloc(tot);
    b->GU.alloc(tot);
    int64_t kept_total = 0;
    for (int l = 0; l < b->levels; l++)
    {
        b->view.ml_off[l] = kept_total;
        kept_total += (b->N >> l) / 2;
    }
    b->ML.alloc((size_t) std::max<int64_t>(kept_total, 1) * bb);
    b->MU.alloc((size_t) std::max<int64_t>(kept_total, 1) * bb);
    const size_t npad = (size_t) b->N * b->B;
    b->xb.alloc(npad);
    b->rb.alloc(npad);
    b->t.alloc(npad);
    b->view.Dinv = b->Dinv.get();
    b->view.GL = b->GL.get();
    b->view.GU = b->GU.get();
    b->view.ML = b->ML.get();
    b->view.MU = b->MU.get();
    b->view.N = b->N;
    b->view.B = b->B;
    b->view.levels = b->levels;
    return b.release();
}

// one BCR solve in place on the padded vector f (N*B entries)
static void bcr_solve_inplace(BandSolve* b, double* f, cudaStream_t st)
{
    for (int l = 0; l < b->first_top; l++)
    {
        const int64_t nkept = (b->N >> l) / 2;
        if (b->B <= 16)
            bcr_forward_kernel<16><<<(unsigned) ((nkept + 3) / 4), 128, 0, st>>>(b->view, f, l, nkept);
        else
            bcr_forward_kernel<32><<<(unsigned) ((nkept + 3) / 4), 128, 0, st>>>(b->view, f, l, nkept);
        b->launches++;
    }
    if (b->B <= 16)
        bcr_top_kernel<16><<<1, kTopThreads, 0, st>>>(b->view, f, b->first_top);
    else
        bcr_top_kernel<32><<<1, kTopThreads, 0, st>>>(b->view, f, b->first_top);
    b->launches++;
    for (int l = b->first_top - 1; l >= 0; l--)
    {
        const int64_t nelim = ((b->N >> l) + 1) / 2;
        if (b->B <= 16)
            bcr_backward_kernel<16><<<(unsigned) ((nelim + 3) / 4), 128, 0, st>>>(b->view, f, l, nelim);
        else
            bcr_backward_kernel<32><<<(unsigned) ((nelim + 3) / 4), 128, 0, st>>>(b->view, f, l, nelim);
        b->launches++;
    }
    SB200_CUDA_CHECK(cudaGetLastError());
}

// y = beta y + alpha M x for one B x B block of the factors
static void thomas_gemv(BandSolve* b, const double* M, const double* x, double* y, double alpha, double beta, cudaStream_t st)
{
    const int64_t B = b->B;
    const unsigned g = (unsigned) ((B + 31) / 32);
    if (b->tsplit > 1)
        block_gemv_split_kernel<<<dim3(g, (unsigned) b->tsplit), 256, 0, st>>>(M, x, y, B, alpha, beta, b->tpart.get(), b->ttick.get());
    else
        block_gemv_kernel<<<g, 256, 0, st>>>(M, x, y, B, alpha, beta);
}

// one block-elimination solve in place on the padded vector f (N*B entries): 3 N - 2 block products
static void thomas_enqueue(BandSolve* b, double* f, cudaStream_t st)
{
    const int64_t B = b->B, N = b->N;
    const size_t bb = (size_t) B * B;
    double* X = b->tX.get();
    for (int64_t i = 1; i < N; i++)  // g_i = f_i - ML_i g_{i-1}
        thomas_gemv(b, b->tML.get() + (size_t) i * bb, f + (i - 1) * B, f + i * B, -1.0, 1.0, st);
    for (int64_t i = N - 1; i >= 0; i--)  // x_i = Sinv_i g_i - GU_i x_{i+1}
    {
        thomas_gemv(b, b->tSinv.get() + (size_t) i * bb, f + i * B, X + i * B, 1.0, 0.0, st);
        if (i + 1 < N)
            thomas_gemv(b, b->tGU.get() + (size_t) i * bb, X + (i + 1) * B, X + i * B, -1.0, 1.0, st);
    }
    SB200_CUDA_CHECK(cudaMemcpyAsync(f, X, sizeof(double) * (size_t) (N * B), cudaMemcpyDeviceToDevice, st));
}

static void thomas_solve_inplace(BandSolve* b, double* f, cudaStream_t st)
{
#ifndef SB200_EMU
    static const bool use_graph = [] { const char* e = std::getenv("SB200_SHIFT_GRAPH"); return !(e && e[0] == '0'); }();
    const int which = f == b->xb.get() ? 0 : (f == b->rb.get() ? 1 : -1);
    if (use_graph && which >= 0)
    {
        if (!b->tgraph[which])
        {
            cudaGraph_t graph = nullptr;
            SB200_CUDA_CHECK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            try
            {
                thomas_enqueue(b, f, st);
            }
            catch (...)
            {
                cudaStreamEndCapture(st, &graph);
                if (graph)
                    cudaGraphDestroy(graph);
                throw;
            }
            SB200_CUDA_CHECK(cudaStreamEndCapture(st, &graph));
            const cudaError_t e = cudaGraphInstantiate(&b->tgraph[which], graph, 0);
            cudaGraphDestroy(graph);
            SB200_CUDA_CHECK(e);
        }
        SB200_CUDA_CHECK(cudaGraphLaunch(b->tgraph[which], st));
        b->launches += 1;
        return;
    }
#endif
    thomas_enqueue(b, f, st);
    b->launches += 3 * b->N - 1;
    SB200_CUDA_CHECK(cudaGetLastError());
}

// y = (A - sigma I)^{-1} x, device pointers (n entries each; x == y allowed)
void band_solve_device(sb200_op* op, const double* x, double* y)
{
    BandSolve* b = op->band;
    SB200_REQUIRE(b && b->factored, SB200_LOGIC, "SparseSymShiftSolve: set_shift() has not been called");
    cudaStream_t st = op->stream;
    if (b->thomas)
    {
        const int64_t n = b->n, npad = b->N * b->B;
        pad_copy_kernel<<<grid_for(npad), 256, 0, st>>>(x, b->xb.get(), n, npad);
        thomas_solve_inplace(b, b->xb.get(), st);
        b->launches++;
        if (b->refine > 0)
        {
            launch_spmv(op->A, op->plan, b->xb.get(), b->t.get(), st);
            refine_residual_kernel<<<grid_for(npad), 256, 0, st>>>(x, b->t.get(), b->xb.get(), b->sigma, b->rb.get(), n, npad);
            thomas_solve_inplace(b, b->rb.get(), st);
            add_out_kernel<<<grid_for(n), 256, 0, st>>>(b->xb.get(), b->rb.get(), y, n);
            b->launches += 3;
        }
        else
        {
            add_out_kernel<<<grid_for(n), 256, 0, st>>>(b->xb.get(), nullptr, y, n);
            b->launches++;
        }
        SB200_CUDA_CHECK(cudaGetLastError());
        b->solves++;
        return;
    }
    if (b->dense)
    {
        const int64_t n = b->n;
        const unsigned g = (unsigned) ((n + 31) / 32);
        dense_gemv_kernel<<<g, 256, 0, st>>>(b->Iden.get(), x, b->xb.get(), n);
        if (b->refine > 0)
        {
            launch_spmv(op->A, op->plan, b->xb.get(), b->t.get(), st);
            refine_residual_kernel<<<grid_for(n), 256, 0, st>>>(x, b->t.get(), b->xb.get(), b->sigma, b->rb.get(), n, n);
            dense_gemv_kernel<<<g, 256, 0, st>>>(b->Iden.get(), b->rb.get(), b->t.get(), n);
            add_out_kernel<<<grid_for(n), 256, 0, st>>>(b->xb.get(), b->t.get(), y, n);
            b->launches += 5;
        }
        else
        {
            add_out_kernel<<<grid_for(n), 256, 0, st>>>(b->xb.get(), nullptr, y, n);
            b->launches += 2;
        }
        SB200_CUDA_CHECK(cudaGetLastError());
        b->solves++;
        return;
    }
    const int64_t n = b->n, npad = b->N * b->B;
    pad_copy_kernel<<<grid_for(npad), 256, 0, st>>>(x, b->xb.get(), n, npad);
    bcr_solve_inplace(b, b->xb.get(), st);
    b->launches++;
    if (b->refine > 0)
    {
        // r = x - (A - sigma I) y0 ; y = y0 + solve(r)
        launch_spmv(op->A, op->plan, b->xb.get(), b->t.get(), st);
        refine_residual_kernel<<<grid_for(npad), 256, 0, st>>>(x, b->t.get(), b->xb.get(), b->sigma, b->rb.get(), n, npad);
        bcr_solve_inplace(b, b->rb.get(), st);
        add_out_kernel<<<grid_for(n), 256, 0, st>>>(b->xb.get(), b->rb.get(), y, n);
        b->launches += 3;
    }
    else
    {
        add_out_kernel<<<grid_for(n), 256, 0, st>>>(b->xb.get(), nullptr, y, n);
        b->launches++;
    }
    SB200_CUDA_CHECK(cudaGetLastError());
    b->solves++;
}

static void factor_dense(sb200_op* op, BandSolve* b, double sigma)
{
    const DeviceCsr& A = op->A;
    cudaStream_t st = op->stream;
    const int64_t n = A.n;
    b->Mden.zero(st);
    b->Iden.zero(st);
    auto scat = [&](const int* rp, const int* ci, const double* v) { dense_scatter_kernel<<<grid_for(n), 256, 0, st>>>(rp, ci, v, n, b->Mden.get()); };
    if (A.blocks.empty())
        scat(A.rowptr.get(), A.col.get(), A.val.get());
    else
        for (const CsrBlock& blk : A.blocks)
            scat(blk.rowptr.get(), blk.col.get(), blk.val.get());
    dense_init_kernel<<<grid_for(n), 256, 0, st>>>(b->Mden.get(), b->Iden.get(), n, sigma);
    const int ge = grid_for(2 * n * n);
    for (int k = 0; k < (int) n; k++)
    {
        gj_pivot_kernel<<<1, 1024, 0, st>>>(b->Mden.get(), n, k, b->ipiv.get(), b->scal.get(), b->flag.get());
        gj_save_kk_kernel<<<1, 1, 0, st>>>(b->Mden.get(), n, k, b->scal.get() + 1);
        gj_swap_scale_kernel<<<grid_for(2 * n), 256, 0, st>>>(b->Mden.get(), b->Iden.get(), n, k, b->ipiv.get(), b->scal.get(), b->colk.get());
        gj_fix_colk_kernel<<<1, 1, 0, st>>>(b->colk.get(), k, b->ipiv.get(), b->scal.get() + 1);
        gj_eliminate_kernel<<<ge, 256, 0, st>>>(b->Mden.get(), b->Iden.get(), n, k, b->colk.get());
    }
    SB200_CUDA_CHECK(cudaGetLastError());
}

// wide bands: S_i = D_i - ML_i U_{i-1}, ML_i = L_i Sinv_{i-1}, Sinv_i = S_i^{-1}, GU_i = Sinv_i U_i; L_i = A(i, i-1) and U_i = A(i+1, i)^T from the CSR
static void factor_thomas(sb200_op* op, BandSolve* b, double sigma)
{
    const DeviceCsr& A = op->A;
    cudaStream_t st = op->stream;
    const int B = b->B;
    const int64_t N = b->N, n = A.n;
    const size_t bb = (size_t) B * B;
    const int gb = grid_for((int64_t) bb), gr = grid_for(B), ge = grid_for(2 * (int64_t) bb);
    // every CSR kernel runs once per column block of the operator (a column block holds part of every row)
    auto for_blocks = [&](auto&& fn) {
        if (A.blocks.empty())
            fn(A.rowptr.get(), A.col.get(), A.val.get());
        else
            for (const CsrBlock& blk : A.blocks)
                fn(blk.rowptr.get(), blk.col.get(), blk.val.get());
    };
    b->tML.zero(st);
    b->tSinv.zero(st);
    b->tGU.zero(st);
    const bool blocked = [&] {
        const char* e = std::getenv("SB200_SHIFT_GJ");
        if (e && std::string(e) == "rank1")
            return false;
        if (e && std::string(e) == "blocked")
            return true;
        return B >= 128;
    }();
    if (blocked && b->twork.n == 0)
    {
        b->twork.alloc((size_t) B * kGjP);
        b->tark.alloc((size_t) B * kGjP);
        b->ttk.alloc((size_t) kGjP * 2 * (size_t) B);
        b->tkk.alloc((size_t) kGjP * kGjP);
        b->tpiv.alloc(kGjP);
    }
    for (int64_t i = 0; i < N; i++)
    {
        double* S = b->tS.get();
        double* Sinv = b->tSinv.get() + (size_t) i * bb;
        double* ML = b->tML.get() + (size_t) i * bb;
        SB200_CUDA_CHECK(cudaMemsetAsync(S, 0, sizeof(double) * bb, st));
        for_blocks([&](const int* rp, const int* ci, const double* v) { thomas_diag_scatter_kernel<<<gr, 256, 0, st>>>(rp, ci, v, n, B, i, S); });
        thomas_diag_shift_kernel<<<gr, 256, 0, st>>>(S, Sinv, n, B, i, sigma, b->torig.get(), b->torig.get() + B);
        if (i > 0)
        {
            const double* Sprev = b->tSinv.get() + (size_t) (i - 1) * bb;
            for_blocks([&](const int* rp, const int* ci, const double* v) { thomas_sparse_dense_kernel<<<gb, 256, 0, st>>>(rp, ci, v, n, B, i, i - 1, Sprev, ML, 1.0); });
            for_blocks([&](const int* rp, const int* ci, const double* v) { thomas_dense_sparse_t_kernel<<<gb, 256, 0, st>>>(rp, ci, v, n, B, i, i - 1, ML, S, -1.0); });
        }
        // Sinv_i = S^{-1} by Gauss-Jordan on the (S | I) tableau: blocked (panels of kGjP pivots) for large blocks, rank-1 updates otherwise
        if (blocked)
        {
            const dim3 gt((unsigned) ((B + 63) / 64), (unsigned) ((2 * (int64_t) B + 63) / 64));
            for (int k0 = 0; k0 < B; k0 += kGjP)
            {
                const int pw = std::min(kGjP, B - k0);
                gjb_panel_kernel<<<1, 1024, 0, st>>>(S, B, k0, pw, b->twork.get(), b->tpiv.get(), b->flag.get());
                gjb_swap_kernel<<<grid_for(2 * (int64_t) B), 256, 0, st>>>(S, Sinv, B, k0, pw, b->tpiv.get(), b->torig.get(), b->torig.get() + B);
                gjb_kk_kernel<<<1, 32, 0, st>>>(S, B, k0, pw, b->tkk.get(), b->flag.get());
                gjb_rows_kernel<<<grid_for(2 * (int64_t) B), 256, 0, st>>>(S, Sinv, B, k0, pw, b->tkk.get(), b->tark.get(), b->ttk.get(), b->torig.get() + B);
                gjb_update_kernel<<<gt, 256, 0, st>>>(S, Sinv, B, k0, pw, b->tark.get(), b->ttk.get(), b->torig.get() + B);
                gjb_clear_panel_kernel<<<grid_for((int64_t) B * pw), 256, 0, st>>>(S, B, k0, pw);
            }
        }
        else
        for (int k = 0; k < B; k++)
        {
            gj3_pivot_kernel<<<1, 1024, 0, st>>>(S, B, k, b->ipiv.get(), b->scal.get(), b->flag.get(), b->torig.get(), b->torig.get() + B);
            gj3_swap_scale_kernel<<<grid_for(2 * (int64_t) B), 256, 0, st>>>(S, Sinv, B, k, b->ipiv.get(), b->scal.get(), b->colk.get());
            gj3_eliminate_kernel<<<ge, 256, 0, st>>>(S, Sinv, B, k, b->colk.get(), b->torig.get() + B);
        }
        if (i + 1 < N)
        {
            double* GU = b->tGU.get() + (size_t) i * bb;
            for_blocks([&](const int* rp, const int* ci, const double* v) { thomas_dense_sparse_t_kernel<<<gb, 256, 0, st>>>(rp, ci, v, n, B, i + 1, i, Sinv, GU, 1.0); });
        }
        SB200_CUDA_CHECK(cudaGetLastError());
    }
}

static void factor_band(sb200_op* op, BandSolve* b, double sigma)
{
    const DeviceCsr& A = op->A;
    cudaStream_t st = op->stream;
    const int B = b->B;
    const size_t bb = (size_t) B * B;
    b->D.zero(st);
    b->Lo.zero(st);
    b->Up.zero(st);
    auto scatter = [&](const int* rp, const int* ci, const double* v) {
        band_scatter_kernel<<<grid_for(A.n), 256, 0, st>>>(rp, ci, v, A.n, B, b->D.get(), b->Lo.get(), b->Up.get(), b->flag.get());
    };
    if (A.blocks.empty())
        scatter(A.rowptr.get(), A.col.get(), A.val.get());
    else
        for (const CsrBlock& blk : A.blocks)
            scatter(blk.rowptr.get(), blk.col.get(), blk.val.get());
    band_diag_kernel<<<grid_for(b->N * B), 256, 0, st>>>(b->D.get(), A.n, b->N, B, sigma);
    SB200_CUDA_CHECK(cudaGetLastError());
    const size_t smem = 3 * bb * sizeof(double);
    SB200_CUDA_CHECK(cudaFuncSetAttribute((const void*) bcr_eliminate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    SB200_CUDA_CHECK(cudaFuncSetAttribute((const void*) bcr_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    for (int l = 0; l < b->levels; l++)
    {
        const int64_t cnt = b->N >> l, nelim = (cnt + 1) / 2, nkept = cnt / 2;
        bcr_eliminate_kernel<<<(unsigned) nelim, 32, smem, st>>>(b->D.get(), b->Lo.get(), b->Up.get(), b->Dinv.get(), b->GL.get(), b->GU.get(), b->N, B, l, cnt,
                                                                  b->flag.get());
        if (nkept > 0)
            bcr_update_kernel<<<(unsigned) nkept, 32, smem, st>>>(b->D.get(), b->Lo.get(), b->Up.get(), b->Dinv.get(), b->GL.get(), b->GU.get(),
                                                                   b->ML.get() + (size_t) b->view.ml_off[l] * bb, b->MU.get() + (size_t) b->view.ml_off[l] * bb, b->N,
                                                                   B, l, cnt);
    }
    // top row: the eliminate kernel at level `levels` with a single active row (k = 1)
    bcr_eliminate_kernel<<<1, 32, smem, st>>>(b->D.get(), b->Lo.get(), b->Up.get(), b->Dinv.get(), b->GL.get(), b->GU.get(), b->N, B, b->levels, 1, b->flag.get());
    SB200_CUDA_CHECK(cudaGetLastError());
}

// set_shift (SparseSymShiftSolve.h:85-95)
void band_set_shift(sb200_op* op, double sigma)
{
    BandSolve* b = op->band;
    SB200_REQUIRE(b != nullptr, SB200_INVALID_ARGUMENT, "operator is not a shift-solve operator");
    const DeviceCsr& A = op->A;
    cudaStream_t st = op->stream;
    const int B = b->B;
    b->factored = false;
    b->sigma = sigma;
    SB200_CUDA_CHECK(cudaMemsetAsync(b->flag.get(), 0, sizeof(int), st));
    if (b->thomas)
        factor_thomas(op, b, sigma);
    else if (b->dense)
        factor_dense(op, b, sigma);
    else
        factor_band(op, b, sigma);
    int h = 0;
    SB200_CUDA_CHECK(cudaMemcpyAsync(&h, b->flag.get(), sizeof(int), cudaMemcpyDeviceToHost, st));
    SB200_CUDA_CHECK(cudaStreamSynchronize(st));
    SB200_REQUIRE(h != 2, SB200_LOGIC, "band scatter met an entry outside the block tridiagonal");
    SB200_REQUIRE(h == 0, SB200_INVALID_ARGUMENT, "SparseSymShiftSolve: factorization failed with the given shift");
    b->factored = true;
    // verification solve: relative residual of one deterministic right-hand side.  It also decides whether the solves need their
    // refinement step: when the plain solve already reaches working precision (pivoted large blocks, diagonally dominant bands) the second
    // sweep is dropped; block cyclic reduction on an indefinite shifted band (residual ~1e-11) keeps it.
    {
        std::vector<double> hx((size_t) A.n), hy((size_t) A.n), ht((size_t) A.n);
        uint64_t sstate = 0x9E3779B97F4A7C15ULL;
        for (int64_t i = 0; i < A.n; i++)
        {
            sstate = sstate * 6364136223846793005ULL + 1442695040888963407ULL;
            hx[(size_t) i] = (double) (sstate >> 11) * (1.0 / 9007199254740992.0) - 0.5;
        }
        const size_t npad_v = std::max<size_t>((size_t) b->N * B, (size_t) A.n);
        DevBuf<double> dx((size_t) A.n), dy(npad_v), dt(npad_v);
        SB200_CUDA_CHECK(cudaMemcpyAsync(dx.get(), hx.data(), sizeof(double) * A.n, cudaMemcpyHostToDevice, st));
        auto residual = [&]() {
            dy.zero(st);
            band_solve_device(op, dx.get(), dy.get());
            launch_spmv(op->A, op->plan, dy.get(), dt.get(), st);
            SB200_CUDA_CHECK(cudaMemcpyAsync(hy.data(), dy.get(), sizeof(double) * A.n, cudaMemcpyDeviceToHost, st));
            SB200_CUDA_CHECK(cudaMemcpyAsync(ht.data(), dt.get(), sizeof(double) * A.n, cudaMemcpyDeviceToHost, st));
            SB200_CUDA_CHECK(cudaStreamSynchronize(st));
            double rn = 0.0, xn = 0.0;
            for (int64_t i = 0; i < A.n; i++)
            {
                const double r = hx[(size_t) i] - (ht[(size_t) i] - sigma * hy[(size_t) i]);
                rn += r * r;
                xn += hx[(size_t) i] * hx[(size_t) i];
            }
            return std::sqrt(rn / xn);
        };
        double rel;
        b->verify_rel_unrefined = -1.0;
        if (b->refine_mode < 0)
        {
            b->refine = 0;
            rel = b->verify_rel_unrefined = residual();
            if (!(rel <= 5e-14))
            {
                b->refine = 1;
                rel = residual();
            }
        }
        else
        {
            b->refine = b->refine_mode;
            rel = residual();
        }
        b->verify_rel = rel;
        if (!(rel <= 1e-8))
        {
            b->factored = false;
            throw Error(SB200_INVALID_ARGUMENT, "SparseSymShiftSolve: factorization failed with the given shift (verification residual " + std::to_string(rel) + ")");
        }
        b->solves = 0;
        b->launches = 0;
    }
}

void band_info(const sb200_op* op, int* half_bandwidth, int* block, int64_t* block_rows, int* levels)
{
    const BandSolve* b = op->band;
    SB200_REQUIRE(b != nullptr, SB200_INVALID_ARGUMENT, "operator is not a shift-solve operator");
    if (half_bandwidth)
        *half_bandwidth = b->bw;
    if (block)
        *block = b->B;
    if (block_rows)
        *block_rows = b->N;
    if (levels)
        *levels = b->levels;
}

void band_status(const sb200_op* op, int* refine_steps, double* verify_residual, double* unrefined_residual)
{
    const BandSolve* b = op->band;
    SB200_REQUIRE(b != nullptr, SB200_INVALID_ARGUMENT, "operator is not a shift-solve operator");
    SB200_REQUIRE(b->factored, SB200_LOGIC, "SparseSymShiftSolve: set_shift() has not been called");
    if (refine_steps)
        *refine_steps = b->refine;
    if (verify_residual)
        *verify_residual = b->verify_rel;
    if (unrefined_residual)
        *unrefined_residual = b->verify_rel_unrefined;
}

void band_set_refine(sb200_op* op, int steps)
{
    SB200_REQUIRE(op->band != nullptr, SB200_INVALID_ARGUMENT, "operator is not a shift-solve operator");
    // steps < 0: back to the automatic choice made by the next set_shift(); 0 / >= 1: fixed from now on (also for the current factorisation)
    op->band->refine_mode = steps < 0 ? -1 : (steps > 0 ? 1 : 0);
    if (steps >= 0)
        op->band->refine = op->band->refine_mode;
}

}  // namespace sb200

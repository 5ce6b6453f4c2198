// BLAS-1 helpers used outside the steady-state loop: Arnoldi::init (Arnoldi.h:136-195),
// expand_basis (Arnoldi.h:66-115) and the rare restart tests of Lanczos.h:99-121.
// Reductions are bit-reproducible (fixed combine order).
#include "kernels.h"

namespace sb200 {

namespace {

constexpr int kVecBlock = 256;

template <int OP>
__global__ void __launch_bounds__(kVecBlock) vec_reduce_kernel(const double* __restrict__ x, const double* __restrict__ y, int64_t n, double* out,
                                                               double* partials, unsigned int* ticket)
{
    double acc = 0.0;
    for (int64_t i = (int64_t) blockIdx.x * kVecBlock + threadIdx.x; i < n; i += (int64_t) gridDim.x * kVecBlock)
    {
        const double a = x[i];
        if (OP == VR_SUMSQ)
            acc = fma(a, a, acc);
        else if (OP == VR_DOT)
            acc = fma(a, y[i], acc);
        else if (OP == VR_CDOT_IM)
            acc = (i & 1) ? fma(-a, y[i - 1], acc) : fma(a, y[i + 1], acc);  // Im conj(x) y = x_re y_im - x_im y_re (n even)
        else if (OP == VR_CMAXABS)
            acc = (i & 1) ? acc : fmax(acc, hypot(a, x[i + 1]));
        else
            acc = fmax(acc, fabs(a));
    }
    __shared__ double s_w[kVecBlock / 32];
    if (OP == VR_MAXABS || OP == VR_CMAXABS)
    {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
            acc = fmax(acc, __shfl_xor_sync(0xffffffffu, acc, o));
    }
    else
        acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0)
        s_w[threadIdx.x >> 5] = acc;
    __syncthreads();
    double cta = 0.0;
    if (threadIdx.x == 0)
    {
        for (int q = 0; q < kVecBlock / 32; q++)
            cta = (OP == VR_MAXABS || OP == VR_CMAXABS) ? fmax(cta, s_w[q]) : cta + s_w[q];
    }
    if (OP == VR_MAXABS || OP == VR_CMAXABS)
    {
        // max is order independent: reuse the ticket scheme with a max combine
        __shared__ bool s_last;
        if (threadIdx.x == 0)
        {
            partials[(size_t) blockIdx.x * 128] = cta;
            __threadfence();
            const unsigned int t = atomicAdd(ticket, 1u);
            s_last = (t == gridDim.x - 1);
        }
        __syncthreads();
        if (s_last)
        {
            __threadfence();
            double mx = 0.0;
            for (int b = threadIdx.x; b < (int) gridDim.x; b += kVecBlock)
                mx = fmax(mx, ld_cg_f64(partials + (size_t) b * 128));
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
                mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            __syncthreads();
            if ((threadIdx.x & 31) == 0)
                s_w[threadIdx.x >> 5] = mx;
            __syncthreads();
            if (threadIdx.x == 0)
            {
                double r = 0.0;
                for (int q = 0; q < kVecBlock / 32; q++)
                    r = fmax(r, s_w[q]);
                out[0] = r;
                *ticket = 0u;
            }
        }
    }
    else
    {
        grid_reduce_fixed_order<kVecBlock>(cta, 1, partials, ticket, out);
    }
}

__global__ void vec_scale_kernel(const double* __restrict__ x, double s, int divide, double* __restrict__ y, int64_t n)
{
    for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t) gridDim.x * blockDim.x)
        y[i] = divide ? x[i] / s : x[i] * s;
}

__global__ void vec_axpy_kernel(const double* __restrict__ w, const double* __restrict__ v, double a, double* __restrict__ f, int64_t n)
{
    for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t) gridDim.x * blockDim.x)
        f[i] = w[i] - v[i] * a;  // f = w - v * H(0,0)   (Arnoldi.h:177)
}

// interleaved complex vectors of n doubles: f = w - v * (ar + i ai)   (Arnoldi.h:177 with a complex Scalar)
__global__ void vec_caxpy_kernel(const double* __restrict__ w, const double* __restrict__ v, double ar, double ai, double* __restrict__ f, int64_t n)
{
    for (int64_t i = 2 * ((int64_t) blockIdx.x * blockDim.x + threadIdx.x); i + 1 < n; i += 2 * (int64_t) gridDim.x * blockDim.x)
    {
        const double vr = v[i], vi = v[i + 1];
        f[i] = w[i] - (vr * ar - vi * ai);
        f[i + 1] = w[i + 1] - (vr * ai + vi * ar);
    }
}

// Complex restart GEMM, last step: A = V Re(Q), B = V Im(Q) were formed by two real GEMMs on the interleaved basis (rows = 2 x complex
// rows); V(:, c) = A(:, c) + i B(:, c), i.e. (a_re - b_im) + i (a_im + b_re), for c < kk.   (Arnoldi.h:320-335 with a complex Q)
__global__ void zcombine_kernel(const double* __restrict__ A, const double* __restrict__ B, int64_t ldab, double* __restrict__ V, int64_t ldv, int64_t n, int kk)
{
    const int64_t pairs = n / 2;
    for (int64_t t = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; t < pairs * kk; t += (int64_t) gridDim.x * blockDim.x)
    {
        const int64_t c = t / pairs, i = 2 * (t % pairs);
        const double ar = A[i + c * ldab], ai = A[i + 1 + c * ldab], br = B[i + c * ldab], bi = B[i + 1 + c * ldab];
        V[i + c * ldv] = ar - bi;
        V[i + 1 + c * ldv] = ai + br;
    }
}

// f <- f * Q(m-1, kk-2) + V(:, kk-1) * H(kk-1, kk-2) with complex coefficients read on the device   (Arnoldi.h:337)
__global__ void zf_update_kernel(double* __restrict__ f, const double* __restrict__ vk, const double* __restrict__ Qr, const double* __restrict__ Qi,
                                 const double* __restrict__ Hr, const double* __restrict__ Hi, int m, int kk, int64_t n)
{
    const double qr = Qr[(m - 1) + (int64_t) (kk - 2) * m], qi = Qi[(m - 1) + (int64_t) (kk - 2) * m];
    const double hr = Hr[(kk - 1) + (int64_t) (kk - 2) * m], hi = Hi[(kk - 1) + (int64_t) (kk - 2) * m];
    for (int64_t i = 2 * ((int64_t) blockIdx.x * blockDim.x + threadIdx.x); i + 1 < n; i += 2 * (int64_t) gridDim.x * blockDim.x)
    {
        const double fr = f[i], fi = f[i + 1], vr = vk[i], vi = vk[i + 1];
        f[i] = (fr * qr - fi * qi) + (vr * hr - vi * hi);
        f[i + 1] = (fr * qi + fi * qr) + (vr * hi + vi * hr);
    }
}

__global__ void set_beta_kernel(FacCtl* ctl, const double* slot, int take_sqrt)
{
    const double v = *slot;
    ctl->beta = take_sqrt ? sqrt(v) : v;
}

__global__ void set_scalar_kernel(double* dst, double v) { *dst = v; }

__global__ void step_scale_kernel(const double* __restrict__ f, const FacCtl* ctl, double* __restrict__ vi, int64_t n)
{
    const double beta = ctl->beta;
    for (int64_t r = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t) gridDim.x * blockDim.x)
        vi[r] = f[r] / beta;
}

template <bool SYM>
__global__ void __launch_bounds__(kVecBlock)
    step_epilogue_kernel(double* __restrict__ w, const double* __restrict__ V, int64_t ldv, int64_t n, FacCtl* ctl, double* H, int m, int i, int restarted,
                         double* partials, unsigned int* ticket)
{
    const double hsub = restarted ? 0.0 : ctl->beta;
    const double* __restrict__ vi = V + (int64_t) i * ldv;
    const double* __restrict__ vp = V + (int64_t) (i - 1) * ldv;
    double part = 0.0;
    if (SYM)
    {
        for (int64_t r = (int64_t) blockIdx.x * kVecBlock + threadIdx.x; r < n; r += (int64_t) gridDim.x * kVecBlock)
        {
            const double wr = w[r] - hsub * vp[r];  // Lanczos.h:139
            w[r] = wr;
            part = fma(vi[r], wr, part);            // Lanczos.h:142
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
    {
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
        __shared__ double s_w[kVecBlock / 32];
        part = warp_sum(part);
        if ((threadIdx.x & 31) == 0)
            s_w[threadIdx.x >> 5] = part;
        __syncthreads();
        double cta = 0.0;
        if (threadIdx.x == 0)
            for (int q = 0; q < kVecBlock / 32; q++)
                cta += s_w[q];
        grid_reduce_fixed_order<kVecBlock>(cta, 1, partials, ticket, ctl->red_a);
    }
}

int vec_grid(int64_t n)
{
    const int sms = device_info().sm_count;
    const int64_t need = (n + kVecBlock - 1) / kVecBlock;
    return (int) std::max<int64_t>(1, std::min<int64_t>(need, (int64_t) sms * 4));
}

}  // namespace

void launch_vec_reduce(int op, const double* x, const double* y, int64_t n, double* out, const RedScratch& rs, cudaStream_t stream)
{
    const int grid = vec_grid(n);
    SB200_REQUIRE(grid <= rs.max_grid, SB200_LOGIC, "vec_reduce: reduction scratch too small");
    switch (op)
    {
        case VR_SUMSQ: vec_reduce_kernel<VR_SUMSQ><<<grid, kVecBlock, 0, stream>>>(x, y, n, out, rs.partials, rs.ticket); break;
        case VR_DOT: vec_reduce_kernel<VR_DOT><<<grid, kVecBlock, 0, stream>>>(x, y, n, out, rs.partials, rs.ticket); break;
        case VR_MAXABS: vec_reduce_kernel<VR_MAXABS><<<grid, kVecBlock, 0, stream>>>(x, y, n, out, rs.partials, rs.ticket); break;
        case VR_CDOT_IM: vec_reduce_kernel<VR_CDOT_IM><<<grid, kVecBlock, 0, stream>>>(x, y, n, out, rs.partials, rs.ticket); break;
        case VR_CMAXABS: vec_reduce_kernel<VR_CMAXABS><<<grid, kVecBlock, 0, stream>>>(x, y, n, out, rs.partials, rs.ticket); break;
        default: throw Error(SB200_LOGIC, "bad reduce op");
    }
    SB200_CUDA_CHECK(cudaGetLastError());
}

void launch_vec_scale(const double* x, double s, int divide, double* y, int64_t n, cudaStream_t stream)
{
    vec_scale_kernel<<<vec_grid(n), kVecBlock, 0, stream>>>(x, s, divide, y, n);
    SB200_CUDA_CHECK(cudaGetLastError());
}

void launch_vec_axpy(const double* w, const double* v, double a, double* f, int64_t n, cudaStream_t stream)
{
    vec_axpy_kernel<<<vec_grid(n), kVecBlock, 0, stream>>>(w, v, a, f, n);
    SB200_CUDA_CHECK(cudaGetLastError());
}

void launch_vec_caxpy(const double* w, const double* v, double ar, double ai, double* f, int64_t n, cudaStream_t stream)
{
    vec_caxpy_kernel<<<vec_grid(n / 2), kVecBlock, 0, stream>>>(w, v, ar, ai, f, n);
    SB200_CUDA_CHECK(cudaGetLastError());
}

void launch_zcombine(const double* A, const double* B, int64_t ldab, double* V, int64_t ldv, int64_t n, int kk, cudaStream_t stream)
{
    zcombine_kernel<<<vec_grid((n / 2) * kk), kVecBlock, 0, stream>>>(A, B, ldab, V, ldv, n, kk);
    SB200_CUDA_CHECK(cudaGetLastError());
}

void launch_zf_update(double* f, const double* vk, const double* Qr, const double* Qi, const double* Hr, const double* Hi, int m, int kk, int64_t n,
                      cudaStream_t stream)
{
    zf_update_kernel<<<vec_grid(n / 2), kVecBlock, 0, stream>>>(f, vk, Qr, Qi, Hr, Hi, m, kk, n);
    SB200_CUDA_CHECK(cudaGetLastError());
}

void launch_step_scale(const double* f, const FacCtl* ctl, double* vi, int64_t n, cudaStream_t stream)
{
    step_scale_kernel<<<vec_grid(n), kVecBlock, 0, stream>>>(f, ctl, vi, n);
    SB200_CUDA_CHECK(cudaGetLastError());
}

void launch_step_epilogue(double* w, const double* V, int64_t ldv, int64_t n, FacCtl* ctl, double* H, int m, int i, int restarted, bool symmetric,
                          const RedScratch& rs, cudaStream_t stream)
{
    const int grid = symmetric ? vec_grid(n) : 1;
    SB200_REQUIRE(grid <= rs.max_grid, SB200_LOGIC, "step epilogue: reduction scratch too small");
    if (symmetric)
        step_epilogue_kernel<true><<<grid, kVecBlock, 0, stream>>>(w, V, ldv, n, ctl, H, m, i, restarted, rs.partials, rs.ticket);
    else
        step_epilogue_kernel<false><<<grid, kVecBlock, 0, stream>>>(w, V, ldv, n, ctl, H, m, i, restarted, rs.partials, rs.ticket);
    SB200_CUDA_CHECK(cudaGetLastError());
}

void launch_set_beta(FacCtl* ctl, const double* red_slot, int take_sqrt, cudaStream_t stream)
{
    set_beta_kernel<<<1, 1, 0, stream>>>(ctl, red_slot, take_sqrt);
    SB200_CUDA_CHECK(cudaGetLastError());
}

void launch_set_scalar(double* dst, double v, cudaStream_t stream)
{
    set_scalar_kernel<<<1, 1, 0, stream>>>(dst, v);
    SB200_CUDA_CHECK(cudaGetLastError());
}

}  // namespace sb200

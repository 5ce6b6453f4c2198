// Microbenchmarks that pin the roofs the SpMV is reported against (tools/gather_roof.py).  Not on the product path.
//
// gather_bench_kernel: every thread performs independent 8-byte read-only loads x[h(k)] with pseudo-random, uniformly distributed
// indices over a vector of n doubles, four in flight, exactly the access a CSR SpMV with uniformly random column ids makes for its
// operand -- without the matrix stream, the reduction and the output.  Its time for nnz gathers over one column-block slice is the
// floor of the SpMV's gather phase: each gather is its own 128 B line (one L1TEX wavefront) and its own 32 B L2 sector.
#include "kernels.h"

namespace sb200 {

namespace {

__device__ __forceinline__ unsigned int mix32(unsigned int h)
{
    // lowbias32 (public-domain integer hash): good avalanche in 2 multiplications
    h ^= h >> 16;
    h *= 0x7feb352du;
    h ^= h >> 15;
    h *= 0x846ca68bu;
    h ^= h >> 16;
    return h;
}

__global__ void __launch_bounds__(256, 8) gather_bench_kernel(const double* __restrict__ x, unsigned int n, int per_thread, double* __restrict__ out)
{
    const uint64_t pol_keep = l2_policy_evict_last();
    const unsigned int tid = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned int k = tid * (unsigned int) per_thread;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int t = 0;
    for (; t + 4 <= per_thread; t += 4, k += 4)
    {
        const unsigned int i0 = __umulhi(mix32(k + 0), n), i1 = __umulhi(mix32(k + 1), n), i2 = __umulhi(mix32(k + 2), n), i3 = __umulhi(mix32(k + 3), n);
        a0 += ld_keep_f64(x + i0, pol_keep);
        a1 += ld_keep_f64(x + i1, pol_keep);
        a2 += ld_keep_f64(x + i2, pol_keep);
        a3 += ld_keep_f64(x + i3, pol_keep);
    }
    for (; t < per_thread; t++, k++)
        a0 += ld_keep_f64(x + __umulhi(mix32(k), n), pol_keep);
    out[tid] = (a0 + a1) + (a2 + a3);
}

// stream_gather_bench_kernel: the same gathers, but the indices and a coefficient per gather are STREAMED from HBM exactly as a sliced
// (ELL-like) SpMV streams its (col, val) arrays -- 12 bytes per gather in fully coalesced 128 B / 256 B warp loads, four steps in flight
// -- and each gather depends on its index load.  No rows, no padding, no reduction tree, no output vector: acc += val[k] * x[col[k]].
// This is the floor of ANY SpMV kernel on this access pattern (stream + dependent gather through one L1TEX and the L2), i.e. the
// practical roof the product kernels are compared with.
__global__ void __launch_bounds__(512, 3) stream_gather_bench_kernel(const int* __restrict__ col, const double* __restrict__ val, const double* __restrict__ x,
                                                                    int64_t count, double* __restrict__ out)
{
    const uint64_t pol_stream = l2_policy_evict_first();
    const uint64_t pol_keep = l2_policy_evict_last();
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t) blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t) gridDim.x * blockDim.x) >> 5;
    // each warp owns contiguous runs of 32 * 20 entries (one "slice" of a 20-per-row matrix), strided over the grid like the windows of the product kernel
    constexpr int RUN = 20;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    const int64_t nruns = count / (32 * RUN);
    for (int64_t r = warp; r < nruns; r += nwarps)
    {
        const int* cp = col + r * (32 * RUN) + lane;
        const double* vp = val + r * (32 * RUN) + lane;
#pragma unroll
        for (int t = 0; t < RUN; t += 4)
        {
            const int c0 = ld_stream_s32(cp + (t + 0) * 32, pol_stream), c1 = ld_stream_s32(cp + (t + 1) * 32, pol_stream);
            const int c2 = ld_stream_s32(cp + (t + 2) * 32, pol_stream), c3 = ld_stream_s32(cp + (t + 3) * 32, pol_stream);
            const double v0 = ld_stream_f64(vp + (t + 0) * 32, pol_stream), v1 = ld_stream_f64(vp + (t + 1) * 32, pol_stream);
            const double v2 = ld_stream_f64(vp + (t + 2) * 32, pol_stream), v3 = ld_stream_f64(vp + (t + 3) * 32, pol_stream);
            a0 = fma(v0, ld_keep_f64(x + c0, pol_keep), a0);
            a1 = fma(v1, ld_keep_f64(x + c1, pol_keep), a1);
            a2 = fma(v2, ld_keep_f64(x + c2, pol_keep), a2);
            a3 = fma(v3, ld_keep_f64(x + c3, pol_keep), a3);
        }
    }
    out[(int64_t) blockIdx.x * blockDim.x + threadIdx.x] = (a0 + a1) + (a2 + a3);
}

// col[k] = uniform pseudo-random index below n (band == 0) or k / 20 + (k % 20) - 10 clamped (band != 0: neighbouring columns, coalesced gathers)
__global__ void fill_stream_kernel(int* col, double* val, int64_t count, unsigned int n, int band)
{
    for (int64_t k = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; k < count; k += (int64_t) gridDim.x * blockDim.x)
    {
        if (band)
        {
            // entry k of the step-major slice layout: run r = k / 640, step t = (k % 640) / 32, lane = k % 32  ->  row = 32 r + lane, column = row + t - 10
            const int64_t r = k / 640, t = (k % 640) / 32, lane = k % 32;
            int64_t c = (32 * r + lane) % n + t - 10;
            c = c < 0 ? 0 : (c >= (int64_t) n ? (int64_t) n - 1 : c);
            col[k] = (int) c;
        }
        else
            col[k] = (int) __umulhi(mix32((unsigned int) k * 2654435761u + (unsigned int) (k >> 32)), n);
        val[k] = 1.0;
    }
}

__global__ void fill_ones_kernel(double* x, int64_t n)
{
    for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t) gridDim.x * blockDim.x)
        x[i] = 1.0;
}

}  // namespace

// `gathers` random 8-byte loads from a vector of n doubles; returns the average time of `repeat` launches (ms) and, in
// *checksum, the sum of everything loaded divided by the number of gathers (1.0 for the vector of ones that is used).
float bench_gather(int64_t n, int64_t gathers, int repeat, double* checksum)
{
    SB200_REQUIRE(n >= 1 && n < (1LL << 31) && gathers >= 1, SB200_INVALID_ARGUMENT, "bench_gather: bad sizes");
    const int sms = device_info().sm_count;
    const int grid = sms * 8, block = 256;
    const int64_t threads = (int64_t) grid * block;
    const int per_thread = (int) std::max<int64_t>(1, (gathers + threads - 1) / threads);
    DevBuf<double> x((size_t) n), out((size_t) threads);
    cudaStream_t st;
    SB200_CUDA_CHECK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    cudaEvent_t e0, e1;
    SB200_CUDA_CHECK(cudaEventCreate(&e0));
    SB200_CUDA_CHECK(cudaEventCreate(&e1));
    fill_ones_kernel<<<sms * 8, 256, 0, st>>>(x.get(), n);
    gather_bench_kernel<<<grid, block, 0, st>>>(x.get(), (unsigned int) n, per_thread, out.get());  // warm-up (fills L2)
    SB200_CUDA_CHECK(cudaEventRecord(e0, st));
    for (int r = 0; r < std::max(repeat, 1); r++)
        gather_bench_kernel<<<grid, block, 0, st>>>(x.get(), (unsigned int) n, per_thread, out.get());
    SB200_CUDA_CHECK(cudaEventRecord(e1, st));
    SB200_CUDA_CHECK(cudaEventSynchronize(e1));
    SB200_CUDA_CHECK(cudaGetLastError());
    float ms = 0.f;
    SB200_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
    std::vector<double> h((size_t) threads);
    SB200_CUDA_CHECK(cudaMemcpy(h.data(), out.get(), sizeof(double) * h.size(), cudaMemcpyDeviceToHost));
    double s = 0.0;
    for (double v : h)
        s += v;
    if (checksum)
        *checksum = s / (double(per_thread) * double(threads));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaStreamDestroy(st);
    return ms / float(std::max(repeat, 1));
}

// `gathers` streamed (index, coefficient) pairs + dependent gathers from a vector of n doubles (see stream_gather_bench_kernel);
// band != 0 uses neighbouring columns instead of uniformly random ones.  Returns the average time of `repeat` launches (ms).
float bench_stream_gather(int64_t n, int64_t gathers, int band, int repeat, double* checksum)
{
    SB200_REQUIRE(n >= 1 && n < (1LL << 31) && gathers >= 640, SB200_INVALID_ARGUMENT, "bench_stream_gather: bad sizes");
    const int sms = device_info().sm_count;
    const int grid = sms * 3, block = 512;
    const int64_t threads = (int64_t) grid * block;
    const int64_t count = gathers / 640 * 640;
    DevBuf<double> x((size_t) n), out((size_t) threads), val((size_t) count);
    DevBuf<int> col((size_t) count);
    cudaStream_t st;
    SB200_CUDA_CHECK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    cudaEvent_t e0, e1;
    SB200_CUDA_CHECK(cudaEventCreate(&e0));
    SB200_CUDA_CHECK(cudaEventCreate(&e1));
    fill_ones_kernel<<<sms * 8, 256, 0, st>>>(x.get(), n);
    fill_stream_kernel<<<sms * 8, 256, 0, st>>>(col.get(), val.get(), count, (unsigned int) n, band);
    stream_gather_bench_kernel<<<grid, block, 0, st>>>(col.get(), val.get(), x.get(), count, out.get());  // warm-up
    SB200_CUDA_CHECK(cudaEventRecord(e0, st));
    for (int r = 0; r < std::max(repeat, 1); r++)
        stream_gather_bench_kernel<<<grid, block, 0, st>>>(col.get(), val.get(), x.get(), count, out.get());
    SB200_CUDA_CHECK(cudaEventRecord(e1, st));
    SB200_CUDA_CHECK(cudaEventSynchronize(e1));
    SB200_CUDA_CHECK(cudaGetLastError());
    float ms = 0.f;
    SB200_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
    std::vector<double> h((size_t) threads);
    SB200_CUDA_CHECK(cudaMemcpy(h.data(), out.get(), sizeof(double) * h.size(), cudaMemcpyDeviceToHost));
    double sum = 0.0;
    for (double v : h)
        sum += v;
    if (checksum)
        *checksum = sum / double(count);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaStreamDestroy(st);
    return ms / float(std::max(repeat, 1));
}

}  // namespace sb200

// Shared host/device helpers for the sm_100a eigensolver library.
#pragma once

#ifdef SB200_EMU
// kernel-logic emulation build (tools/cuda_emu/, test infrastructure only): the CUDA execution model on CPU fibers
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#endif
#include <stdint.h>

#include <cmath>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/spectra_b200.h"

namespace sb200 {

// ---------------------------------------------------------------------------------------------
// Errors: C++ exceptions inside the library, converted to sb200_status at the C ABI.
// ---------------------------------------------------------------------------------------------
struct Error : public std::runtime_error
{
    int status;
    Error(int st, const std::string& msg) : std::runtime_error(msg), status(st) {}
};

void set_last_error(const std::string& msg);

#define SB200_CUDA_CHECK(expr)                                                                                        \
    do                                                                                                                \
    {                                                                                                                 \
        cudaError_t _e = (expr);                                                                                      \
        if (_e != cudaSuccess)                                                                                        \
            throw ::sb200::Error(SB200_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" + __FILE__ + ":" + std::to_string(__LINE__) + ")"); \
    } while (0)

#define SB200_REQUIRE(cond, status, msg)       \
    do                                         \
    {                                          \
        if (!(cond))                           \
            throw ::sb200::Error(status, msg); \
    } while (0)

// ---------------------------------------------------------------------------------------------
// Constants of the reference algorithm (Util/TypeTraits.h:63-74, Arnoldi.h:50-52)
// ---------------------------------------------------------------------------------------------
constexpr double kEps = 2.220446049250313e-16;
constexpr double kMin = 2.2250738585072014e-308;
constexpr double kNear0 = kMin * 10.0;

constexpr int kMaxNcv = 128;  // upper bound on the Krylov dimension handled by the device kernels

// ---------------------------------------------------------------------------------------------
// RAII device buffer
// ---------------------------------------------------------------------------------------------
template <typename T>
struct DevBuf
{
    T* p = nullptr;
    size_t n = 0;
    DevBuf() {}
    explicit DevBuf(size_t count) { alloc(count); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n)
    {
        o.p = nullptr;
        o.n = 0;
    }
    DevBuf& operator=(DevBuf&& o) noexcept
    {
        if (this != &o)
        {
            release();
            p = o.p;
            n = o.n;
            o.p = nullptr;
            o.n = 0;
        }
        return *this;
    }
    ~DevBuf() { release(); }
    void alloc(size_t count)
    {
        release();
        n = count;
        if (count)
            SB200_CUDA_CHECK(cudaMalloc(&p, count * sizeof(T)));
    }
    void release()
    {
        if (p)
            cudaFree(p);
        p = nullptr;
        n = 0;
    }
    void zero(cudaStream_t s) { SB200_CUDA_CHECK(cudaMemsetAsync(p, 0, n * sizeof(T), s)); }
    T* get() const { return p; }
    size_t bytes() const { return n * sizeof(T); }
};

template <typename T>
struct PinnedBuf
{
    T* p = nullptr;
    size_t n = 0;
    PinnedBuf() {}
    explicit PinnedBuf(size_t count) { alloc(count); }
    PinnedBuf(const PinnedBuf&) = delete;
    PinnedBuf& operator=(const PinnedBuf&) = delete;
    ~PinnedBuf()
    {
        if (p)
            cudaFreeHost(p);
    }
    void alloc(size_t count)
    {
        if (p)
            cudaFreeHost(p);
        p = nullptr;
        n = count;
        if (count)
            SB200_CUDA_CHECK(cudaMallocHost(&p, count * sizeof(T)));
    }
    T* get() const { return p; }
};

inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

// ---------------------------------------------------------------------------------------------
// Device info (cached)
// ---------------------------------------------------------------------------------------------
struct DeviceInfo
{
    int device = -1;
    int sm_count = 0;
    int cc_major = 0, cc_minor = 0;
    size_t total_mem = 0;
    int l2_bytes = 0;
};
const DeviceInfo& device_info();  // throws Error(SB200_CUDA) when no usable device

// ---------------------------------------------------------------------------------------------
// Profiling: optional per-kernel-class CUDA-event timing
// ---------------------------------------------------------------------------------------------
enum KernelClass
{
    KC_SPMV = 0,
    KC_PANEL,
    KC_COMPRESS,
    KC_SMALL,
    KC_COMM,
    KC_COUNT
};
extern int g_profiling_level;

struct Profiler
{
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    double ms[KC_COUNT] = {0, 0, 0, 0, 0};
    int64_t launches = 0;
    void reset()
    {
        for (double& m : ms)
            m = 0;
        launches = 0;
    }
    void ensure()
    {
        if (!ev0)
        {
            SB200_CUDA_CHECK(cudaEventCreate(&ev0));
            SB200_CUDA_CHECK(cudaEventCreate(&ev1));
        }
    }
    ~Profiler()
    {
        if (ev0)
            cudaEventDestroy(ev0);
        if (ev1)
            cudaEventDestroy(ev1);
    }
};

struct ScopedKernelTimer
{
    Profiler* prof;
    cudaStream_t stream;
    int cls;
    bool on;
    ScopedKernelTimer(Profiler* p, cudaStream_t s, int c, int nlaunch = 1) : prof(p), stream(s), cls(c), on(p && g_profiling_level > 0)
    {
        if (prof)
            prof->launches += nlaunch;
        if (on)
        {
            prof->ensure();
            cudaEventRecord(prof->ev0, stream);
        }
    }
    ~ScopedKernelTimer()
    {
        if (on)
        {
            cudaEventRecord(prof->ev1, stream);
            cudaEventSynchronize(prof->ev1);
            float t = 0.f;
            cudaEventElapsedTime(&t, prof->ev0, prof->ev1);
            prof->ms[cls] += t;
        }
    }
};

// ---------------------------------------------------------------------------------------------
// Device-side helpers
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int L>
__device__ __forceinline__ double subwarp_sum(double v)
{
#pragma unroll
    for (int o = L / 2; o > 0; o >>= 1)
        v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

#ifdef SB200_EMU
// emulation: cache hints have no meaning on the CPU; the loads are plain loads
__device__ __forceinline__ uint64_t l2_policy_evict_first() { return 0; }
__device__ __forceinline__ uint64_t l2_policy_evict_last() { return 0; }
__device__ __forceinline__ double ld_stream_f64(const double* p, uint64_t) { return *p; }
__device__ __forceinline__ int ld_stream_s32(const int* p, uint64_t) { return *p; }
__device__ __forceinline__ double2 ld_stream_f64x2(const double* p, uint64_t) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ void ld_stream_f64x4(const double* p, double (&v)[4])
{
    for (int q = 0; q < 4; q++)
        v[q] = p[q];
}
__device__ __forceinline__ double ld_keep_f64(const double* p, uint64_t) { return *p; }
__device__ __forceinline__ double2 ld_keep_f64x2(const double* p, uint64_t) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ double ld_cg_f64(const double* p) { return *p; }
// system-scope accesses of the peer-memory kernels: the emulated ranks are OS threads of one process
__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
__device__ __forceinline__ void st_sys_f64(double* p, double v)
{
    unsigned long long b;
    memcpy(&b, &v, sizeof(b));
    __atomic_store_n(reinterpret_cast<unsigned long long*>(p), b, __ATOMIC_RELAXED);
}
__device__ __forceinline__ double ld_sys_f64(const double* p)
{
    const unsigned long long b = __atomic_load_n(reinterpret_cast<const unsigned long long*>(p), __ATOMIC_RELAXED);
    double v;
    memcpy(&v, &b, sizeof(v));
    return v;
}
__device__ __forceinline__ void st_peer_f64x2(double* p, double2 v) { p[0] = v.x; p[1] = v.y; }
#else
// L2 cache policies (createpolicy): streamed-once data is marked evict_first so that the gathered
// operand vector / small reused vectors keep their L2 residency (evict_last).
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
// Streaming (read-once) loads: no L1 allocation + the given L2 policy.
__device__ __forceinline__ double ld_stream_f64(const double* p, uint64_t pol)
{
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ int ld_stream_s32(const int* p, uint64_t pol)
{
    int v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ double2 ld_stream_f64x2(const double* p, uint64_t pol)
{
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(pol));
    return v;
}
// 256-bit streaming load (sm_100+: LDG.E.256), 32-byte aligned address.
__device__ __forceinline__ void ld_stream_f64x4(const double* p, double (&v)[4])
{
    asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
}
// Gathered operand: keep in L2.  SB200_GATHER_MODE (build-time experiment knob): 0 = read-only path with the L2
// evict_last hint (default), 1 = the same without allocating in L1, 2 = plain read-only load, no hints.
#ifndef SB200_GATHER_MODE
#define SB200_GATHER_MODE 0
#endif
__device__ __forceinline__ double ld_keep_f64(const double* p, uint64_t pol)
{
    double v;
#if SB200_GATHER_MODE == 0
    asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
#elif SB200_GATHER_MODE == 1
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
#else
    (void) pol;
    v = __ldg(p);
#endif
    return v;
}
// 16-byte gather (one complex operand entry), kept in L2
__device__ __forceinline__ double2 ld_keep_f64x2(const double* p, uint64_t pol)
{
    double2 v;
    asm volatile("ld.global.nc.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p), "l"(pol));
    return v;
}
// L2-coherent load (skips the non-coherent L1) for data written by other CTAs of the same kernel.
__device__ __forceinline__ double ld_cg_f64(const double* p)
{
    double v;
    asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
// System-scope accesses used on peer (NVLink-mapped) memory: release / acquire flags, relaxed data.
__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_sys_f64(double* p, double v) { asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory"); }
__device__ __forceinline__ double ld_sys_f64(const double* p)
{
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
// 16-byte store of two consecutive rows into a (possibly remote) operand buffer; ordered for the peers by the kernel boundary and the
// system-scope release of the all-reduce that follows
__device__ __forceinline__ void st_peer_f64x2(double* p, double2 v) { asm volatile("st.global.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory"); }
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
    int spins = 0;
    while (!done)
    {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (!done && ++spins > (1 << 26))
            __trap();  // never hang the device on a lost transaction
    }
}
// 1-D bulk TMA copy global -> shared, completion signalled on an mbarrier (bytes: multiple of 16, both addresses 16 B aligned)
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes),
                 "r"(smem_u32(bar))
                 : "memory");
}
#endif  // SB200_EMU

// Grid-wide deterministic reduction of K (<= 128) per-CTA partial values.
//   partials : gridDim.x * 128 doubles (scratch);  ticket : zero-initialised counter, left at zero.
// Every CTA deposits its K partials (thread t holds index t, t < K; BLOCK >= 128).  The CTA that
// arrives last sums them in an order that depends only on (gridDim.x, BLOCK, K<=64): thread (k, s)
// adds CTAs s, s+S, s+2S, ... and the S slices are then added 0..S-1.  Returns true in the last
// CTA (all of its threads) with out[k] written (out may be shared or global memory).
template <int BLOCK>
__device__ __forceinline__ bool grid_reduce_fixed_order(double my_val, int K, double* partials, unsigned int* ticket, double* out)
{
    static_assert(BLOCK % 128 == 0, "BLOCK must be a multiple of 128");
    __shared__ bool s_last;
    __shared__ double s_slice[BLOCK];
    const int tid = threadIdx.x;
    if (tid < K)
        partials[(size_t) blockIdx.x * 128 + tid] = my_val;
    __threadfence();
    __syncthreads();
    if (tid == 0)
    {
        const unsigned int t = atomicAdd(ticket, 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last)
        return false;
    __threadfence();
    const int kp = (K <= 64) ? 64 : 128;  // slots per slice
    const int S = BLOCK / kp;             // number of slices
    const int k = tid & (kp - 1), sl = tid / kp;
    double s = 0.0;
    if (k < K)
    {
        const int nb = gridDim.x;
        int b = sl;
        for (; b + 3 * S < nb; b += 4 * S)
        {
            const double a0 = ld_cg_f64(partials + (size_t) (b) * 128 + k);
            const double a1 = ld_cg_f64(partials + (size_t) (b + S) * 128 + k);
            const double a2 = ld_cg_f64(partials + (size_t) (b + 2 * S) * 128 + k);
            const double a3 = ld_cg_f64(partials + (size_t) (b + 3 * S) * 128 + k);
            s += a0;
            s += a1;
            s += a2;
            s += a3;
        }
        for (; b < nb; b += S)
            s += ld_cg_f64(partials + (size_t) b * 128 + k);
    }
    s_slice[tid] = s;
    __syncthreads();
    if (tid < K)
    {
        double t = s_slice[tid];
        for (int q = 1; q < S; q++)
            t += s_slice[q * kp + tid];
        out[tid] = t;
    }
    if (tid == 0)
        *ticket = 0u;
    __syncthreads();
    return true;
}

#endif  // __CUDACC__

}  // namespace sb200

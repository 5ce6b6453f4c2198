// cuda_emu.h -- a small CUDA-on-CPU execution model, TEST INFRASTRUCTURE ONLY.
//
// Purpose: run the *source* of the sm_100a kernels (spectra_b200/csrc/*.cu) on the CPU to check their logic (indexing, barriers,
// shuffles, reductions, host sequencing) when no GPU is at hand.  tools/cuda_emu/emu_build.py rewrites the `<<< >>>` launches and
// `extern __shared__` declarations, compiles the sources with g++ against this header and links tests/_emu/libspectra_b200_emu.so.
// Nothing in spectra_b200/ loads that library; it is used by `tests/test_emu_*.py` (-m "not gpu") only, exactly like the oracle.
// It is NOT a fallback of the product (the product has none), it says nothing about performance, and it cannot run the kernels that
// are written in PTX (TMA / mbarrier / DMMA in gemm_dmma.cu).
//
// Model: one OS thread; the threads of ONE CTA are ucontext fibers that run round-robin and switch only at synchronisation points
// (__syncthreads, __syncwarp, warp shuffles).  CTAs of a grid run one after the other, so `__shared__` can be a plain static and
// atomics are trivially atomic.  Threads that have returned count as arrived at every later barrier (CUDA semantics).  A round in
// which no fiber makes progress aborts with a message (deadlock = a barrier some thread never reaches).
#pragma once

#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <functional>
#include <mutex>
#include <vector>

#define __CUDACC__ 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __launch_bounds__(...)
#define __shared__ static
#define __constant__ static
#define __align__(n) __attribute__((aligned(n)))
#ifndef __restrict__
#define __restrict__ __restrict
#endif

// ---------------------------------------------------------------------------------------------
// vector types
// ---------------------------------------------------------------------------------------------
struct uint3
{
    unsigned int x, y, z;
};
struct dim3
{
    unsigned int x, y, z;
    dim3(unsigned int x_ = 1, unsigned int y_ = 1, unsigned int z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct alignas(16) double2
{
    double x, y;
};
struct alignas(16) int4
{
    int x, y, z, w;
};
struct alignas(8) int2
{
    int x, y;
};
inline double2 make_double2(double x, double y) { return double2{x, y}; }
inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
inline int2 make_int2(int x, int y) { return int2{x, y}; }

// Fiber switch.  x86-64: a 12-instruction register switch (callee-saved registers + stack pointer; defined once in the emu_stubs
// translation unit) -- glibc's swapcontext costs two rt_sigprocmask system calls per switch, which dominated the emulated run time.
// Elsewhere: ucontext.
#if defined(__x86_64__) && !defined(CUDA_EMU_USE_UCONTEXT)
#define CUDA_EMU_ASM_SWITCH 1
extern "C" void cuda_emu_ctx_switch(void** save_sp, void* load_sp);
#endif

namespace emu {

#ifdef CUDA_EMU_ASM_SWITCH
struct Ctx
{
    void* sp = nullptr;
};
inline void ctx_switch(Ctx* from, Ctx* to) { cuda_emu_ctx_switch(&from->sp, to->sp); }
inline void ctx_make(Ctx* c, char* stack, size_t bytes, void (*entry)())
{
    // initial frame: six callee-saved register slots, the entry address that `ret` pops, one fake return-address slot;
    // at entry %rsp must be 8 modulo 16 (as after a call)
    uintptr_t top = ((uintptr_t) stack + bytes) & ~(uintptr_t) 15;
    void** slot = (void**) (top - 16);
    slot[0] = (void*) entry;
    slot[1] = nullptr;
    void** sp = slot - 6;
    for (int q = 0; q < 6; q++)
        sp[q] = nullptr;
    c->sp = sp;
}
#else
struct Ctx
{
    ucontext_t uc;
};
inline void ctx_switch(Ctx* from, Ctx* to) { swapcontext(&from->uc, &to->uc); }
inline void ctx_make(Ctx* c, char* stack, size_t bytes, void (*entry)())
{
    getcontext(&c->uc);
    c->uc.uc_stack.ss_sp = stack;
    c->uc.uc_stack.ss_size = bytes;
    c->uc.uc_link = nullptr;
    makecontext(&c->uc, entry, 0);
}
#endif

struct Warp
{
    uint32_t live = 0;      // lanes that have not returned
    uint32_t mask = 0;      // mask of the collective in progress
    int gen = 0;            // completed collectives
    int arrived = 0;
    uint64_t buf[2][32];
};

struct State
{
    // current thread / CTA
    uint3 tid{0, 0, 0}, bid{0, 0, 0};
    dim3 bdim, gdim;
    // CTA barrier
    int live = 0, bar_count = 0, bar_gen = 0;
    std::vector<Warp> warps;
    unsigned char* dyn_smem = nullptr;
    // fibers
    Ctx main_ctx;
    std::vector<Ctx> ctx;
    std::vector<char*> stacks;
    std::vector<char> done;
    int cur = -1;
    uint64_t progress = 0;
    const std::function<void()>* body = nullptr;
    int64_t launches = 0;
    bool reverse_order = [] { const char* e = getenv("CUDA_EMU_ORDER"); return e && e[0] == 'r' && e[1] == 'e'; }();
    int order_mode = [] { const char* e = getenv("CUDA_EMU_ORDER"); return (e && e[0] == 'r' && e[1] == 'o') ? 2 : 0; }();  // "rotate"
    uint64_t round = 0;
};
// One State per OS thread: the ranks of an emulated multi-GPU run are threads of one process (emu_comm.cpp).  Kernel launches of
// different ranks are serialised by launch_mutex because `__shared__` variables are process-wide statics.
inline thread_local State g;
inline std::mutex launch_mutex;

constexpr size_t kStackBytes = 256 * 1024;

inline void yield_to_main() { ctx_switch(&g.ctx[(size_t) g.cur], &g.main_ctx); }

inline void warp_try_release(Warp& w)
{
    if (w.arrived > 0 && w.arrived == __builtin_popcount(w.mask & w.live))
    {
        w.arrived = 0;
        w.gen++;
        g.progress++;
    }
}

inline void cta_try_release()
{
    if (g.bar_count > 0 && g.bar_count >= g.live)
    {
        g.bar_count = 0;
        g.bar_gen++;
        g.progress++;
    }
}

inline void fiber_entry()
{
    (*g.body)();
    // thread exit: counts as arrived everywhere from now on
    const int t = g.cur;
    g.done[(size_t) t] = 1;
    g.live--;
    Warp& w = g.warps[(size_t) t / 32];
    w.live &= ~(1u << (t & 31));
    g.progress++;
    warp_try_release(w);
    cta_try_release();
    ctx_switch(&g.ctx[(size_t) t], &g.main_ctx);
    abort();  // a finished fiber is never resumed
}

inline void syncthreads()
{
    const int my = g.bar_gen;
    g.bar_count++;
    g.progress++;
    cta_try_release();
    while (g.bar_gen == my)
        yield_to_main();
}

// all lanes in (mask & live) deposit `bits`; returns the generation the values were written in
inline int warp_collective(uint32_t mask, uint64_t bits)
{
    const int t = g.cur, lane = t & 31;
    Warp& w = g.warps[(size_t) t / 32];
    const int my = w.gen;
    w.buf[my & 1][lane] = bits;
    w.mask = mask;
    w.arrived++;
    g.progress++;
    warp_try_release(w);
    while (w.gen == my)
        yield_to_main();
    return my;
}

template <typename T>
inline uint64_t to_bits(T v)
{
    static_assert(sizeof(T) <= 8, "shuffle of a type wider than 64 bits");
    uint64_t b = 0;
    memcpy(&b, &v, sizeof(T));
    return b;
}
template <typename T>
inline T from_bits(uint64_t b)
{
    T v;
    memcpy(&v, &b, sizeof(T));
    return v;
}

template <typename T>
inline T shfl_from(uint32_t mask, T v, int src_lane)
{
    const int my = warp_collective(mask, to_bits(v));
    const Warp& w = g.warps[(size_t) g.cur / 32];
    return from_bits<T>(w.buf[my & 1][src_lane & 31]);
}

struct Launch
{
    dim3 grid, block;
    size_t smem;
};

inline void run_cta(const std::function<void()>& body, const Launch& L)
{
    const int nthreads = (int) (L.block.x * L.block.y * L.block.z);
    g.bdim = L.block;
    g.gdim = L.grid;
    g.live = nthreads;
    g.bar_count = 0;
    g.bar_gen = 0;
    g.warps.assign((size_t) (nthreads + 31) / 32, Warp());
    for (int t = 0; t < nthreads; t++)
        g.warps[(size_t) t / 32].live |= 1u << (t & 31);
    if ((int) g.ctx.size() < nthreads)
    {
        g.ctx.resize((size_t) nthreads);
        while ((int) g.stacks.size() < nthreads)
            g.stacks.push_back((char*) malloc(kStackBytes));
    }
    g.done.assign((size_t) nthreads, 0);
    g.body = &body;
    for (int t = 0; t < nthreads; t++)
    {
        ctx_make(&g.ctx[(size_t) t], g.stacks[(size_t) t], kStackBytes, fiber_entry);
    }
    int remaining = nthreads;
    while (remaining > 0)
    {
        const uint64_t before = g.progress;
        g.round++;
        remaining = 0;
        for (int tt = 0; tt < nthreads; tt++)
        {
            // CUDA_EMU_ORDER=reverse runs the fibers of a CTA in descending thread order: results that change with the order expose
            // code that relies on warp-lockstep execution between two synchronisation points (a race under independent thread scheduling)
            // order 2 ("rotate"): ascending from a start thread that moves every scheduling round, so that neither end is always first
            const int t = g.order_mode == 2 ? (tt + (int) ((g.round * 13) % (uint64_t) nthreads)) % nthreads : (g.reverse_order ? nthreads - 1 - tt : tt);
            if (g.done[(size_t) t])
                continue;
            g.cur = t;
            g.tid.x = (unsigned) t % L.block.x;
            g.tid.y = ((unsigned) t / L.block.x) % L.block.y;
            g.tid.z = (unsigned) t / (L.block.x * L.block.y);
            ctx_switch(&g.main_ctx, &g.ctx[(size_t) t]);
            if (!g.done[(size_t) t])
                remaining++;
        }
        if (remaining > 0 && g.progress == before)
        {
            fprintf(stderr, "cuda_emu: deadlock in CTA (%u,%u,%u): %d threads wait at a barrier/shuffle nobody else reaches\n", g.bid.x, g.bid.y, g.bid.z, remaining);
            abort();
        }
    }
    g.cur = -1;
}

inline void launch_impl(const Launch& L, const std::function<void()>& body)
{
    std::lock_guard<std::mutex> lock(launch_mutex);
    g.launches++;
    // dynamic shared memory: 128-byte aligned, NaN-filled (like device memory here), followed by a guard zone that must stay intact
    constexpr size_t kGuard = 4096;
    std::vector<unsigned char> smem(L.smem + 128 + kGuard);
    unsigned char* sp = smem.data();
    sp += (128 - ((uintptr_t) sp & 127)) & 127;
    memset(sp, 0xFF, L.smem);
    memset(sp + L.smem, 0xA5, kGuard);
    g.dyn_smem = sp;
    for (unsigned bz = 0; bz < L.grid.z; bz++)
        for (unsigned by = 0; by < L.grid.y; by++)
            for (unsigned bx = 0; bx < L.grid.x; bx++)
            {
                g.bid = uint3{bx, by, bz};
                run_cta(body, L);
            }
    for (size_t q = 0; q < kGuard; q++)
        if (sp[L.smem + q] != 0xA5)
        {
            fprintf(stderr, "cuda_emu: a kernel wrote %zu bytes past its %zu bytes of dynamic shared memory\n", q + 1, L.smem);
            abort();
        }
    g.dyn_smem = nullptr;
}

template <typename F>
inline void launch(dim3 grid, dim3 block, F&& f)
{
    launch_impl(Launch{grid, block, 0}, std::function<void()>(f));
}
template <typename F>
inline void launch(dim3 grid, dim3 block, size_t smem, F&& f)
{
    launch_impl(Launch{grid, block, smem}, std::function<void()>(f));
}
template <typename S, typename F>
inline void launch(dim3 grid, dim3 block, size_t smem, S /*stream*/, F&& f)
{
    launch_impl(Launch{grid, block, smem}, std::function<void()>(f));
}

}  // namespace emu

#define threadIdx (::emu::g.tid)
#define blockIdx (::emu::g.bid)
#define blockDim (::emu::g.bdim)
#define gridDim (::emu::g.gdim)

// ---------------------------------------------------------------------------------------------
// device builtins
// ---------------------------------------------------------------------------------------------
inline void __syncthreads() { ::emu::syncthreads(); }
inline void __syncwarp(unsigned mask = 0xffffffffu) { ::emu::warp_collective(mask, 0); }
inline void __threadfence() {}
inline void __threadfence_block() {}
inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }  // emulated ranks are OS threads: a real fence

template <typename T>
inline T __shfl_sync(unsigned mask, T v, int src, int width = 32)
{
    const int lane = ::emu::g.cur & 31;
    const int base = lane & ~(width - 1);
    return ::emu::shfl_from<T>(mask, v, base + (src & (width - 1)));
}
// warp vote: every participating lane deposits its predicate; the result is assembled from the collective's buffer
inline unsigned __ballot_sync(unsigned mask, int pred)
{
    const int my = ::emu::warp_collective(mask, pred ? 1u : 0u);
    const ::emu::Warp& w = ::emu::g.warps[(size_t) ::emu::g.cur / 32];
    unsigned r = 0;
    for (int l = 0; l < 32; l++)
        if (((mask & w.live) >> l) & 1u)
            r |= (w.buf[my & 1][l] ? 1u : 0u) << l;
    return r;
}
inline int __clz(int x) { return x == 0 ? 32 : __builtin_clz((unsigned) x); }
template <typename T>
inline T __shfl_xor_sync(unsigned mask, T v, int lane_mask, int width = 32)
{
    const int lane = ::emu::g.cur & 31;
    const int src = lane ^ lane_mask;
    const int base = lane & ~(width - 1);
    // a source outside the lane's own segment returns the lane's own value
    const T r = ::emu::shfl_from<T>(mask, v, (src >= base && src < base + width) ? src : lane);
    return r;
}
template <typename T>
inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32)
{
    const int lane = ::emu::g.cur & 31;
    const int base = lane & ~(width - 1);
    const int src = lane - (int) delta;
    return ::emu::shfl_from<T>(mask, v, src >= base ? src : lane);
}
template <typename T>
inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32)
{
    const int lane = ::emu::g.cur & 31;
    const int base = lane & ~(width - 1);
    const int src = lane + (int) delta;
    return ::emu::shfl_from<T>(mask, v, src < base + width ? src : lane);
}

template <typename T>
inline T __ldg(const T* p)
{
    return *p;
}
inline unsigned int __umulhi(unsigned int a, unsigned int b) { return (unsigned int) (((uint64_t) a * (uint64_t) b) >> 32); }
inline double rsqrt(double x) { return 1.0 / sqrt(x); }
// CUDA's global min / max overloads
#define EMU_MINMAX(T)                               \
    inline T min(T a, T b) { return b < a ? b : a; } \
    inline T max(T a, T b) { return a < b ? b : a; }
EMU_MINMAX(int)
EMU_MINMAX(unsigned int)
EMU_MINMAX(long)
EMU_MINMAX(unsigned long)
EMU_MINMAX(long long)
EMU_MINMAX(unsigned long long)
EMU_MINMAX(float)
EMU_MINMAX(double)
#undef EMU_MINMAX
inline int __popc(unsigned int x) { return __builtin_popcount(x); }

template <typename T>
inline T atomicAdd(T* p, T v)
{
    const T old = *p;
    *p = old + v;
    return old;
}
inline unsigned int atomicAdd(unsigned int* p, int v) { return atomicAdd<unsigned int>(p, (unsigned int) v); }
template <typename T>
inline T atomicMax(T* p, T v)
{
    const T old = *p;
    if (v > old)
        *p = v;
    return old;
}

// ---------------------------------------------------------------------------------------------
// runtime API subset (synchronous; "device memory" is host memory filled with 0xFF = NaN doubles / -1 ints)
// ---------------------------------------------------------------------------------------------
typedef int cudaError_t;
constexpr cudaError_t cudaSuccess = 0;
constexpr cudaError_t cudaErrorMemoryAllocation = 2;
struct CUstream_st
{
    int dummy;
};
typedef CUstream_st* cudaStream_t;
struct CUevent_st
{
    std::chrono::steady_clock::time_point t;
};
typedef CUevent_st* cudaEvent_t;
enum cudaMemcpyKind
{
    cudaMemcpyHostToHost = 0,
    cudaMemcpyHostToDevice = 1,
    cudaMemcpyDeviceToHost = 2,
    cudaMemcpyDeviceToDevice = 3,
    cudaMemcpyDefault = 4
};
constexpr unsigned int cudaStreamNonBlocking = 1;
constexpr unsigned int cudaEventDisableTiming = 2;
enum cudaFuncAttribute
{
    cudaFuncAttributeMaxDynamicSharedMemorySize = 8
};
struct cudaDeviceProp
{
    char name[256];
    int multiProcessorCount, major, minor, l2CacheSize;
    size_t totalGlobalMem;
};

inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "cuda_emu error"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int* n)
{
    *n = 1;
    return cudaSuccess;
}
inline cudaError_t cudaGetDevice(int* d)
{
    *d = 0;
    return cudaSuccess;
}
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int)
{
    memset(p, 0, sizeof(*p));
    strcpy(p->name, "cuda_emu (CPU, test infrastructure)");
    const char* e = getenv("CUDA_EMU_SMS");
    p->multiProcessorCount = e ? atoi(e) : 2;
    p->major = 10;
    p->minor = 0;
    p->l2CacheSize = 126 << 20;
    p->totalGlobalMem = (size_t) 8 << 30;
    return cudaSuccess;
}
// "Device" allocations carry a 64-byte guard zone on either side (0xA5) that cudaFree verifies: an out-of-bounds WRITE of any kernel
// or copy aborts the test run with the size of the block (reads are not detected).  Layout: [size_t size | pad][guard][user bytes][guard]
constexpr size_t kMemGuard = 64;
template <typename T>
inline cudaError_t cudaMalloc(T** p, size_t bytes)
{
    unsigned char* raw = (unsigned char*) malloc(bytes + 3 * kMemGuard);
    if (!raw)
        return cudaErrorMemoryAllocation;
    memcpy(raw, &bytes, sizeof(size_t));
    memset(raw + kMemGuard, 0xA5, kMemGuard);
    memset(raw + 2 * kMemGuard, 0xFF, bytes);
    memset(raw + 2 * kMemGuard + bytes, 0xA5, kMemGuard);
    *p = (T*) (raw + 2 * kMemGuard);
    return cudaSuccess;
}
inline cudaError_t cudaFree(void* p)
{
    if (!p)
        return cudaSuccess;
    unsigned char* user = (unsigned char*) p;
    unsigned char* raw = user - 2 * kMemGuard;
    size_t bytes = 0;
    memcpy(&bytes, raw, sizeof(size_t));
    for (size_t q = 0; q < kMemGuard; q++)
        if (raw[kMemGuard + q] != 0xA5 || user[bytes + q] != 0xA5)
        {
            fprintf(stderr, "cuda_emu: out-of-bounds write detected around a device allocation of %zu bytes (%s the block)\n", bytes,
                    raw[kMemGuard + q] != 0xA5 ? "before" : "after");
            abort();
        }
    free(raw);
    return cudaSuccess;
}
template <typename T>
inline cudaError_t cudaMallocHost(T** p, size_t bytes)
{
    *p = (T*) malloc(bytes ? bytes : 1);
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
inline cudaError_t cudaFreeHost(void* p)
{
    free(p);
    return cudaSuccess;
}
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind)
{
    memmove(d, s, n);
    return cudaSuccess;
}
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind k, cudaStream_t = nullptr) { return cudaMemcpy(d, s, n, k); }
inline cudaError_t cudaMemcpy2D(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, cudaMemcpyKind)
{
    for (size_t r = 0; r < h; r++)
        memmove((char*) d + r * dp, (const char*) s + r * sp, w);
    return cudaSuccess;
}
inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, cudaMemcpyKind k, cudaStream_t = nullptr)
{
    return cudaMemcpy2D(d, dp, s, sp, w, h, k);
}
inline cudaError_t cudaMemset(void* p, int v, size_t n)
{
    memset(p, v, n);
    return cudaSuccess;
}
inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t = nullptr) { return cudaMemset(p, v, n); }
// a fixed, generous figure: the emulator has no device memory of its own (host allocations)
inline cudaError_t cudaMemGetInfo(size_t* free_b, size_t* total_b)
{
    *free_b = (size_t) 8 << 30;
    *total_b = (size_t) 16 << 30;
    return cudaSuccess;
}
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned int)
{
    *s = new CUstream_st{0};
    return cudaSuccess;
}
inline cudaError_t cudaStreamDestroy(cudaStream_t s)
{
    delete s;
    return cudaSuccess;
}
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e)
{
    *e = new CUevent_st{std::chrono::steady_clock::now()};
    return cudaSuccess;
}
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned int) { return cudaEventCreate(e); }
inline cudaError_t cudaEventDestroy(cudaEvent_t e)
{
    delete e;
    return cudaSuccess;
}
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr)
{
    e->t = std::chrono::steady_clock::now();
    return cudaSuccess;
}
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b)
{
    *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count();
    return cudaSuccess;
}
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned int = 0) { return cudaSuccess; }
template <typename F>
inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int)
{
    return cudaSuccess;
}

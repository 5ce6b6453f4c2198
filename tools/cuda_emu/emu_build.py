"""Builds tests/_emu/libspectra_b200_emu.so: the library's CUDA sources compiled for the CPU against tools/cuda_emu/cuda_emu.h.

TEST INFRASTRUCTURE ONLY (see cuda_emu.h).  The sources are the product's own .cu files, rewritten textually in two places that
are not C++:  `kernel<<<grid, block[, smem[, stream]]>>>(args)`  ->  `::emu::launch(grid, block[, smem[, stream]], [&]() { kernel(args); })`
and  `extern __shared__ T name[];`  ->  `T* name = reinterpret_cast<T*>(::emu::g.dyn_smem);`.
gemm_dmma.cu (TMA / mbarrier / DMMA in PTX) cannot be emulated; its entry point is replaced by the FMA restart GEMM of panel.cu.

    python tools/cuda_emu/emu_build.py [--force]
"""
from __future__ import annotations

import os
import re
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "spectra_b200", "csrc")
OUT = os.path.join(ROOT, "tests", "_emu")
SRC_OUT = os.path.join(OUT, "src")
LIB = os.path.join(OUT, "libspectra_b200_emu.so")
SKIP = {"gemm_dmma.cu", "comm.cu"}
CXX = os.environ.get("CXX", "g++")
CXXFLAGS = ["-O1", "-g", "-std=c++17", "-fPIC", "-fvisibility=hidden", "-DSB200_EMU", "-w", f"-I{HERE}", f"-I{CSRC}"]

STUBS = r'''
// emulation stub for gemm_dmma.cu (PTX kernels cannot run under cuda_emu): the FMA restart GEMM of panel.cu stands in
#include "kernels.h"
namespace sb200 {
void launch_compress_dmma(const double* V, int64_t ldv, int64_t nrows, int m, const double* Q, int kk, double* Vout, int64_t ldo, double* f, const double* H,
                          double* red_out, const RedScratch& rs, cudaStream_t stream)
{
    launch_compress_fma(V, ldv, nrows, m, Q, kk, Vout, ldo, f, H, red_out, rs, stream);
}
}  // namespace sb200

// ---- emulation of comm.cu (NCCL binding): the ranks of a communicator are OS threads of this process -------------------------
// Collectives rendezvous on a shared "world" keyed by the 128-byte unique id: publish the send pointer, barrier, every rank computes its
// result from all ranks' buffers in rank order (so results are bitwise identical on every rank, like NCCL's), barrier, write back.
#include <condition_variable>
#include <cstring>
#include <map>
#include <memory>
#include "host.h"
namespace sb200 {
namespace {
struct EmuWorld
{
    int nranks = 0, arrived = 0, gen = 0;
    std::mutex mu;
    std::condition_variable cv;
    std::vector<const double*> send;
    void barrier()
    {
        std::unique_lock<std::mutex> lk(mu);
        const int my = gen;
        if (++arrived == nranks)
        {
            arrived = 0;
            gen++;
            cv.notify_all();
        }
        else
            cv.wait(lk, [&] { return gen != my; });
    }
};
std::mutex g_worlds_mu;
std::map<uint64_t, std::shared_ptr<EmuWorld>> g_worlds;
uint64_t g_next_id = 1;
struct EmuComm
{
    std::shared_ptr<EmuWorld> w;
};
EmuWorld& world_of(sb200_comm* c) { return *static_cast<EmuComm*>(c->nccl)->w; }
}  // namespace

void nccl_unique_id(void* id128)
{
    std::lock_guard<std::mutex> lk(g_worlds_mu);
    memset(id128, 0, 128);
    const uint64_t id = g_next_id++;
    memcpy(id128, &id, sizeof(id));
}
void nccl_comm_init(sb200_comm* c, const void* id128)
{
    uint64_t id = 0;
    memcpy(&id, id128, sizeof(id));
    std::lock_guard<std::mutex> lk(g_worlds_mu);
    std::shared_ptr<EmuWorld>& w = g_worlds[id];
    if (!w)
    {
        w = std::make_shared<EmuWorld>();
        w->nranks = c->nranks;
        w->send.assign((size_t) c->nranks, nullptr);
    }
    SB200_REQUIRE(w->nranks == c->nranks, SB200_NCCL, "emulated communicator: rank count mismatch");
    c->nccl = new EmuComm{w};
}
void nccl_comm_destroy(sb200_comm* c)
{
    delete static_cast<EmuComm*>(c->nccl);
    c->nccl = nullptr;
}
static void allreduce(sb200_comm* c, double* buf, size_t count, bool is_max)
{
    EmuWorld& w = world_of(c);
    w.send[(size_t) c->rank] = buf;
    w.barrier();
    std::vector<double> tmp(count);
    for (size_t i = 0; i < count; i++)
    {
        double a = w.send[0][i];
        for (int r = 1; r < w.nranks; r++)
            a = is_max ? std::max(a, w.send[(size_t) r][i]) : a + w.send[(size_t) r][i];
        tmp[i] = a;
    }
    w.barrier();
    memcpy(buf, tmp.data(), sizeof(double) * count);
    w.barrier();
}
// peer windows: the "peers" are threads of this process, so a window is a plain allocation whose address every rank learns through the
// world; the peer kernels (peer.cu) then run unchanged, their system-scope flags being atomics of this process
bool peer_window_create(sb200_comm* c, size_t bytes, PeerWindow& w, cudaStream_t)
{
    if (const char* e = std::getenv("SB200_EMU_NO_PEER"))
        if (e[0] == '1')
            return false;  // test knob: behave like a node without peer access (NCCL path)
    EmuWorld& wd = world_of(c);
    void* local = nullptr;
    SB200_CUDA_CHECK(cudaMalloc(&local, bytes));
    memset(local, 0, bytes);
    wd.send[(size_t) c->rank] = static_cast<const double*>(local);
    wd.barrier();
    w.local = local;
    w.bytes = bytes;
    w.nranks = c->nranks;
    for (int r = 0; r < c->nranks; r++)
        w.peer[r] = const_cast<double*>(wd.send[(size_t) r]);
    wd.barrier();
    return true;
}
void peer_window_destroy(sb200_comm*, PeerWindow& w)
{
    if (w.local)
        cudaFree(w.local);
    w = PeerWindow();
}
void nccl_allreduce_sum(sb200_comm* c, double* buf, size_t count, cudaStream_t) { allreduce(c, buf, count, false); }
void nccl_allreduce_max(sb200_comm* c, double* buf, size_t count, cudaStream_t) { allreduce(c, buf, count, true); }
void nccl_allgather(sb200_comm* c, const double* send, double* recv, size_t count_per_rank, cudaStream_t)
{
    EmuWorld& w = world_of(c);
    w.send[(size_t) c->rank] = send;
    w.barrier();
    // send may alias a slice of recv on the calling rank only; other ranks' send buffers are disjoint from this recv
    for (int r = 0; r < w.nranks; r++)
        if (w.send[(size_t) r] != recv + (size_t) r * count_per_rank)
            memmove(recv + (size_t) r * count_per_rank, w.send[(size_t) r], sizeof(double) * count_per_rank);
    w.barrier();
}
}  // namespace sb200

// fiber switch for x86-64 (cuda_emu.h, CUDA_EMU_ASM_SWITCH): save the callee-saved registers and the stack pointer of the running fiber,
// load the other one's.  void cuda_emu_ctx_switch(void** save_sp, void* load_sp)
#ifdef CUDA_EMU_ASM_SWITCH
asm(R"(
    .text
    .globl cuda_emu_ctx_switch
    .hidden cuda_emu_ctx_switch
    .type cuda_emu_ctx_switch,@function
cuda_emu_ctx_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
    .size cuda_emu_ctx_switch,.-cuda_emu_ctx_switch
    .section .note.GNU-stack,"",@progbits
    .text
)");
#endif

// test control: run the fibers of every CTA in descending thread order (see cuda_emu.h, CUDA_EMU_ORDER)
// mode 0: ascending thread order, 1: descending, 2: ascending from a start thread that rotates every scheduling round
extern "C" __attribute__((visibility("default"))) void cuda_emu_set_reverse(int mode)
{
    ::emu::g.reverse_order = (mode == 1);
    ::emu::g.order_mode = (mode == 2) ? 2 : 0;
}
'''


def _match_back_angle(s: str, i: int) -> int:
    """s[i] == '>': index of the matching '<' (balanced), scanning backwards."""
    depth = 0
    while i >= 0:
        if s[i] == ">":
            depth += 1
        elif s[i] == "<":
            depth -= 1
            if depth == 0:
                return i
        i -= 1
    raise ValueError("unbalanced template brackets before <<<")


def _match_paren(s: str, i: int) -> int:
    """s[i] == '(': index of the matching ')'."""
    depth = 0
    while i < len(s):
        if s[i] == "(":
            depth += 1
        elif s[i] == ")":
            depth -= 1
            if depth == 0:
                return i
        i += 1
    raise ValueError("unbalanced parentheses after >>>")


def rewrite_launches(s: str) -> str:
    out = []
    pos = 0
    while True:
        k = s.find("<<<", pos)
        if k < 0:
            out.append(s[pos:])
            break
        # kernel expression before <<<
        j = k - 1
        while s[j].isspace():
            j -= 1
        if s[j] == ">":
            j = _match_back_angle(s, j) - 1
        while j >= 0 and (s[j].isalnum() or s[j] in "_:"):
            j -= 1
        kern = s[j + 1:k].strip()
        e = s.find(">>>", k)
        cfg = s[k + 3:e].strip()
        a0 = e + 3
        while s[a0].isspace() or s[a0] == "\\":
            a0 += 1
        assert s[a0] == "(", f"expected '(' after >>> near: {s[k - 40:k + 80]!r}"
        a1 = _match_paren(s, a0)
        args = s[a0 + 1:a1]
        out.append(s[pos:j + 1])
        out.append(f"::emu::launch({cfg}, [&]() {{ {kern}({args}); }})")
        pos = a1 + 1
    return "".join(out)


_EXT_SH = re.compile(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?([\w:]+(?:\s+[\w:]+)*?)\s+(\w+)\s*\[\s*\]\s*;")


def rewrite(s: str) -> str:
    s = _EXT_SH.sub(lambda m: f"{m.group(1)}* {m.group(2)} = reinterpret_cast<{m.group(1)}*>(::emu::g.dyn_smem);", s)
    return rewrite_launches(s)


def _compile(name: str, text: str, force: bool) -> str:
    src = os.path.join(SRC_OUT, name + ".cpp")
    obj = os.path.join(SRC_OUT, name + ".o")
    old = open(src).read() if os.path.exists(src) else None
    if old != text:
        with open(src, "w") as f:
            f.write(text)
    deps = [src, os.path.join(HERE, "cuda_emu.h")] + [os.path.join(CSRC, h) for h in os.listdir(CSRC) if h.endswith((".h", ".cuh"))]
    deps.append(os.path.join(ROOT, "include", "spectra_b200.h"))
    if force or old != text or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(d) for d in deps):
        r = subprocess.run([CXX, *CXXFLAGS, "-c", src, "-o", obj], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"emu compile failed for {name}:\n{r.stderr[:6000]}")
    return obj


def build(force: bool = False) -> str:
    os.makedirs(SRC_OUT, exist_ok=True)
    jobs = [("emu_stubs", STUBS)]
    for f in sorted(os.listdir(CSRC)):
        if f.endswith(".cu") and f not in SKIP:
            jobs.append((f[:-3], f'#line 1 "{os.path.join(CSRC, f)}"\n' + rewrite(open(os.path.join(CSRC, f)).read())))
    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(lambda j: _compile(j[0], j[1], force), jobs))
    if force or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        r = subprocess.run([CXX, "-shared", "-o", LIB, *objs, "-ldl", "-lpthread"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"emu link failed:\n{r.stderr[:4000]}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))

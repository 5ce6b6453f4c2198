// NCCL binding (row-sharded multi-GPU runs).  libnccl.so.2 is resolved with dlopen
// at first use: a process that already loaded the library (e.g. through torch) shares it, and a
// single-GPU run never touches NCCL.
#include <dlfcn.h>

#include <cstring>
#include <mutex>
#include <vector>

#include "host.h"

namespace sb200 {

namespace {

// Minimal NCCL ABI (stable across 2.x): opaque comm, 128-byte unique id, enum values from nccl.h.
typedef struct
{
    char internal[128];
} NcclUniqueId;
typedef void* NcclComm;
constexpr int kNcclFloat64 = 8;  // ncclDouble
constexpr int kNcclSum = 0;      // ncclSum
constexpr int kNcclMax = 2;      // ncclMax

struct NcclApi
{
    void* handle = nullptr;
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};

NcclApi& api()
{
    static NcclApi a;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* nm : names)
        {
            a.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (a.handle)
                break;
        }
        if (!a.handle)
            return;
        a.GetUniqueId = (int (*)(NcclUniqueId*)) dlsym(a.handle, "ncclGetUniqueId");
        a.CommInitRank = (int (*)(NcclComm*, int, NcclUniqueId, int)) dlsym(a.handle, "ncclCommInitRank");
        a.CommDestroy = (int (*)(NcclComm)) dlsym(a.handle, "ncclCommDestroy");
        a.AllReduce = (int (*)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t)) dlsym(a.handle, "ncclAllReduce");
        a.AllGather = (int (*)(const void*, void*, size_t, int, NcclComm, cudaStream_t)) dlsym(a.handle, "ncclAllGather");
        a.GetErrorString = (const char* (*) (int) ) dlsym(a.handle, "ncclGetErrorString");
    });
    if (!a.handle || !a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.AllReduce || !a.AllGather)
        throw Error(SB200_NCCL, "NCCL (libnccl.so.2) could not be loaded");
    return a;
}

void check(int rc, const char* what)
{
    if (rc != 0)
    {
        const char* msg = api().GetErrorString ? api().GetErrorString(rc) : "?";
        throw Error(SB200_NCCL, std::string(what) + ": " + msg);
    }
}

}  // namespace

void nccl_unique_id(void* id128)
{
    NcclUniqueId id;
    check(api().GetUniqueId(&id), "ncclGetUniqueId");
    memcpy(id128, &id, sizeof(id));
}

void nccl_comm_init(sb200_comm* c, const void* id128)
{
    NcclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    NcclComm comm = nullptr;
    check(api().CommInitRank(&comm, c->nranks, id, c->rank), "ncclCommInitRank");
    c->nccl = comm;
}

void nccl_comm_destroy(sb200_comm* c)
{
    if (c->nccl)
        api().CommDestroy((NcclComm) c->nccl);
    c->nccl = nullptr;
}

void nccl_allreduce_sum(sb200_comm* c, double* buf, size_t count, cudaStream_t s)
{
    check(api().AllReduce(buf, buf, count, kNcclFloat64, kNcclSum, (NcclComm) c->nccl, s), "ncclAllReduce");
}

void nccl_allreduce_max(sb200_comm* c, double* buf, size_t count, cudaStream_t s)
{
    check(api().AllReduce(buf, buf, count, kNcclFloat64, kNcclMax, (NcclComm) c->nccl, s), "ncclAllReduce");
}

void nccl_allgather(sb200_comm* c, const double* send, double* recv, size_t count_per_rank, cudaStream_t s)
{
    check(api().AllGather(send, recv, count_per_rank, kNcclFloat64, (NcclComm) c->nccl, s), "ncclAllGather");
}

// ---- peer windows (CUDA IPC) ----
// Every rank allocates `bytes` (rounded up to 2 MiB so that the allocation is not shared with unrelated data), exports its IPC handle,
// all-gathers the 64-byte handles through the NCCL communicator and opens the peers' handles.  A final all-reduce agrees on success.
bool peer_window_create(sb200_comm* c, size_t bytes, PeerWindow& w, cudaStream_t s)
{
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    const int P = c->nranks;
    if (P > kMaxPeers)
        return false;
    const size_t gran = (size_t) 2 << 20;
    bytes = (bytes + gran - 1) / gran * gran;
    double failed = 0.0;
    void* local = nullptr;
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof(mine));
    if (cudaMalloc(&local, bytes) != cudaSuccess)
    {
        cudaGetLastError();
        local = nullptr;
        failed = 1.0;
    }
    else if (cudaIpcGetMemHandle(&mine, local) != cudaSuccess)
    {
        cudaGetLastError();
        failed = 1.0;
    }
    // exchange the handles (8 doubles each) and the failure flags through the communicator
    DevBuf<double> dsend(8), drecv((size_t) 8 * P), dflag(1);
    SB200_CUDA_CHECK(cudaMemcpyAsync(dsend.get(), &mine, 64, cudaMemcpyHostToDevice, s));
    nccl_allgather(c, dsend.get(), drecv.get(), 8, s);
    std::vector<cudaIpcMemHandle_t> all((size_t) P);
    SB200_CUDA_CHECK(cudaMemcpyAsync(all.data(), drecv.get(), (size_t) 64 * P, cudaMemcpyDeviceToHost, s));
    SB200_CUDA_CHECK(cudaMemcpyAsync(dflag.get(), &failed, sizeof(double), cudaMemcpyHostToDevice, s));
    nccl_allreduce_sum(c, dflag.get(), 1, s);
    double any_failed = 0.0;
    SB200_CUDA_CHECK(cudaMemcpyAsync(&any_failed, dflag.get(), sizeof(double), cudaMemcpyDeviceToHost, s));
    SB200_CUDA_CHECK(cudaStreamSynchronize(s));
    void* mapped[kMaxPeers] = {};
    failed = any_failed;
    if (failed == 0.0)
        for (int r = 0; r < P; r++)
        {
            if (r == c->rank)
            {
                mapped[r] = local;
                continue;
            }
            if (cudaIpcOpenMemHandle(&mapped[r], all[(size_t) r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess)
            {
                cudaGetLastError();
                mapped[r] = nullptr;
                failed = 1.0;
                break;
            }
        }
    SB200_CUDA_CHECK(cudaMemcpyAsync(dflag.get(), &failed, sizeof(double), cudaMemcpyHostToDevice, s));
    nccl_allreduce_sum(c, dflag.get(), 1, s);
    SB200_CUDA_CHECK(cudaMemcpyAsync(&any_failed, dflag.get(), sizeof(double), cudaMemcpyDeviceToHost, s));
    SB200_CUDA_CHECK(cudaStreamSynchronize(s));
    if (any_failed != 0.0)
    {
        for (int r = 0; r < P; r++)
            if (r != c->rank && mapped[r])
                cudaIpcCloseMemHandle(mapped[r]);
        if (local)
            cudaFree(local);
        cudaGetLastError();
        return false;
    }
    w.local = local;
    w.bytes = bytes;
    w.nranks = P;
    for (int r = 0; r < P; r++)
        w.peer[r] = mapped[r];
    SB200_CUDA_CHECK(cudaMemsetAsync(local, 0, bytes, s));
    // nobody may write into a window before its owner has cleared it
    nccl_allreduce_sum(c, dflag.get(), 1, s);
    SB200_CUDA_CHECK(cudaStreamSynchronize(s));
    return true;
}

void peer_window_destroy(sb200_comm* c, PeerWindow& w)
{
    if (!w.local)
        return;
    for (int r = 0; r < w.nranks; r++)
        if (c && r != c->rank && w.peer[r])
            cudaIpcCloseMemHandle(w.peer[r]);
    cudaFree(w.local);
    w = PeerWindow();
}

}  // namespace sb200

// NCCL binding (row-sharded multi-GPU runs).  libnccl.so.2 is resolved with dlopen
// at first use: a process that already loaded the library (e.g. through torch) shares it, and a
// single-GPU run never touches NCCL.
#include <dlfcn.h>

#include <mutex>

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

}  // namespace sb200

// Host-side objects behind the C ABI handles.
#pragma once

#include <memory>

#include "kernels.h"

struct sb200_comm
{
    int rank = 0, nranks = 1;
    void* nccl = nullptr;  // ncclComm_t
};

namespace sb200 {

// ---- NCCL, loaded lazily with dlopen so that single-GPU use has no NCCL dependency ----
void nccl_unique_id(void* id128);
void nccl_comm_init(sb200_comm* c, const void* id128);
void nccl_comm_destroy(sb200_comm* c);
// in-place sum / max all-reduce of `count` doubles, all-gather of `count` doubles per rank
void nccl_allreduce_sum(sb200_comm* c, double* buf, size_t count, cudaStream_t s);
void nccl_allreduce_max(sb200_comm* c, double* buf, size_t count, cudaStream_t s);
void nccl_allgather(sb200_comm* c, const double* send, double* recv, size_t count_per_rank, cudaStream_t s);

// ---- peer memory over NVLink (row-sharded runs on one node): symmetric device allocations mapped into every rank's address space
// (CUDA IPC handles exchanged through the NCCL communicator), used by the peer kernels of peer.cu -- one-shot all-reduce of the
// dot products and the direct write of each rank's new residual slice into every peer's SpMV operand buffer ----
constexpr int kMaxPeers = 16;
struct PeerWindow
{
    void* local = nullptr;
    size_t bytes = 0;
    void* peer[kMaxPeers] = {};   // peer[r] = rank r's allocation as mapped here; peer[own rank] == local
    int nranks = 0;
    bool ok() const { return local != nullptr; }
};
// Collective over the communicator.  Returns false on every rank (and allocates nothing) when any rank cannot map its peers
// (no IPC / peer access): the caller then stays on the NCCL collectives.
bool peer_window_create(sb200_comm* c, size_t bytes, PeerWindow& w, cudaStream_t s);
void peer_window_destroy(sb200_comm* c, PeerWindow& w);

// ---- banded shift-solve operator (band_solve.cu; SparseSymShiftSolve.h:85-109) ----
struct BandSolve;
BandSolve* band_create(sb200_op* op);
void band_destroy(BandSolve* b);
void band_set_shift(sb200_op* op, double sigma);
void band_solve_device(sb200_op* op, const double* x_dev, double* y_dev);
void band_info(const sb200_op* op, int* half_bandwidth, int* block, int64_t* block_rows, int* levels);
void band_set_refine(sb200_op* op, int steps);
void band_status(const sb200_op* op, int* refine_steps, double* verify_residual, double* unrefined_residual);

}  // namespace sb200

// Device-resident operator: full CSR rows [row0, row0+nrows) of an n x n matrix.
struct sb200_op
{
    sb200::DeviceCsr A;
    sb200::SpmvPlan plan;
    sb200_comm* comm = nullptr;
    cudaStream_t stream = nullptr;
    int64_t slab = 0;                 // rows per rank (ceil(n / nranks)); nrows <= slab
    sb200::DevBuf<double> x_full;     // n_pad = slab * nranks entries (sharded), else n
    sb200::DevBuf<double> y_loc;      // host-pointer perform_op staging
    sb200::DevBuf<double> x_stage;    // slab-sized send buffer for the all-gather
    // chunked all-gather / SpMV overlap (sharded operators, DeviceCsr::chunk_len): operand in chunk-major layout, a second
    // stream for the collectives, one event per chunk
    sb200::DevBuf<double> x_chunks;   // nchunks * nranks * chunk_len (NCCL path / single-GPU test layout)
    double* xc = nullptr;             // the chunk-major operand buffer in use: x_chunks, or the local part of win_x in peer mode
    // peer mode (peer.cu): operand buffer and reduction mailboxes live in symmetric windows every rank can write
    sb200::PeerWindow win_x, win_ctl;
    sb200::DevBuf<unsigned long long> peer_seq;  // round counter of the one-shot all-reduce (identical on every rank)
    bool peer_mode() const { return win_x.ok() && win_ctl.ok(); }
    cudaStream_t comm_stream = nullptr;
    cudaEvent_t ev_ready = nullptr;
    std::vector<cudaEvent_t> ev_chunk;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool symmetric_hint = false;      // created through a SYM mode
    // user-defined host operator (the reference's OpType concept, SymEigsSolver.h:99-114): y = fn(x) on host memory
    void (*cb)(const double*, double*, void*) = nullptr;
    void* cb_user = nullptr;
    sb200::PinnedBuf<double> hx, hy;  // pinned staging of the callback path
    // shift-solve operator: perform_op is y = (A - sigma I)^{-1} x instead of y = A x (A stays available for refinement)
    sb200::BandSolve* band = nullptr;
    // complex Hermitian operator (SparseHermMatProd): the matrix lives in Az; A only carries n / nrows / nnz.
    // Vectors are interleaved complex, i.e. 2 n doubles; single GPU.
    bool cplx = false;
    sb200::DeviceCsrZ Az;
    // true when perform_op is not the fused CSR SpMV (user callback, shift-solve, complex operand): the solvers take the unfused step path
    bool indirect() const { return cb != nullptr || band != nullptr || cplx; }

    int nranks() const { return comm ? comm->nranks : 1; }
    int rank() const { return comm ? comm->rank : 0; }
    ~sb200_op();
};

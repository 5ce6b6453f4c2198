// Peer-memory kernels for row-sharded runs on one NVLink / NVSwitch node.
//
// The reference has no multi-GPU path; what is sharded here is its Lanczos / Arnoldi step (Lanczos.h:102-182, Arnoldi.h:236-290): every
// step needs (1) the new residual f on every rank as the SpMV operand and (2) the sums over ranks of the <= 65 partial dot products of
// each panel pass.  With NCCL that is one all-gather and two to three latency-bound all-reduces per step.  Here every rank maps the
// other ranks' buffers (CUDA IPC, comm.cu) and
//   * the correction pass (panel.cu, PANEL_CORR) stores its rows of the new residual straight into every rank's operand buffer while it
//     streams V -- the exchange rides on the NVLink bandwidth the HBM-bound pass leaves idle, and no all-gather is left;
//   * peer_allreduce_kernel combines the partial sums in one launch: data into every peer's mailbox, a system-scope release of one flag
//     per peer, acquire of the peers' flags, summation in rank order (bitwise identical on all ranks).  Its release / acquire pair is
//     also the barrier that makes the residual rows written by the preceding pass visible before the next SpMV reads them.
// Mailboxes and flags are double-buffered by round parity: a rank can be at most one round ahead of a peer (it needs that peer's flag
// of the current round to finish it), so round r + 2 never overwrites data of round r that is still being read.
#include "peer_dev.cuh"
#ifdef SB200_EMU
#include <sched.h>
#endif

namespace sb200 {

namespace {

__global__ void __launch_bounds__(128) peer_allreduce_kernel(PeerCtl pc, double* buf, int count, int op, const int* abort)
{
    if (abort != nullptr && *abort != 0)
        return;  // sweep mode after an abort: every rank skips the same rounds (the flag derives from reduced, identical values)
    peer_allreduce_body(pc, buf, count, op);
}

__global__ void __launch_bounds__(256) peer_push_kernel(PeerX px, const double* __restrict__ f_loc, int64_t nrows_ld)
{
    const int64_t pairs = (px.rows < nrows_ld ? px.rows : nrows_ld) / 2;
    for (int64_t q = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; q < pairs; q += (int64_t) gridDim.x * blockDim.x)
    {
        const int64_t r0 = 2 * q;
        const double2 v = *reinterpret_cast<const double2*>(f_loc + r0);
        const int64_t c = r0 / px.len;
        const int64_t dst = c * px.stride + (int64_t) px.rank * px.len + (r0 - c * px.len);
        for (int p = 0; p < px.np; p++)
            st_peer_f64x2(px.dst[p] + dst, v);
    }
}

}  // namespace

void launch_peer_allreduce(const PeerCtl& pc, double* buf, int count, int op, cudaStream_t stream, const int* abort)
{
    SB200_REQUIRE(count >= 1 && count <= kRedStride && pc.nranks >= 1 && pc.nranks <= kPeerMax, SB200_LOGIC, "peer all-reduce: bad arguments");
#ifdef SB200_EMU
    // Kernel-logic emulator: launches of different emulated ranks are serialised (process-wide __shared__ statics), so a kernel that
    // spins on a peer's flag can never be answered.  The same mailbox protocol is therefore run by the rank's host thread, outside
    // the launch lock -- it exercises the protocol (parity, flags, rank-order sum) and the drivers' use of it, not the kernel text.
    (void) stream;
    if (abort != nullptr && *abort != 0)
        return;
    const int P = pc.nranks, me = pc.rank;
    const unsigned long long seq = *pc.seq + 1ull;
    const int par = (int) (seq & 1ull);
    for (int t = 0; t < count; t++)
        for (int p = 0; p < P; p++)
            st_sys_f64(pc.slots[p] + ((size_t) (par * P + me)) * kRedStride + t, buf[t]);
    __atomic_thread_fence(__ATOMIC_SEQ_CST);
    for (int p = 0; p < P; p++)
        st_release_sys_u64(pc.flags[p] + (par * P + me), seq);
    for (int q = 0; q < P; q++)
        while (ld_acquire_sys_u64(pc.flags[me] + (par * P + q)) != seq)
            sched_yield();
    for (int t = 0; t < count; t++)
    {
        const double* box = pc.slots[me] + (size_t) par * P * kRedStride + t;
        double a = ld_sys_f64(box);
        for (int q = 1; q < P; q++)
        {
            const double b = ld_sys_f64(box + (size_t) q * kRedStride);
            a = (op == 1) ? std::max(a, b) : a + b;
        }
        buf[t] = a;
    }
    *pc.seq = seq;
#else
    peer_allreduce_kernel<<<1, 128, 0, stream>>>(pc, buf, count, op, abort);
    SB200_CUDA_CHECK(cudaGetLastError());
#endif
}

void launch_peer_push(const PeerX& px, const double* f_loc, int64_t nrows_ld, cudaStream_t stream)
{
    if (px.np <= 0)
        return;
    const int64_t pairs = std::min<int64_t>(px.rows, nrows_ld) / 2;
    const int grid = (int) std::max<int64_t>(1, std::min<int64_t>((pairs + 255) / 256, (int64_t) device_info().sm_count * 4));
    peer_push_kernel<<<grid, 256, 0, stream>>>(px, f_loc, nrows_ld);
    SB200_CUDA_CHECK(cudaGetLastError());
}

}  // namespace sb200

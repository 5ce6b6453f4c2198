// Device body of the one-shot mailbox all-reduce (peer.cu): shared by peer_allreduce_kernel and by the decide kernels that absorb the
// all-reduce of a panel pass (panel.cu), so that a sharded Lanczos step launches one small kernel per panel pass instead of two.
// Block of >= 128 threads, all of which must call it; on return buf[0..count) holds the rank-order sum (or max) on every rank.
#pragma once

#include "kernels.h"

namespace sb200 {

#ifdef __CUDACC__
__device__ __forceinline__ void peer_allreduce_body(const PeerCtl& pc, double* buf, int count, int op)
{
    const int t = threadIdx.x;
    const int P = pc.nranks, me = pc.rank;
    __shared__ unsigned long long s_seq;
    if (t == 0)
        s_seq = *pc.seq + 1ull;
    __syncthreads();
    const unsigned long long seq = s_seq;
    const int par = (int) (seq & 1ull);
    if (t < count)
    {
        const double v = buf[t];
        for (int p = 0; p < P; p++)
            st_sys_f64(pc.slots[p] + ((size_t) (par * P + me)) * kRedStride + t, v);
    }
    __threadfence_system();
    __syncthreads();
    if (t < P)
    {
        st_release_sys_u64(pc.flags[t] + (par * P + me), seq);                  // tell rank t that my contribution of this round has landed
        const unsigned long long* mine = pc.flags[me] + (par * P + t);
        unsigned int spins = 0;
        while (ld_acquire_sys_u64(mine) != seq)                                 // wait for rank t's contribution
        {
#ifndef SB200_EMU
            if (++spins > (1u << 27))
                __trap();  // a peer that never arrives (crashed rank, mismatched call sequence) must not hang the device for good
#endif
        }
        (void) spins;
    }
    __syncthreads();
    if (t < count)
    {
        const double* box = pc.slots[me] + (size_t) par * P * kRedStride + t;
        double a = ld_sys_f64(box);
        for (int q = 1; q < P; q++)
        {
            const double b = ld_sys_f64(box + (size_t) q * kRedStride);
            a = (op == 1) ? fmax(a, b) : a + b;
        }
        buf[t] = a;
    }
    if (t == 0)
        *pc.seq = seq;
    __syncthreads();
}
#endif

}  // namespace sb200

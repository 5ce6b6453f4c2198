// Restart GEMM on the fp64 tensor pipe, fed by TMA (K5/K6 of SURVEY.md §2.1):
//   V[:, :kk] <- V[:, :m] * Q[:, :kk]      Arnoldi::compress_V   (LinAlg/Arnoldi.h:320-340)
//   X = V * S                              eigenvectors()        (HermEigsBase.h:467, GenEigsBase.h:600)
// plus the fused residual update  f <- f*Q(m-1,k-1) + V_new[:,k]*H(k,k-1)  and ||f||^2  (Arnoldi.h:337-339).
//
// Shape: tall-skinny, n x 64 times 64 x <=64, fp64.  tcgen05 has no f64 kind, so the tensor path on
// sm_100a is DMMA: mma.sync.aligned.m8n8k4.row.col.f64.  A CTA of 4 warps owns 64-row tiles of V:
//   * the tile (64 rows x m columns, each column 512 contiguous bytes in the column-major V) is brought in by one
//     1-D bulk TMA copy per column (cp.async.bulk ... mbarrier::complete_tx::bytes), double buffered, so the copy of
//     tile t+2 overlaps the MMAs of tile t+1;
//   * Q (zero padded to 64 x 64) sits in shared memory for the whole kernel;
//   * warp w computes rows 16w..16w+15 against all column tiles with 2 x NT accumulator fragments in registers
//     (16 k-steps x (2 A + NT B fragment loads, 2*NT DMMA));
//   * shared-memory columns are 68 doubles apart, which makes every fragment load bank-conflict free.
// The product is written in place (a tile's rows are read completely before they are written; other tiles touch
// other rows).  Algorithmic bytes: 8 n (m + kk) + 16 n; flops 2 n m kk.
#include "kernels.h"

namespace sb200 {

namespace {

constexpr int kTR = 64;       // tile rows
constexpr int kTC = 64;       // padded panel width / output width
constexpr int kLds = 68;      // shared-memory column stride in doubles (68 mod 16 == 4 -> conflict-free fragments)
constexpr int kStages = 2;
constexpr int kDmmaBlock = 128;

__device__ __forceinline__ void dmma_m8n8k4(double& d0, double& d1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int NT>  // number of 8-column output tiles actually computed (ceil(kk / 8))
__global__ void __launch_bounds__(kDmmaBlock, 2)
    compress_dmma_kernel(const double* V, int64_t ldv, int64_t nrows, int m, const double* __restrict__ Q, int kk, double* Vout, int64_t ldo, double* f,
                         const double* __restrict__ H, double* red_out, double* partials, unsigned int* ticket)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* Qs = reinterpret_cast<double*>(smem_raw);  // Qs[n * kLds + k] = Q(k, n), zero padded
    double* Vs = Qs + kTC * kLds;                      // kStages tiles: Vs[stage][c * kLds + r]
    __shared__ __align__(8) uint64_t full_bar[kStages];
    __shared__ int s_kend[kTC / 8];  // per 8-column output tile: number of k-steps that touch a nonzero of Q
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int grp = lane >> 2, tig = lane & 3;

    for (int idx = tid; idx < kTC * kLds; idx += kDmmaBlock)
    {
        const int n = idx / kLds, k = idx % kLds;
        Qs[idx] = (k < m && n < kk) ? Q[k + (int64_t) n * m] : 0.0;
    }
    for (int idx = tid; idx < kStages * kTC * kLds; idx += kDmmaBlock)
        Vs[idx] = 0.0;  // columns >= m are never written by the TMA and must read as zero
    if (tid == 0)
    {
        for (int s = 0; s < kStages; s++)
            mbar_init(&full_bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // order the generic-proxy initialisation of the tile buffers before the async-proxy (TMA) writes
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();

    // Q of a restart is banded below (column i has m-k+i+1 leading nonzeros, Arnoldi.h:330; the rest are exact zeros that
    // no rotation ever touched), so trailing k-steps of the left column tiles are skipped -- the reference's flop saving.
    if (tid < kTC / 8)
    {
        int last = -1;
        for (int n = tid * 8; n < tid * 8 + 8; n++)
            for (int k = m - 1; k > last; k--)
                if (Qs[n * kLds + k] != 0.0)
                {
                    last = k;
                    break;
                }
        s_kend[tid] = (last + 4) / 4;
    }
    __syncthreads();

    const int64_t ntiles = (nrows + kTR - 1) / kTR;
    int kend[NT], ksteps = 0;
#pragma unroll
    for (int nt = 0; nt < NT; nt++)
    {
        kend[nt] = s_kend[nt];
        ksteps = max(ksteps, kend[nt]);
    }
    double fq = 0.0, fh = 0.0;
    if (f)
    {
        fq = Q[(m - 1) + (int64_t) (kk - 2) * m];   // Q(m-1, k-1)
        fh = H[(kk - 1) + (int64_t) (kk - 2) * m];  // H(k, k-1)
    }
    double nrm = 0.0;

    // warp 0 arms the barrier (lane 0) and launches one bulk copy per panel column (64 rows * 8 B = 512 B each), the columns spread over
    // its lanes so that the 60 copy instructions of a tile issue in two rounds instead of one after the other
    auto issue = [&](int64_t tile, int stage) {
        if (lane == 0)
            mbar_expect_tx(&full_bar[stage], (uint32_t) (m * kTR * sizeof(double)));
        __syncwarp();
        const double* src = V + tile * kTR;
        double* dst = Vs + (size_t) stage * kTC * kLds;
        for (int c = lane; c < m; c += 32)
            tma_load_1d(dst + c * kLds, src + (int64_t) c * ldv, kTR * sizeof(double), &full_bar[stage]);
    };

    int64_t tile = blockIdx.x;
    if (warp == 0)
        for (int s = 0; s < kStages; s++)
            if (tile + (int64_t) s * gridDim.x < ntiles)
                issue(tile + (int64_t) s * gridDim.x, s);

    for (int it = 0; tile < ntiles; it++, tile += gridDim.x)
    {
        const int stage = it % kStages;
        const uint32_t parity = (uint32_t) ((it / kStages) & 1);
        mbar_wait(&full_bar[stage], parity);
        const double* Vt = Vs + (size_t) stage * kTC * kLds;

        double acc[2][NT][2];
#pragma unroll
        for (int mt = 0; mt < 2; mt++)
#pragma unroll
            for (int nt = 0; nt < NT; nt++)
                acc[mt][nt][0] = acc[mt][nt][1] = 0.0;

        for (int ks = 0; ks < ksteps; ks++)
        {
            const int kc = ks * 4 + tig;  // k index of this lane's A / B elements
            const double a0 = Vt[kc * kLds + warp * 16 + grp];
            const double a1 = Vt[kc * kLds + warp * 16 + 8 + grp];
#pragma unroll
            for (int nt = 0; nt < NT; nt++)
            {
                if (ks < kend[nt])  // warp-uniform
                {
                    const double b = Qs[(nt * 8 + grp) * kLds + kc];
                    dmma_m8n8k4(acc[0][nt][0], acc[0][nt][1], a0, b);
                    dmma_m8n8k4(acc[1][nt][0], acc[1][nt][1], a1, b);
                }
            }
        }
        // every warp has consumed its fragments of this stage: the buffer may be refilled
        __syncthreads();
        if (warp == 0 && tile + (int64_t) kStages * gridDim.x < ntiles)
            issue(tile + (int64_t) kStages * gridDim.x, stage);

        // C fragment: row = grp, cols = 2*tig + {0,1} of each 8 x 8 tile
        const int64_t r0 = tile * kTR + warp * 16;
#pragma unroll
        for (int mt = 0; mt < 2; mt++)
        {
            const int64_t row = r0 + mt * 8 + grp;
            if (row < nrows)
            {
#pragma unroll
                for (int nt = 0; nt < NT; nt++)
                {
                    const int c0 = nt * 8 + tig * 2;
                    if (c0 < kk)
                        Vout[row + (int64_t) c0 * ldo] = acc[mt][nt][0];
                    if (c0 + 1 < kk)
                        Vout[row + (int64_t) (c0 + 1) * ldo] = acc[mt][nt][1];
                    if (f)
                    {
                        // the lane holding column kk-1 of this row updates the residual (Arnoldi.h:337)
                        if (c0 == kk - 1 || c0 + 1 == kk - 1)
                        {
                            const double vk = (c0 == kk - 1) ? acc[mt][nt][0] : acc[mt][nt][1];
                            const double fn = f[row] * fq + vk * fh;
                            f[row] = fn;
                            nrm = fma(fn, fn, nrm);
                        }
                    }
                }
            }
        }
    }

    if (f)
    {
        __shared__ double s_w[kDmmaBlock / 32];
        nrm = warp_sum(nrm);
        if (lane == 0)
            s_w[warp] = nrm;
        __syncthreads();
        double cta = 0.0;
        if (tid == 0)
            for (int q = 0; q < kDmmaBlock / 32; q++)
                cta += s_w[q];
        grid_reduce_fixed_order<kDmmaBlock>(cta, 1, partials, ticket, red_out);
    }
}

template <int NT>
void launch_nt(int grid, size_t smem, cudaStream_t stream, const double* V, int64_t ldv, int64_t nrows, int m, const double* Q, int kk, double* Vout, int64_t ldo,
               double* f, const double* H, double* red_out, const RedScratch& rs)
{
    SB200_CUDA_CHECK(cudaFuncSetAttribute((const void*) compress_dmma_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    compress_dmma_kernel<NT><<<grid, kDmmaBlock, smem, stream>>>(V, ldv, nrows, m, Q, kk, Vout, ldo, f, H, red_out, rs.partials, rs.ticket);
}

}  // namespace

// Requirements: ldv, ldo multiples of 64 rows (tile loads read whole 64-row tiles; padding rows are zero), m <= 64.
void launch_compress_dmma(const double* V, int64_t ldv, int64_t nrows, int m, const double* Q, int kk, double* Vout, int64_t ldo, double* f, const double* H,
                          double* red_out, const RedScratch& rs, cudaStream_t stream)
{
    SB200_REQUIRE(m >= 1 && m <= kTC && kk >= 1 && kk <= kTC, SB200_INVALID_ARGUMENT, "compress_dmma: bad dimensions");
    SB200_REQUIRE(ldv % kTR == 0 && ldo % kTR == 0, SB200_LOGIC, "compress_dmma: leading dimensions must be multiples of 64");
    const int sms = device_info().sm_count;
    const int64_t ntiles = (nrows + kTR - 1) / kTR;
    const int grid = (int) std::max<int64_t>(1, std::min<int64_t>(ntiles, (int64_t) sms * 2));
    SB200_REQUIRE(grid <= rs.max_grid, SB200_LOGIC, "compress_dmma: reduction scratch too small");
    const size_t smem = sizeof(double) * (size_t) (kTC * kLds * (1 + kStages));
    const int nt = (kk + 7) / 8;
    switch (nt)
    {
        case 1: launch_nt<1>(grid, smem, stream, V, ldv, nrows, m, Q, kk, Vout, ldo, f, H, red_out, rs); break;
        case 2: launch_nt<2>(grid, smem, stream, V, ldv, nrows, m, Q, kk, Vout, ldo, f, H, red_out, rs); break;
        case 3: launch_nt<3>(grid, smem, stream, V, ldv, nrows, m, Q, kk, Vout, ldo, f, H, red_out, rs); break;
        case 4: launch_nt<4>(grid, smem, stream, V, ldv, nrows, m, Q, kk, Vout, ldo, f, H, red_out, rs); break;
        case 5: launch_nt<5>(grid, smem, stream, V, ldv, nrows, m, Q, kk, Vout, ldo, f, H, red_out, rs); break;
        case 6: launch_nt<6>(grid, smem, stream, V, ldv, nrows, m, Q, kk, Vout, ldo, f, H, red_out, rs); break;
        case 7: launch_nt<7>(grid, smem, stream, V, ldv, nrows, m, Q, kk, Vout, ldo, f, H, red_out, rs); break;
        default: launch_nt<8>(grid, smem, stream, V, ldv, nrows, m, Q, kk, Vout, ldo, f, H, red_out, rs); break;
    }
    SB200_CUDA_CHECK(cudaGetLastError());
}

}  // namespace sb200

// Device-resident Krylov factorisation state shared by the Lanczos (solver_sym.cu) and Arnoldi
// (solver_gen.cu) drivers: A V = V H + f e_k'.  Mirrors the members and the non-virtual parts of
// LinAlg/Arnoldi.h (state :47-61, expand_basis :66-115, init :136-195, compress_V :320-340).
#pragma once

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <numeric>

#include "host.h"

namespace sb200 {

// Util/SimpleRandom.h:30-123 (default residual and expand_basis vectors); host side, O(n) integers.
struct SimpleRandom
{
    long m_rand;
    explicit SimpleRandom(unsigned long init_seed)
    {
        const unsigned long m_max = 2147483647L;
        m_rand = init_seed ? (long) (init_seed & m_max) : 1;
    }
    static long next_long_rand(long seed)
    {
        const unsigned int m_a = 16807;
        const unsigned long m_max = 2147483647L;
        unsigned long lo, hi;
        lo = (unsigned long) m_a * (unsigned long) (seed & 0xFFFFUL);
        hi = (unsigned long) m_a * (unsigned long) ((unsigned long) seed >> 16);
        lo += (hi & 0x7FFF) << 16;
        if (lo > m_max)
        {
            lo &= m_max;
            ++lo;
        }
        lo += hi >> 15;
        if (lo > m_max)
        {
            lo &= m_max;
            ++lo;
        }
        return (long) lo;
    }
    double random()
    {
        m_rand = next_long_rand(m_rand);
        return double(m_rand) / double(2147483647L) - 0.5;
    }
    void random_vec(double* v, int64_t len)
    {
        for (int64_t i = 0; i < len; i++)
            v[i] = random();
    }
};

void launch_trim_h(double* H, int m, int from_k, cudaStream_t s);
void op_spmv_device(sb200_op* op, const double* x_dev, double* y_dev);

struct FacBase
{
    sb200_op* op = nullptr;
    // n / nloc / ld count DOUBLES: for a complex (Hermitian) operator every vector holds interleaved (re, im) pairs and is handled
    // as a real vector of twice the length by every kernel except the operator itself and the panel passes (cw = 2).
    int64_t n = 0, nloc = 0, ld = 0;
    int cw = 1;
    bool is_cplx() const { return cw == 2; }
    int nev = 0, m = 0;

    DevBuf<double> V, f, w, t0, H, Q, S, X;
    // w of the step being built: `w`, or -- when a sweep overlaps the tail of a correction pass with the first operator kernels of the
    // next step (solver_sym.cu, factorize_from) -- alternately `w` and `w_alt`
    DevBuf<double> w_alt;
    double* wp = nullptr;
    cudaStream_t aux_stream = nullptr;                 // carries the speculatively launched operator kernels of the next step
    cudaEvent_t ev_part_a = nullptr, ev_k0 = nullptr;  // first part of the residual final / those kernels done
    DevBuf<double> Hi;        // complex Arnoldi only: imaginary parts of the Hessenberg matrix (H holds the real parts)
    bool complex_h = false;   // keep the imaginary parts of the projected matrix (general complex operator; Hermitian ones drop them)
    DevBuf<FacCtl> ctl;
    DevBuf<double> partials;
    DevBuf<unsigned int> ticket;
    RedScratch rs;
    PinnedBuf<char> hstat;   // status readback
    PinnedBuf<double> hred;  // reduction readback (rare paths)

    int64_t k = 0;  // current subspace dimension (m_k)
    int64_t nmatop = 0, niter = 0;
    int info = SB200_NOT_COMPUTED;
    double h_beta = 0.0;
    bool initialised = false;

    // peer mode (row-sharded runs with NVLink-mapped windows, peer.cu): mailboxes of the one-shot all-reduce, destinations of the residual
    // rows in every rank's operand buffer, and whether those buffers currently hold f
    PeerCtl pctl{};
    PeerX px{};
    bool peer = false;
    bool x_published = false;
    // sweep mode: the steps of one factorize_from() are enqueued back to back without reading the status in between; kernels carry the
    // device flag FacCtl::abort that hands control back when a step needs the host (see lanczos_decide_kernel)
    bool in_sweep = false;
    const int* abort_flag() { return in_sweep ? &ctl.get()->abort : nullptr; }
    bool sweep_capable() const { return !op->indirect() && !is_cplx() && (op->nranks() == 1 || peer) && !sweep_disabled(); }
    static bool sweep_disabled()
    {
        const char* e = std::getenv("SB200_SWEEP");  // read per factorisation: an A/B and test knob
        return e && e[0] == '0';
    }
    void clear_abort() { SB200_CUDA_CHECK(cudaMemsetAsync(&ctl.get()->abort, 0, sizeof(int), stream())); }
    // Overlapped sweeps (opt-in, SB200_OVERLAP=1; >= 2 column blocks; one GPU in the natural layout, or peer mode): the correction pass
    // runs in two row ranges, and as soon as the rows that the first column blocks of the NEXT operator application gather from are
    // final, those (gather-bound) kernels start on a second stream next to the (HBM-bound) second range.  Measured SLOWER than the plain
    // sweep (438 vs 487 SpMV-iters/s at n = 1e7, profiles/r2g_quick_*overlap_n1e7.log): the two kernels do not share SMs (registers), so
    // each runs on a subset of the SMs, and the gather kernel's bound is per SM (L1TEX wavefronts) -- unlike the fused operator kernel,
    // whose two phases alternate inside every SM.  Kept as a tested code path, off by default.
    int num_blocks() const { return op->A.blocks.empty() ? 1 : (int) op->A.blocks.size(); }
    bool overlap_capable() const
    {
        const char* e = std::getenv("SB200_OVERLAP");  // read per sweep: a test knob, not a hot path
        const bool on = e && e[0] == '1';
        if (!on || !sweep_capable() || num_blocks() < 2)
            return false;
        return peer || (op->nranks() == 1 && op->A.chunk_len == 0);
    }
    // local rows [0, h) are what column blocks 0 .. nb-2 gather from (this rank's share of them in the chunk-major layout)
    int64_t overlap_split_rows() const
    {
        const int64_t per = op->A.chunk_len ? op->A.chunk_len : op->A.col_block_width;
        const int64_t h = std::min<int64_t>(nloc, per * (num_blocks() - 1));
        return h & ~int64_t(1);
    }
    void ensure_overlap_resources()
    {
        if (!aux_stream)
        {
            SB200_CUDA_CHECK(cudaStreamCreateWithFlags(&aux_stream, cudaStreamNonBlocking));
            SB200_CUDA_CHECK(cudaEventCreateWithFlags(&ev_part_a, cudaEventDisableTiming));
            SB200_CUDA_CHECK(cudaEventCreateWithFlags(&ev_k0, cudaEventDisableTiming));
        }
        if (w_alt.n < (size_t) ld)
        {
            w_alt.alloc((size_t) ld);
            w_alt.zero(stream());
        }
    }
    // column blocks 0 .. nb-2 of the operator application of step i (raw products accumulated into wdst) on stream st
    void launch_head_blocks(int i, cudaStream_t st, double* wdst)
    {
        const DeviceCsr& A = op->A;
        const int nb = num_blocks();
        ScopedKernelTimer t(&prof, st, KC_SPMV, nb - 1);
        for (int c = 0; c + 1 < nb; c++)
        {
            const double* xb = A.chunk_len ? op->xc + (int64_t) c * A.chunk_stride() : f.get();
            launch_spmv_step_block(A, op->plan, c, xb, f.get(), V.get(), ld, wdst, ctl.get(), H.get(), m, i, 0, true, rs, st, nullptr);
        }
    }
    // last column block of step i on the main stream: step head + (sliced layout) c = V^T w; returns whether the panel pass was fused
    bool launch_last_block(int i, bool symmetric, double* wbuf)
    {
        const DeviceCsr& A = op->A;
        const int nb = num_blocks();
        const double* xb = A.chunk_len ? op->xc + (int64_t) (nb - 1) * A.chunk_stride() : f.get();
        bool fused;
        {
            ScopedKernelTimer t(&prof, stream(), KC_SPMV);
            fused = launch_spmv_step_block(A, op->plan, nb - 1, xb, f.get(), V.get(), ld, wbuf, ctl.get(), H.get(), m, i, 0, symmetric, rs, stream(), ctl.get()->red);
        }
        stats.spmv_launches++;
        nmatop++;
        return count_fused(fused, i);
    }

    Profiler prof;
    sb200_stats stats;
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr;

    cudaStream_t stream() const { return op->stream; }
    int P() const { return op->nranks(); }

    ~FacBase()
    {
        if (ev_begin)
            cudaEventDestroy(ev_begin);
        if (ev_end)
            cudaEventDestroy(ev_end);
        if (ev_part_a)
            cudaEventDestroy(ev_part_a);
        if (ev_k0)
            cudaEventDestroy(ev_k0);
        if (aux_stream)
            cudaStreamDestroy(aux_stream);
    }

    void alloc_common(sb200_op* op_, int64_t nev_, int64_t m_)
    {
        op = op_;
        cw = op->cplx ? 2 : 1;
        SB200_REQUIRE(cw == 1 || op->nranks() == 1, SB200_INVALID_ARGUMENT, "complex operators are single-GPU in this build");
        n = op->A.n * cw;
        nloc = op->A.nrows * cw;
        // leading dimension: multiple of 16 doubles (128 B) and at least the all-gather slab
        // (64-row tiles of the restart GEMM; chunked all-gathers read nchunks * chunk_len local entries)
        const int64_t chunk_rows = (op->A.chunk_len && op->nranks() > 1) ? op->A.chunk_len * (int64_t) op->A.blocks.size() : 0;
        ld = round_up(std::max<int64_t>(std::max<int64_t>(std::max<int64_t>(nloc, op->slab * cw), chunk_rows), 2), 64);
        nev = (int) nev_;
        m = (int) m_;
        V.alloc((size_t) ld * m);
        f.alloc((size_t) ld);
        w.alloc((size_t) ld);
        wp = w.get();
        t0.alloc((size_t) std::max<int64_t>(ld, n));
        H.alloc((size_t) m * m);
        if (complex_h)
            Hi.alloc((size_t) m * m);
        Q.alloc((size_t) m * m);
        S.alloc((size_t) 2 * m * m);
        ctl.alloc(1);
        const int max_grid = reduction_max_grid(device_info().sm_count);  // 16 MB of partials
        partials.alloc((size_t) max_grid * kRedStride);
        ticket.alloc(1);
        ticket.zero(op->stream);
        rs.partials = partials.get();
        rs.ticket = ticket.get();
        rs.max_grid = max_grid;
        hstat.alloc(256);
        hred.alloc(kRedStride + 8);
        peer = op->peer_mode();
        if (peer)
        {
            const int Pn = op->nranks();
            for (int r = 0; r < Pn; r++)
            {
                pctl.slots[r] = static_cast<double*>(op->win_ctl.peer[r]);
                pctl.flags[r] = reinterpret_cast<unsigned long long*>(static_cast<double*>(op->win_ctl.peer[r]) + (size_t) 2 * Pn * kRedStride);
                px.dst[r] = static_cast<double*>(op->win_x.peer[r]);
            }
            pctl.seq = op->peer_seq.get();
            pctl.rank = op->rank();
            pctl.nranks = Pn;
            px.np = Pn;
            px.rank = op->rank();
            px.len = op->A.chunk_len;
            px.stride = op->A.chunk_stride();
            px.rows = op->A.chunk_len * (int64_t) op->A.blocks.size();
        }
        SB200_CUDA_CHECK(cudaEventCreate(&ev_begin));
        SB200_CUDA_CHECK(cudaEventCreate(&ev_end));
        std::memset(&stats, 0, sizeof(stats));
    }

    // ---- small helpers -------------------------------------------------------------------------
    const FacCtl* read_status()
    {
        SB200_CUDA_CHECK(cudaMemcpyAsync(hstat.get(), ctl.get(), kFacCtlStatusBytes, cudaMemcpyDeviceToHost, stream()));
        SB200_CUDA_CHECK(cudaStreamSynchronize(stream()));
        stats.host_syncs++;
        return reinterpret_cast<const FacCtl*>(hstat.get());
    }
    void allreduce_sum(double* buf, size_t count)
    {
        if (P() > 1)
        {
            ScopedKernelTimer t(&prof, stream(), KC_COMM, peer ? 1 : 0);
            if (peer)
                launch_peer_allreduce(pctl, buf, (int) count, 0, stream(), abort_flag());
            else
                nccl_allreduce_sum(op->comm, buf, count, stream());
        }
    }
    void allreduce_max(double* buf, size_t count)
    {
        if (P() > 1)
        {
            ScopedKernelTimer t(&prof, stream(), KC_COMM, peer ? 1 : 0);
            if (peer)
                launch_peer_allreduce(pctl, buf, (int) count, 1, stream());
            else
                nccl_allreduce_max(op->comm, buf, count, stream());
        }
    }
    // peer mode: make every rank's operand buffer hold the current residual f (cold paths -- after init, a restart, expand_basis or a
    // zeroed residual; on the hot path the correction pass has already written it).  The all-reduce is the barrier for the peer writes.
    void publish_f()
    {
        ScopedKernelTimer t(&prof, stream(), KC_COMM, 2);
        launch_peer_push(px, f.get(), ld, stream());
        launch_peer_allreduce(pctl, ctl.get()->red_a + 3, 1, 0, stream());
        x_published = true;
    }
    // x_full <- all-gather of a local vector (sharded); returns the pointer the SpMV must read
    const double* gather_full(const double* local)
    {
        if (P() == 1)
            return local;
        ScopedKernelTimer t(&prof, stream(), KC_COMM, 0);
        // local vectors have ld >= slab entries allocated; padding rows are zero
        nccl_allgather(op->comm, local, op->x_full.get(), (size_t) op->slab, stream());
        return op->x_full.get();
    }
    double reduce_scalar(int opk, const double* x, const double* y)
    {
        double* slot = ctl.get()->red_a + 1;
        {
            ScopedKernelTimer t(&prof, stream(), KC_PANEL);
            launch_vec_reduce(opk, x, y, nloc, slot, rs, stream());
        }
        if (opk == VR_MAXABS || opk == VR_CMAXABS)
            allreduce_max(slot, 1);
        else
            allreduce_sum(slot, 1);
        SB200_CUDA_CHECK(cudaMemcpyAsync(hred.get(), slot, sizeof(double), cudaMemcpyDeviceToHost, stream()));
        SB200_CUDA_CHECK(cudaStreamSynchronize(stream()));
        stats.host_syncs++;
        return hred.get()[0];
    }
    void set_beta_host(double b)
    {
        h_beta = b;
        launch_set_scalar(&ctl.get()->beta, b, stream());
        prof.launches++;
    }
    void spmv_plain(const double* xfull, double* y)
    {
        {
            ScopedKernelTimer t(&prof, stream(), KC_SPMV);
            op_spmv_device(op, xfull, y);
        }
        stats.spmv_launches++;
        nmatop++;
    }
    // K-A of one step.  want_dot: ask the operator kernel to also deliver ctl->red[0..i] = V[:, :i+1]^T w (the first panel pass);
    // returns true when it did (sliced layout), false when the caller has to run the panel pass itself.
    bool spmv_step(int i, bool restarted, bool symmetric, bool want_dot = false)
    {
        double* dot_out = want_dot ? ctl.get()->red : nullptr;
        if (op->indirect())
        {
            // user-defined host operator / device shift-solve: v_i = f/beta on the device, w = op(v_i), epilogue on the device
            ScopedKernelTimer t(&prof, stream(), KC_SPMV, 2);
            double* vi = V.get() + (int64_t) i * ld;
            launch_step_scale(f.get(), ctl.get(), vi, nloc, stream());
            op_spmv_device(op, vi, wp);
            launch_step_epilogue(wp, V.get(), ld, nloc, ctl.get(), H.get(), m, i, restarted ? 1 : 0, symmetric, rs, stream());
            stats.spmv_launches++;
            nmatop++;
            return false;
        }
        if (op->A.chunk_len && P() > 1)
        {
            const bool fused = spmv_step_chunked(i, restarted, symmetric, dot_out);
            stats.spmv_launches++;
            nmatop++;
            return count_fused(fused, i);
        }
        if (op->A.chunk_len)
        {
            // single-GPU test layout (SB200_FORCE_CHUNK_RANKS): permute instead of gathering
            ScopedKernelTimer t(&prof, stream(), KC_SPMV, 1 + (int) op->A.blocks.size());
            launch_permute_to_chunks(op->A, f.get(), op->xc, stream());
            const bool fused = launch_spmv_step(op->A, op->plan, op->xc, f.get(), V.get(), ld, wp, ctl.get(), H.get(), m, i, restarted ? 1 : 0,
                                                symmetric, rs, stream(), dot_out);
            stats.spmv_launches++;
            nmatop++;
            return count_fused(fused, i);
        }
        const double* xfull = gather_full(f.get());
        bool fused;
        {
            ScopedKernelTimer t(&prof, stream(), KC_SPMV);
            fused = launch_spmv_step(op->A, op->plan, xfull, f.get(), V.get(), ld, wp, ctl.get(), H.get(), m, i, restarted ? 1 : 0, symmetric, rs, stream(),
                                     dot_out);
        }
        stats.spmv_launches++;
        nmatop++;
        return count_fused(fused, i);
    }
    bool count_fused(bool fused, int i)
    {
        if (fused)
        {
            stats.fused_dot_launches++;
            stats.fused_dot_cols += i;
        }
        return fused;
    }
    // first panel pass of a step: red[0..j) = V[:, :j]^T w, either already delivered by the operator kernel (fused) or run here
    void step_dot(int i, bool restarted, bool symmetric)
    {
        const int j = i + 1;
        if (spmv_step(i, restarted, symmetric, !is_cplx()))
            allreduce_sum(ctl.get()->red, (size_t) j);
        else
            panel(PANEL_DOT, j, wp, nullptr, nullptr);
    }
    // Sharded operator: the operand is all-gathered in chunks on the communication stream; the SpMV of column block c (the
    // columns that chunk c delivers) starts as soon as chunk c has landed, while chunk c+1 is still on the wire.
    bool spmv_step_chunked(int i, bool restarted, bool symmetric, double* dot_out)
    {
        const DeviceCsr& A = op->A;
        const int nb = (int) A.blocks.size();
        const int64_t len = A.chunk_len, stride = A.chunk_stride();
        if (peer)
        {
            // the operand is already in place: the last correction pass (or publish_f) wrote every rank's rows into all operand buffers
            if (!x_published)
                publish_f();
            bool fused_p = false;
            ScopedKernelTimer t(&prof, stream(), KC_SPMV, nb);
            for (int c = 0; c < nb; c++)
                fused_p = launch_spmv_step_block(A, op->plan, c, op->xc + (int64_t) c * stride, f.get(), V.get(), ld, wp, ctl.get(), H.get(), m, i,
                                                 restarted ? 1 : 0, symmetric, rs, stream(), dot_out);
            x_published = false;  // w / f move on; the next correction pass republishes
            return fused_p;
        }
        SB200_CUDA_CHECK(cudaEventRecord(op->ev_ready, stream()));  // f is final on the compute stream
        SB200_CUDA_CHECK(cudaStreamWaitEvent(op->comm_stream, op->ev_ready, 0));
        for (int c = 0; c < nb; c++)
        {
            nccl_allgather(op->comm, f.get() + (int64_t) c * len, op->xc + (int64_t) c * stride, (size_t) len, op->comm_stream);
            SB200_CUDA_CHECK(cudaEventRecord(op->ev_chunk[(size_t) c], op->comm_stream));
        }
        prof.launches += nb;
        bool fused = false;
        for (int c = 0; c < nb; c++)
        {
            {
                // the wait for the chunk is what remains visible of the collective
                ScopedKernelTimer t(&prof, stream(), KC_COMM, 0);
                SB200_CUDA_CHECK(cudaStreamWaitEvent(stream(), op->ev_chunk[(size_t) c], 0));
            }
            ScopedKernelTimer t(&prof, stream(), KC_SPMV);
            fused = launch_spmv_step_block(A, op->plan, c, op->xc + (int64_t) c * stride, f.get(), V.get(), ld, wp, ctl.get(), H.get(), m, i,
                                           restarted ? 1 : 0, symmetric, rs, stream(), dot_out);
        }
        return fused;
    }
    // reduce = false: the caller combines the partial results over the ranks itself (decide_after: all-reduce fused into the decide kernel)
    void panel(int mode, int j, const double* x, double* fo, const double* coef, const int* pred = nullptr, bool reduce = true)
    {
        stats.panel_launches++;
        stats.panel_cols += j;
        const bool push = peer && mode == PANEL_CORR && fo == f.get();
        {
            ScopedKernelTimer t(&prof, stream(), KC_PANEL);
            launch_panel_pass(mode, V.get(), ld, nloc, j, x, fo, coef, ctl.get()->red, rs, stream(), pred, is_cplx(), push ? &px : nullptr, abort_flag());
        }
        if (reduce)
            allreduce_sum(ctl.get()->red, kRedNrm + 1);
        if (push && pred == nullptr)
            x_published = true;  // the all-reduce above is the barrier for the rows just written into the peers' operand buffers
    }
    // One row range [r0, r1) of the correction pass f = x - V c (real path): its partial V^T f and ||f||^2 go to red_out (all-reduced);
    // r1 == nloc covers the zero padding rows up to ld as the full pass does.
    void panel_corr_range(int j, const double* x, const double* coef, int64_t r0, int64_t r1, double* red_out)
    {
        const int64_t limit = (r1 >= nloc ? ld : r1) - r0;
        PeerX pr = px;
        pr.row0 = r0;
        {
            ScopedKernelTimer t(&prof, stream(), KC_PANEL);
            launch_panel_pass(PANEL_CORR, V.get() + r0, ld, std::max<int64_t>(r1 - r0, 0), j, x + r0, f.get() + r0, coef, red_out, rs, stream(), nullptr, false,
                              peer ? &pr : nullptr, abort_flag(), limit);
        }
        allreduce_sum(red_out, kRedNrm + 1);
    }
    // Lanczos (real path): combine ctl->red[0..count) over the ranks, then the decisions of `stage` -- one kernel in peer mode
    // (lanczos_decide_peer_kernel), all-reduce + decide kernel otherwise.
    void decide_after(int stage, int count, double beta_thresh, int sweep)
    {
        if (peer)
        {
            ScopedKernelTimer t(&prof, stream(), KC_COMM, 1);
            launch_lanczos_decide_peer(pctl, count, ctl.get(), H.get(), m, beta_thresh, stage, stream(), sweep, 0);
        }
        else
        {
            allreduce_sum(ctl.get()->red, (size_t) count);
            launch_lanczos_decide(ctl.get(), H.get(), m, beta_thresh, stage, stream(), 0, false, sweep);
            prof.launches++;
        }
    }
    // A speculatively enqueued pass turned out to be skipped on the device: undo its accounting.
    void uncount_panel(int j)
    {
        stats.panel_launches--;
        stats.panel_cols -= j;
    }
    // host copy of red[0..j) and red[kRedNrm]
    void fetch_red(int j, double& ortho_err, double& nrm2)
    {
        SB200_CUDA_CHECK(cudaMemcpyAsync(hred.get(), ctl.get()->red, sizeof(double) * (is_cplx() ? kRedStride : kRedNrm + 1), cudaMemcpyDeviceToHost, stream()));
        SB200_CUDA_CHECK(cudaStreamSynchronize(stream()));
        stats.host_syncs++;
        ortho_err = 0.0;
        for (int q = 0; q < j; q++)
            ortho_err = std::max(ortho_err, is_cplx() ? std::hypot(hred.get()[q], hred.get()[kRedNrm + 1 + q]) : std::fabs(hred.get()[q]));
        nrm2 = hred.get()[kRedNrm];
    }

    // coefficients of the next correction pass <- result of the last panel reduction (complex: Re and Im halves, see panel.cu)
    void copy_red_to_c(int i)
    {
        SB200_CUDA_CHECK(cudaMemcpyAsync(ctl.get()->c, ctl.get()->red, sizeof(double) * (is_cplx() ? kRedStride : i), cudaMemcpyDeviceToDevice, stream()));
    }

    // ---- Arnoldi::expand_basis (Arnoldi.h:66-115), V = first i columns ----
    void expand_basis(int i, int64_t seed)
    {
        stats.expand_calls++;
        std::vector<double> rnd((size_t) n);
        const int64_t row0 = op->A.row0;
        for (int iter = 0; iter < 5; iter++)
        {
            SimpleRandom rng((unsigned long) (seed + 123 * iter));
            rng.random_vec(rnd.data(), n);
            if (iter == 0)
            {
                // the first try forces f into the range of A: f = A * rand
                double* xfull = (P() > 1) ? op->x_full.get() : t0.get();
                SB200_CUDA_CHECK(cudaMemcpyAsync(xfull, rnd.data(), sizeof(double) * n, cudaMemcpyHostToDevice, stream()));
                spmv_plain(xfull, f.get());
            }
            else if (nloc > 0)
            {
                SB200_CUDA_CHECK(cudaMemcpyAsync(f.get(), rnd.data() + row0, sizeof(double) * nloc, cudaMemcpyHostToDevice, stream()));
            }
            // Vf = V^T f ; f -= V Vf ; fnorm ; Vf = V^T f   (:88-95) — the last three in one fused pass
            panel(PANEL_DOT, i, f.get(), nullptr, nullptr);
            copy_red_to_c(i);
            panel(PANEL_CORR, i, f.get(), f.get(), ctl.get()->c);
            double ortho_err, nrm2;
            fetch_red(i, ortho_err, nrm2);
            double fnorm = std::sqrt(nrm2);
            int count = 0;
            while (count < 3 && ortho_err >= kEps * fnorm)
            {
                copy_red_to_c(i);
                panel(PANEL_CORR, i, f.get(), f.get(), ctl.get()->c);
                fetch_red(i, ortho_err, nrm2);
                fnorm = std::sqrt(nrm2);
                count++;
            }
            set_beta_host(fnorm);
            if (ortho_err < kEps * fnorm)
                return;
        }
    }

    // ---- Arnoldi::init (Arnoldi.h:136-195) ----
    void init_factorization(const double* init_resid)
    {
        std::vector<double> gen;
        if (!init_resid)
        {
            // HermEigsBase.h:337-342 / GenEigsBase.h:470-475: SimpleRandom<Scalar> rng(0); random_vec(n)
            gen.resize((size_t) n);
            SimpleRandom rng(0);
            rng.random_vec(gen.data(), n);
            init_resid = gen.data();
        }
        prof.reset();
        std::memset(&stats, 0, sizeof(stats));
        SB200_CUDA_CHECK(cudaEventRecord(ev_begin, stream()));
        V.zero(stream());
        f.zero(stream());
        w.zero(stream());
        t0.zero(stream());
        H.zero(stream());
        if (complex_h)
            Hi.zero(stream());
        ctl.zero(stream());
        nmatop = 0;
        niter = 0;
        k = 0;
        info = SB200_NOT_COMPUTED;

        const int64_t row0 = op->A.row0;
        // full v0 on the device for the first product, local slice in t0
        double* xfull = (P() > 1) ? op->x_full.get() : t0.get();
        SB200_CUDA_CHECK(cudaMemcpyAsync(xfull, init_resid, sizeof(double) * n, cudaMemcpyHostToDevice, stream()));
        if (P() > 1 && nloc > 0)
            SB200_CUDA_CHECK(cudaMemcpyAsync(t0.get(), init_resid + row0, sizeof(double) * nloc, cudaMemcpyHostToDevice, stream()));
        const double v0norm = std::sqrt(reduce_scalar(VR_SUMSQ, t0.get(), nullptr));
        if (v0norm < kNear0)
            throw Error(SB200_INVALID_ARGUMENT, "initial residual vector cannot be zero");
        double* v = V.get();
        spmv_plain(xfull, v);  // v = A * v0 (force v into the range of A)
        const double vnorm = std::sqrt(reduce_scalar(VR_SUMSQ, v, nullptr));
        if (vnorm < kNear0)
            launch_vec_scale(t0.get(), v0norm, 1, v, nloc, stream());  // v = v0 / ||v0||   (:162-165)
        else
            launch_vec_scale(v, vnorm, 1, v, nloc, stream());  // v /= ||v||
        prof.launches++;
        const double* vfull = gather_full(v);
        spmv_plain(vfull, wp);  // w = A * v
        const double h00 = reduce_scalar(VR_DOT, v, wp);
        launch_set_scalar(H.get(), h00, stream());
        double h00_im = 0.0;
        if (complex_h)
        {
            // H(0,0) = v^H w is genuinely complex for a general complex operator (Arnoldi.h:176-177)
            h00_im = reduce_scalar(VR_CDOT_IM, v, wp);
            launch_set_scalar(Hi.get(), h00_im, stream());
            launch_vec_caxpy(wp, v, h00, h00_im, f.get(), nloc, stream());
            prof.launches++;
        }
        else
            launch_vec_axpy(wp, v, h00, f.get(), nloc, stream());  // f = w - v * H(0,0)
        prof.launches += 2;
        const double fmax = reduce_scalar(is_cplx() ? VR_CMAXABS : VR_MAXABS, f.get(), nullptr);  // m_fac_f.cwiseAbs().maxCoeff()
        if (fmax < kEps * std::hypot(h00, h00_im))
        {
            f.zero(stream());
            set_beta_host(0.0);
        }
        else
        {
            set_beta_host(std::sqrt(reduce_scalar(VR_SUMSQ, f.get(), nullptr)));
        }
        k = 1;
        initialised = true;
        x_published = false;
    }

    // ---- Arnoldi::compress_V (Arnoldi.h:320-340) with the new subspace size knew ----
    void compress_v(int knew)
    {
        stats.compress_launches++;
        stats.compress_cols += knew + 1;
        {
            ScopedKernelTimer t(&prof, stream(), KC_COMPRESS);
            launch_compress(V.get(), ld, nloc, m, Q.get(), knew + 1, V.get(), ld, f.get(), H.get(), ctl.get()->red_a + 2, rs, stream());
        }
        allreduce_sum(ctl.get()->red_a + 2, 1);
        launch_set_beta(ctl.get(), ctl.get()->red_a + 2, 1, stream());
        prof.launches++;
        h_beta = read_status()->beta;
        k = knew;
        x_published = false;  // the restart GEMM rewrote f
    }

    void finish_timing()
    {
        SB200_CUDA_CHECK(cudaEventRecord(ev_end, stream()));
        SB200_CUDA_CHECK(cudaEventSynchronize(ev_end));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ev_begin, ev_end);
        stats.ms_total = ms;
        stats.kernel_launches = prof.launches;
        stats.ms_spmv = prof.ms[KC_SPMV];
        stats.ms_panel = prof.ms[KC_PANEL];
        stats.ms_compress = prof.ms[KC_COMPRESS];
        stats.ms_small = prof.ms[KC_SMALL];
        stats.ms_comm = prof.ms[KC_COMM];
    }

    void get_factorization(double* Vh, double* Hh, double* fh, double* beta, int64_t* kk)
    {
        cudaStream_t st = stream();
        if (Vh && nloc > 0)
            SB200_CUDA_CHECK(cudaMemcpy2DAsync(Vh, sizeof(double) * nloc, V.get(), sizeof(double) * ld, sizeof(double) * nloc, m, cudaMemcpyDeviceToHost, st));
        std::vector<double> hre, him;
        if (Hh && complex_h)
        {
            // complex Arnoldi: Hh receives m x m interleaved (re, im) values
            hre.resize((size_t) m * m);
            him.resize((size_t) m * m);
            SB200_CUDA_CHECK(cudaMemcpyAsync(hre.data(), H.get(), sizeof(double) * m * m, cudaMemcpyDeviceToHost, st));
            SB200_CUDA_CHECK(cudaMemcpyAsync(him.data(), Hi.get(), sizeof(double) * m * m, cudaMemcpyDeviceToHost, st));
        }
        else if (Hh)
            SB200_CUDA_CHECK(cudaMemcpyAsync(Hh, H.get(), sizeof(double) * m * m, cudaMemcpyDeviceToHost, st));
        if (fh && nloc > 0)
            SB200_CUDA_CHECK(cudaMemcpyAsync(fh, f.get(), sizeof(double) * nloc, cudaMemcpyDeviceToHost, st));
        SB200_CUDA_CHECK(cudaStreamSynchronize(st));
        for (size_t q = 0; q < hre.size(); q++)
        {
            Hh[2 * q] = hre[q];
            Hh[2 * q + 1] = him[q];
        }
        if (beta)
            *beta = h_beta;
        if (kk)
            *kk = k;
    }
};

}  // namespace sb200

// Host driver of the implicitly restarted Lanczos solver (SymEigsSolver / HermEigsBase).
//
// The host thread only sequences kernels and reads back small status words; V, H, f, the CSR
// operand and all Ritz data stay in HBM for the whole solve.  Control flow follows the reference
// line by line (citations relative to /root/reference/include/Spectra/):
//   init                Arnoldi.h:136-195 (via HermEigsBase.h:309-342)            -> fac_base.h
//   factorize_from      Lanczos.h:62-187   -> fused kernels K-A (spmv.cu) and K-B/K-C (panel.cu)
//   expand_basis        Arnoldi.h:66-115   (rare path, host sequenced)             -> fac_base.h
//   restart             HermEigsBase.h:105-155 -> dense_sym.cu + compress GEMM
//   compute             HermEigsBase.h:366-390
//   sort_ritzpair       HermEigsBase.h:229-251 (+ SymEigsShiftSolver.h:163-169)
//   eigenvalues/vectors HermEigsBase.h:417-470
#include "fac_base.h"

namespace sb200 {

void launch_sym_restart_beta(double* H, int m, int nev, double beta, int selection, double tol, double* ritz_val, double* ritz_est, double* ritz_vec,
                             int* ritz_conv, double* Q, SymRestartOut* out, cudaStream_t stream);

__global__ void trim_h_kernel(double* H, int m, int from_k)
{
    // keep the leading from_k x from_k block (Lanczos.h:85-86, Arnoldi.h:219-220)
    for (int t = threadIdx.x + blockIdx.x * blockDim.x; t < m * m; t += blockDim.x * gridDim.x)
    {
        const int r = t % m, c = t / m;
        if (r >= from_k || c >= from_k)
            H[t] = 0.0;
    }
}

void launch_trim_h(double* H, int m, int from_k, cudaStream_t s)
{
    trim_h_kernel<<<(m * m + 255) / 256, 256, 0, s>>>(H, m, from_k);
    SB200_CUDA_CHECK(cudaGetLastError());
}

}  // namespace sb200

using namespace sb200;

struct sb200_sym_solver : public FacBase
{
    bool shift_mode = false;
    double sigma = 0.0;

    DevBuf<double> ritz_val, ritz_est, ritz_vec;
    DevBuf<int> ritz_conv;
    DevBuf<SymRestartOut> rout;
    std::vector<double> h_ritz_val, h_ritz_vec;
    std::vector<int> h_ritz_conv;

    void init(const double* init_resid)
    {
        ritz_val.zero(stream());
        ritz_est.zero(stream());
        ritz_vec.zero(stream());
        ritz_conv.zero(stream());
        h_ritz_val.assign(m, 0.0);
        h_ritz_vec.assign((size_t) m * nev, 0.0);
        h_ritz_conv.assign(nev, 0);
        init_factorization(init_resid);
    }

    // tail of one Lanczos step on the host: further corrections f -= V Vf while the device asks for them (Lanczos.h:156-182; rare),
    // the zeroed residual of :163-168, and the host copy of beta
    void finish_step(const FacCtl* st, int j, double beta_thresh)
    {
        while (st->need_corr)
        {
            panel(PANEL_CORR, j, f.get(), f.get(), ctl.get()->c);
            launch_lanczos_decide(ctl.get(), H.get(), m, beta_thresh, 2, stream(), 0, is_cplx());
            prof.launches++;
            st = read_status();
        }
        stats.reorth_passes += st->count;
        if (st->f_zeroed)
        {
            f.zero(stream());
            x_published = false;
        }
        h_beta = st->beta;
    }

    // ---- Lanczos::factorize_from (Lanczos.h:62-187) ----
    void factorize_from(int64_t from_k, int64_t to_m)
    {
        if (to_m <= from_k)
            return;
        if (from_k > k)
            throw Error(SB200_INVALID_ARGUMENT, "Lanczos: from_k (= " + std::to_string(from_k) + ") is larger than the current subspace dimension (= " +
                                                    std::to_string(k) + ")");
        const double beta_thresh = kEps * std::sqrt(double(n / cw));  // m_n counts scalars
        const double eps_sqrt = std::sqrt(kEps);
        launch_trim_h(H.get(), m, (int) from_k, stream());
        prof.launches++;

        int i = (int) from_k;
        // Sweep mode (common path): all remaining steps are enqueued without a host round trip; the status is read once at the end.  A
        // step that needs the host -- a second correction pass, a zeroed residual, beta below sqrt(eps) (the restart tests at the head
        // of the next step) -- raises FacCtl::abort on the device, the kernels enqueued behind it return at once, and the loop below
        // finishes that step and the rest of the factorisation one step at a time.
        if (sweep_capable() && i <= (int) to_m - 1 && h_beta >= eps_sqrt)
        {
            struct Snap
            {
                int64_t nmatop;
                sb200_stats stats;
                int64_t launches;
            };
            std::vector<Snap> snap;
            in_sweep = true;
            if (overlap_capable())
            {
                // two-part correction pass; the head blocks of step s + 1 run on the second stream beside the second part of step s
                ensure_overlap_resources();
                const int64_t h = overlap_split_rows();
                double* wother = (wp == w.get()) ? w_alt.get() : w.get();
                if (peer && !x_published)
                    publish_f();
                launch_head_blocks(i, stream(), wp);
                bool aux_used = false;
                for (int s = i; s <= (int) to_m - 1; s++)
                {
                    snap.push_back({nmatop, stats, prof.launches});
                    stats.lanczos_steps++;
                    const int j = s + 1;
                    if (s > i)
                        SB200_CUDA_CHECK(cudaStreamWaitEvent(stream(), ev_k0, 0));  // head blocks of this step (second stream) are done
                    if (launch_last_block(s, true, wp))
                        allreduce_sum(ctl.get()->red, (size_t) j);
                    else
                        panel(PANEL_DOT, j, wp, nullptr, nullptr);
                    launch_lanczos_decide(ctl.get(), H.get(), m, beta_thresh, 0, stream(), 0, false);
                    stats.panel_launches++;
                    stats.panel_cols += j;
                    panel_corr_range(j, wp, ctl.get()->c, 0, h, ctl.get()->red2);
                    if (s < (int) to_m - 1)
                    {
                        SB200_CUDA_CHECK(cudaEventRecord(ev_part_a, stream()));
                        SB200_CUDA_CHECK(cudaStreamWaitEvent(aux_stream, ev_part_a, 0));
                        launch_head_blocks(s + 1, aux_stream, wother);
                        SB200_CUDA_CHECK(cudaEventRecord(ev_k0, aux_stream));
                        aux_used = true;
                    }
                    panel_corr_range(j, wp, ctl.get()->c, h, nloc, ctl.get()->red);
                    launch_lanczos_decide(ctl.get(), H.get(), m, beta_thresh, 1, stream(), 0, false, 1, 1);
                    prof.launches += 2;
                    stats.reorth_passes += 1;
                    std::swap(wp, wother);
                }
                if (aux_used)
                    SB200_CUDA_CHECK(cudaStreamWaitEvent(stream(), ev_k0, 0));
                if (peer)
                    x_published = true;  // both parts of the last correction pass wrote this rank's rows into every operand buffer
            }
            else
            {
                for (int s = i; s <= (int) to_m - 1; s++)
                {
                    snap.push_back({nmatop, stats, prof.launches});
                    stats.lanczos_steps++;
                    const int j = s + 1;
                    int count = j;  // values to combine over the ranks after the first panel pass
                    if (!spmv_step(s, false, true, true))
                    {
                        panel(PANEL_DOT, j, wp, nullptr, nullptr, nullptr, false);
                        count = kRedNrm + 1;
                    }
                    decide_after(0, count, beta_thresh, 0);
                    panel(PANEL_CORR, j, wp, f.get(), ctl.get()->c, nullptr, false);
                    decide_after(1, kRedNrm + 1, beta_thresh, 1);
                    stats.reorth_passes += 1;
                }
            }
            in_sweep = false;
            const FacCtl* st = read_status();
            if (!st->abort)
            {
                h_beta = st->beta;
                k = to_m;
                return;
            }
            // step st->i ran up to its first correction pass and raised the flag; nothing after it executed
            const int ia = st->i;
            const Snap& sn = snap[(size_t) (ia + 1 - i)];  // counters as they stood before step ia + 1 was enqueued
            if (ia + 1 <= (int) to_m - 1)
            {
                nmatop = sn.nmatop;
                const int64_t syncs = stats.host_syncs;
                stats = sn.stats;
                stats.host_syncs = syncs;
                prof.launches = sn.launches;
            }
            stats.reorth_passes -= 1;  // re-counted from the device counter below
            clear_abort();
            finish_step(st, ia + 1, beta_thresh);
            i = ia + 1;
        }
        for (; i <= (int) to_m - 1; i++)
        {
            stats.lanczos_steps++;
            bool restart = (h_beta < kNear0);
            if (!restart && h_beta < eps_sqrt)
            {
                // (V_{i-1}^H) v with v = f / beta   (Lanczos.h:107-113)
                double viv = reduce_scalar(VR_DOT, V.get() + (int64_t) (i - 1) * ld, f.get()) / h_beta;
                if (is_cplx())
                    viv = std::hypot(viv, reduce_scalar(VR_CDOT_IM, V.get() + (int64_t) (i - 1) * ld, f.get()) / h_beta);  // |Viv|, complex
                restart = (std::fabs(viv) > eps_sqrt);
            }
            if (restart)
                expand_basis(i, 2 * (int64_t) i);

            // K-A+B: v_i = f/beta, w = A v_i - H(i,i-1) v_{i-1} (Lanczos.h:106,127-139) and c = V[:, :i+1]^T w in the operator kernel
            // (sliced layout; otherwise a separate PANEL_DOT pass); c_i = <v_i, w> = H(i,i) (:142)
            const int j = i + 1;
            step_dot(i, restart, true);
            launch_lanczos_decide(ctl.get(), H.get(), m, beta_thresh, 0, stream(), 0, is_cplx());
            // K-C: f = w - V c, beta = ||f||, Vf = V^T f in one pass (Lanczos.h:145-153 and the first correction :171-179, see
            // lanczos_decide_kernel); then the test of :156 on the device
            panel(PANEL_CORR, j, wp, f.get(), ctl.get()->c);
            launch_lanczos_decide(ctl.get(), H.get(), m, beta_thresh, 1, stream(), 0, is_cplx());
            prof.launches += 2;
            finish_step(read_status(), j, beta_thresh);
        }
        k = to_m;
    }

    SymRestartOut run_restart_kernel(int selection, double tol, int do_restart)
    {
        {
            ScopedKernelTimer t(&prof, stream(), KC_SMALL);
            launch_sym_restart(H.get(), m, nev, ctl.get(), selection, tol, ritz_val.get(), ritz_est.get(), ritz_vec.get(), ritz_conv.get(), Q.get(), rout.get(),
                               do_restart, stream());
        }
        SB200_CUDA_CHECK(cudaMemcpyAsync(hstat.get(), rout.get(), sizeof(SymRestartOut), cudaMemcpyDeviceToHost, stream()));
        SB200_CUDA_CHECK(cudaStreamSynchronize(stream()));
        stats.host_syncs++;
        SymRestartOut o = *reinterpret_cast<const SymRestartOut*>(hstat.get());
        if (o.info != 0)
            throw Error(SB200_RUNTIME, "TridiagEigen: eigen decomposition failed");
        return o;
    }

    // ---- HermEigsBase::compute (HermEigsBase.h:366-390) ----
    int64_t compute(int selection, int64_t maxit, double tol, int sorting)
    {
        SB200_REQUIRE(initialised, SB200_LOGIC, "init() must be called before compute()");
        switch (selection)
        {
            case SB200_LARGEST_MAGN:
            case SB200_LARGEST_ALGE:
            case SB200_SMALLEST_MAGN:
            case SB200_SMALLEST_ALGE:
            case SB200_BOTH_ENDS: break;
            default: throw Error(SB200_INVALID_ARGUMENT, "unsupported selection rule");
        }
        factorize_from(1, m);
        int64_t i, nconv = 0;
        for (i = 0; i < maxit; i++)
        {
            // retrieve_ritzpair + num_converged + (nev_adjusted, shifted-QR chain) in one device kernel
            const SymRestartOut o = run_restart_kernel(selection, tol, 1);
            nconv = o.nconv;
            if (nconv >= nev)
                break;
            // restart(nev_adj): compress_V, factorize_from  (HermEigsBase.h:148-152)
            stats.restarts++;
            if (o.k < m)
            {
                compress_v(o.k);
                factorize_from(o.k, m);
            }
        }
        if (i == maxit)
        {
            // the last restart() ended with retrieve_ritzpair(); the convergence flags stay those of
            // the last num_converged() call (HermEigsBase.h:374-382)
            std::vector<int> conv_keep(nev);
            SB200_CUDA_CHECK(cudaMemcpyAsync(conv_keep.data(), ritz_conv.get(), sizeof(int) * nev, cudaMemcpyDeviceToHost, stream()));
            SB200_CUDA_CHECK(cudaStreamSynchronize(stream()));
            stats.host_syncs++;
            run_restart_kernel(selection, tol, 0);
            SB200_CUDA_CHECK(cudaMemcpyAsync(ritz_conv.get(), conv_keep.data(), sizeof(int) * nev, cudaMemcpyHostToDevice, stream()));
        }
        sort_ritzpair(sorting);
        niter += (i + 1);
        info = (nconv >= nev) ? SB200_SUCCESSFUL : SB200_NOT_CONVERGING;
        finish_timing();
        return std::min<int64_t>(nev, nconv);
    }

    // ---- sort_ritzpair (HermEigsBase.h:229-251; SymEigsShiftSolver.h:163-169) ----
    void sort_ritzpair(int sort_rule)
    {
        SB200_CUDA_CHECK(cudaMemcpyAsync(h_ritz_val.data(), ritz_val.get(), sizeof(double) * m, cudaMemcpyDeviceToHost, stream()));
        SB200_CUDA_CHECK(cudaMemcpyAsync(h_ritz_vec.data(), ritz_vec.get(), sizeof(double) * m * nev, cudaMemcpyDeviceToHost, stream()));
        SB200_CUDA_CHECK(cudaMemcpyAsync(h_ritz_conv.data(), ritz_conv.get(), sizeof(int) * nev, cudaMemcpyDeviceToHost, stream()));
        SB200_CUDA_CHECK(cudaStreamSynchronize(stream()));
        stats.host_syncs++;
        if (shift_mode)
            for (int i = 0; i < nev; i++)
                h_ritz_val[i] = 1.0 / h_ritz_val[i] + sigma;
        if ((sort_rule != SB200_LARGEST_ALGE) && (sort_rule != SB200_LARGEST_MAGN) && (sort_rule != SB200_SMALLEST_ALGE) && (sort_rule != SB200_SMALLEST_MAGN))
            throw Error(SB200_INVALID_ARGUMENT, "unsupported sorting rule");
        std::vector<int> ind(nev);
        std::iota(ind.begin(), ind.end(), 0);
        auto key = [&](int q) {
            const double v = h_ritz_val[q];
            switch (sort_rule)
            {
                case SB200_LARGEST_MAGN: return -std::fabs(v);
                case SB200_LARGEST_ALGE: return -v;
                case SB200_SMALLEST_MAGN: return std::fabs(v);
                default: return v;
            }
        };
        std::sort(ind.begin(), ind.end(), [&](int a, int b) { return key(a) < key(b); });
        std::vector<double> nv(m, 0.0), nvec((size_t) m * nev);
        std::vector<int> nc(nev);
        for (int i = 0; i < nev; i++)
        {
            nv[i] = h_ritz_val[ind[i]];
            std::copy(h_ritz_vec.begin() + (size_t) ind[i] * m, h_ritz_vec.begin() + (size_t) (ind[i] + 1) * m, nvec.begin() + (size_t) i * m);
            nc[i] = h_ritz_conv[ind[i]];
        }
        h_ritz_val.swap(nv);
        h_ritz_vec.swap(nvec);
        h_ritz_conv.swap(nc);
    }

    int64_t count_conv() const
    {
        int64_t c = 0;
        for (int v : h_ritz_conv)
            c += v ? 1 : 0;
        return c;
    }

    // ---- eigenvectors (HermEigsBase.h:447-470): X = V * ritz_vec_conv on the device ----
    int64_t eigenvectors_device(int64_t nvec)
    {
        const int64_t nconv = count_conv();
        nvec = std::min(nvec, nconv);
        if (nvec <= 0)
            return 0;
        std::vector<double> sel((size_t) m * m, 0.0);
        int64_t j = 0;
        for (int i = 0; i < nev && j < nvec; i++)
            if (h_ritz_conv[i])
            {
                std::copy(h_ritz_vec.begin() + (size_t) i * m, h_ritz_vec.begin() + (size_t) (i + 1) * m, sel.begin() + (size_t) j * m);
                j++;
            }
        SB200_CUDA_CHECK(cudaMemcpyAsync(S.get(), sel.data(), sizeof(double) * m * m, cudaMemcpyHostToDevice, stream()));
        SB200_CUDA_CHECK(cudaStreamSynchronize(stream()));
        stats.host_syncs++;  // `sel` is pageable stack-lifetime memory
        if (X.n < (size_t) ld * nev)
            X.alloc((size_t) ld * nev);
        {
            ScopedKernelTimer t(&prof, stream(), KC_COMPRESS);
            launch_compress(V.get(), ld, nloc, m, S.get(), (int) nvec, X.get(), ld, nullptr, nullptr, nullptr, rs, stream());
        }
        return nvec;
    }
};

namespace sb200 {

sb200_sym_solver* sym_create(sb200_op* op, int64_t nev, int64_t ncv, bool shift_mode, double sigma)
{
    device_info();
    SB200_REQUIRE(op != nullptr, SB200_INVALID_ARGUMENT, "null operator");
    const int64_t n = op->A.n;
    // HermEigsBase.h:257-272 — ncv is clamped to n before the checks
    const int64_t m = ncv > n ? n : ncv;
    if (nev < 1 || nev > n - 1)
        throw Error(SB200_INVALID_ARGUMENT, "nev must satisfy 1 <= nev <= n - 1, n is the size of matrix");
    if (ncv <= nev || ncv > n)
        throw Error(SB200_INVALID_ARGUMENT, "ncv must satisfy nev < ncv <= n, n is the size of matrix");
    SB200_REQUIRE(m <= kPanelMaxCols, SB200_INVALID_ARGUMENT, "this build supports ncv <= 64");
    SB200_REQUIRE(!op->cplx || m < kPanelMaxCols, SB200_INVALID_ARGUMENT, "this build supports ncv <= 63 for complex Hermitian operators");
    std::unique_ptr<sb200_sym_solver> s(new sb200_sym_solver());
    s->alloc_common(op, nev, m);
    s->shift_mode = shift_mode;
    s->sigma = sigma;
    s->ritz_val.alloc(m);
    s->ritz_est.alloc(m);
    s->ritz_vec.alloc((size_t) m * nev);
    s->ritz_conv.alloc(nev);
    s->rout.alloc(1);
    SB200_CUDA_CHECK(cudaStreamSynchronize(op->stream));
    return s.release();
}

void sym_init(sb200_sym_solver* s, const double* resid) { s->init(resid); }
int64_t sym_compute(sb200_sym_solver* s, int selection, int64_t maxit, double tol, int sorting) { return s->compute(selection, maxit, tol, sorting); }
void sym_factorize_from(sb200_sym_solver* s, int64_t from_k, int64_t to_m)
{
    SB200_REQUIRE(s->initialised, SB200_LOGIC, "init() must be called first");
    SB200_REQUIRE(to_m <= s->m, SB200_INVALID_ARGUMENT, "to_m exceeds ncv");
    s->factorize_from(from_k, to_m);
}
void sym_get_factorization(sb200_sym_solver* s, double* Vh, double* Hh, double* fh, double* beta, int64_t* kk) { s->get_factorization(Vh, Hh, fh, beta, kk); }

int64_t sym_eigenvalues(const sb200_sym_solver* s, double* out)
{
    int64_t j = 0;
    for (int i = 0; i < s->nev; i++)
        if (s->h_ritz_conv.size() > (size_t) i && s->h_ritz_conv[i])
            out[j++] = s->h_ritz_val[i];
    return j;
}

// local rows only: out is nloc x ncols column-major
int64_t sym_eigenvectors_local(sb200_sym_solver* s, int64_t nvec, double* out)
{
    const int64_t nc = s->eigenvectors_device(nvec);
    if (nc > 0 && s->nloc > 0)
        SB200_CUDA_CHECK(cudaMemcpy2DAsync(out, sizeof(double) * s->nloc, s->X.get(), sizeof(double) * s->ld, sizeof(double) * s->nloc, nc,
                                           cudaMemcpyDeviceToHost, s->stream()));
    SB200_CUDA_CHECK(cudaStreamSynchronize(s->stream()));
    return nc;
}

// full n rows on every rank (sharded runs all-gather column by column)
int64_t sym_eigenvectors_full(sb200_sym_solver* s, int64_t nvec, double* out)
{
    if (s->P() == 1)
        return sym_eigenvectors_local(s, nvec, out);
    const int64_t nc = s->eigenvectors_device(nvec);
    for (int64_t c = 0; c < nc; c++)
    {
        nccl_allgather(s->op->comm, s->X.get() + c * s->ld, s->op->x_full.get(), (size_t) s->op->slab, s->stream());
        SB200_CUDA_CHECK(cudaMemcpyAsync(out + c * s->n, s->op->x_full.get(), sizeof(double) * s->n, cudaMemcpyDeviceToHost, s->stream()));
    }
    SB200_CUDA_CHECK(cudaStreamSynchronize(s->stream()));
    return nc;
}

// unit-tier hook (sb200_dense_sym_restart)
void dense_sym_restart_host(int64_t m, const double* H, double beta, int64_t nev, int selection, double tol, double* ritz_val, double* ritz_est, int32_t* conv,
                            int64_t* nconv, int64_t* k, double* Q, double* Hnew)
{
    device_info();
    DevBuf<double> dH(m * m), dQ(m * m), drv(m), dre(m), dvec(m * nev);
    DevBuf<int> dconv(nev);
    DevBuf<SymRestartOut> dout(1);
    SB200_CUDA_CHECK(cudaMemcpy(dH.get(), H, sizeof(double) * m * m, cudaMemcpyHostToDevice));
    SB200_CUDA_CHECK(cudaMemset(dQ.get(), 0, sizeof(double) * m * m));
    launch_sym_restart_beta(dH.get(), (int) m, (int) nev, beta, selection, tol, drv.get(), dre.get(), dvec.get(), dconv.get(), dQ.get(), dout.get(), 0);
    SB200_CUDA_CHECK(cudaDeviceSynchronize());
    SymRestartOut o;
    SB200_CUDA_CHECK(cudaMemcpy(&o, dout.get(), sizeof(o), cudaMemcpyDeviceToHost));
    if (o.info != 0)
        throw Error(SB200_RUNTIME, "TridiagEigen: eigen decomposition failed");
    SB200_CUDA_CHECK(cudaMemcpy(ritz_val, drv.get(), sizeof(double) * m, cudaMemcpyDeviceToHost));
    SB200_CUDA_CHECK(cudaMemcpy(ritz_est, dre.get(), sizeof(double) * m, cudaMemcpyDeviceToHost));
    SB200_CUDA_CHECK(cudaMemcpy(conv, dconv.get(), sizeof(int) * nev, cudaMemcpyDeviceToHost));
    SB200_CUDA_CHECK(cudaMemcpy(Q, dQ.get(), sizeof(double) * m * m, cudaMemcpyDeviceToHost));
    SB200_CUDA_CHECK(cudaMemcpy(Hnew, dH.get(), sizeof(double) * m * m, cudaMemcpyDeviceToHost));
    *nconv = o.nconv;
    *k = o.k;
}

int sym_info(const sb200_sym_solver* s) { return s->info; }
int64_t sym_niter(const sb200_sym_solver* s) { return s->niter; }
int64_t sym_nops(const sb200_sym_solver* s) { return s->nmatop; }
const sb200_stats& sym_stats(const sb200_sym_solver* s) { return s->stats; }
void sym_destroy(sb200_sym_solver* s) { delete s; }

}  // namespace sb200

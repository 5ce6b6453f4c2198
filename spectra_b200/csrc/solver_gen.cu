// Host driver of the implicitly restarted Arnoldi solver (GenEigsSolver / GenEigsBase, real double).
//
// Same structure as solver_sym.cu; citations relative to /root/reference/include/Spectra/:
//   init                Arnoldi.h:136-195 (via GenEigsBase.h:442-475)             -> fac_base.h
//   factorize_from      Arnoldi.h:198-295  -> K-A (spmv.cu), panel DOT + CORR passes (panel.cu)
//   restart             GenEigsBase.h:204-222 -> dense_gen.cu (RestartArnoldi::run :60-107) + compress GEMM
//   compute             GenEigsBase.h:501-525
//   sort_ritzpair       GenEigsBase.h:345-404
//   eigenvalues/vectors GenEigsBase.h:531-611 (complex results, interleaved re/im at the ABI)
#include <complex>

#include "fac_base.h"

namespace sb200 {

struct GenRestartOut
{
    int nconv;
    int k;
    int info;
    int pad;
};
void launch_gen_restart(double* H, int m, int nev, const FacCtl* ctl, double beta, int use_beta, int selection, double tol, double* ritz_val_ri,
                        double* ritz_est_ri, double* ritz_vec_ri, int* ritz_conv, double* Q, GenRestartOut* out, int do_restart, cudaStream_t stream);
// dense_gen_z.cu: the complex restart kernel (H and Q as separate real / imaginary m x m arrays)
void launch_gen_restart_z(double* Hr, double* Hi, int m, int nev, const FacCtl* ctl, double beta, int use_beta, int selection, double tol, double* ritz_val_ri,
                          double* ritz_est_ri, double* ritz_vec_ri, int* ritz_conv, double* Qr, double* Qi, GenRestartOut* out, int do_restart,
                          cudaStream_t stream);

}  // namespace sb200

using namespace sb200;
using cplx = std::complex<double>;

struct sb200_gen_solver : public FacBase
{
    DevBuf<double> ritz_val, ritz_est, ritz_vec;  // interleaved complex: 2m, 2m, 2 m nev
    DevBuf<int> ritz_conv;
    DevBuf<GenRestartOut> rout;
    std::vector<cplx> h_ritz_val, h_ritz_vec;
    std::vector<int> h_ritz_conv;

    void init(const double* init_resid)
    {
        ritz_val.zero(stream());
        ritz_est.zero(stream());
        ritz_vec.zero(stream());
        ritz_conv.zero(stream());
        h_ritz_val.assign(m, cplx(0, 0));
        h_ritz_vec.assign((size_t) m * nev, cplx(0, 0));
        h_ritz_conv.assign(nev, 0);
        init_factorization(init_resid);
    }

    // ---- Arnoldi::factorize_from (Arnoldi.h:198-295) ----
    void factorize_from(int64_t from_k, int64_t to_m)
    {
        if (to_m <= from_k)
            return;
        if (from_k > k)
            throw Error(SB200_INVALID_ARGUMENT, "Arnoldi: from_k (= " + std::to_string(from_k) + ") is larger than the current subspace dimension (= " +
                                                    std::to_string(k) + ")");
        const double beta_thresh = kEps * std::sqrt(double(n / cw));  // m_n counts scalars
        launch_trim_h(H.get(), m, (int) from_k, stream());
        if (complex_h)
            launch_trim_h(Hi.get(), m, (int) from_k, stream());
        prof.launches++;
        double* hi = complex_h ? Hi.get() : nullptr;

        int i = (int) from_k;
        // Sweep mode (real operators resident on the device; see solver_sym.cu): the remaining steps are enqueued without a host round
        // trip -- K-A(+B), decide, f = w - V h, DGKS decide, and the first correction pass with its decision predicated on the device --
        // and the status is read once.  A step that needs a third pass, zeroes the residual or ends with beta < near_0 raises
        // FacCtl::abort; the per-step loop below takes over from there.
        if (sweep_capable() && !complex_h && i <= (int) to_m - 1 && h_beta >= kNear0)
        {
            struct Snap
            {
                int64_t nmatop;
                sb200_stats stats;
                int64_t launches;
            };
            std::vector<Snap> snap;
            SB200_CUDA_CHECK(cudaMemsetAsync(&ctl.get()->acc_count, 0, 3 * sizeof(int), stream()));
            in_sweep = true;
            for (int s = i; s <= (int) to_m - 1; s++)
            {
                snap.push_back({nmatop, stats, prof.launches});
                stats.lanczos_steps++;
                const int j = s + 1;
                step_dot(s, false, false);
                launch_arnoldi_decide(ctl.get(), H.get(), m, beta_thresh, 0, stream(), 0, nullptr, 1);
                panel(PANEL_CORR, j, wp, f.get(), ctl.get()->c);
                launch_arnoldi_decide(ctl.get(), H.get(), m, beta_thresh, 1, stream(), 0, nullptr, 1);
                panel(PANEL_CORR, j, f.get(), f.get(), ctl.get()->c, &ctl.get()->need_corr);
                launch_arnoldi_decide(ctl.get(), H.get(), m, beta_thresh, 2, stream(), 1, nullptr, 1);
                prof.launches += 3;
            }
            in_sweep = false;
            const FacCtl* st = read_status();
            const int acc_count = st->acc_count, acc_skipped = st->acc_skipped, acc_cols = st->acc_skipped_cols;
            if (!st->abort)
            {
                stats.reorth_passes += acc_count;
                stats.panel_launches -= acc_skipped;
                stats.panel_cols -= acc_cols;
                h_beta = st->beta;
                k = to_m;
                return;
            }
            // step st->i raised the flag; nothing after its deciding kernel executed
            const int ia = st->i;
            if (ia + 1 <= (int) to_m - 1)
            {
                const Snap& sn = snap[(size_t) (ia + 1 - i)];
                nmatop = sn.nmatop;
                const int64_t syncs = stats.host_syncs;
                stats = sn.stats;
                stats.host_syncs = syncs;
                prof.launches = sn.launches;
            }
            stats.reorth_passes += acc_count;
            stats.panel_launches -= acc_skipped;
            stats.panel_cols -= acc_cols;
            clear_abort();
            if (st->count == 0)
                uncount_panel(ia + 1);  // the speculative pass of step ia did not run (the DGKS test or a zeroed residual ended the step)
            finish_step(st, ia + 1, beta_thresh, hi);
            i = ia + 1;
        }
        for (; i <= (int) to_m - 1; i++)
        {
            stats.lanczos_steps++;
            bool restart = false;
            if (h_beta < kNear0)
            {
                expand_basis(i, 2 * (int64_t) i);
                restart = true;
            }
            // K-A: v_i = f/beta, H(i,i-1) = beta (or 0), w = A v_i   (Arnoldi.h:236-243)
            // and h = V^T w  (:251), in the same kernel when the operand carries the sliced layout
            const int j = i + 1;
            step_dot(i, restart, false);
            launch_arnoldi_decide(ctl.get(), H.get(), m, beta_thresh, 0, stream(), 0, hi);
            // f = w - V h, beta, and V^T f for the DGKS test in the same pass  (:254-262)
            panel(PANEL_CORR, j, wp, f.get(), ctl.get()->c);
            launch_arnoldi_decide(ctl.get(), H.get(), m, beta_thresh, 1, stream(), 0, hi);
            prof.launches += 2;
            // correction passes (:266-290); the first one is enqueued speculatively (device-side predicate)
            panel(PANEL_CORR, j, f.get(), f.get(), ctl.get()->c, &ctl.get()->need_corr);
            launch_arnoldi_decide(ctl.get(), H.get(), m, beta_thresh, 2, stream(), 1, hi);
            prof.launches++;
            const FacCtl* st = read_status();
            if (st->count == 0)
                uncount_panel(j);
            finish_step(st, j, beta_thresh, hi);
        }
        k = to_m;
    }

    // tail of one Arnoldi step on the host: further corrections while the device asks for them (Arnoldi.h:266-290), the zeroed residual of
    // :273-278, and the host copy of beta
    void finish_step(const FacCtl* st, int j, double beta_thresh, double* hi)
    {
        while (st->need_corr)
        {
            panel(PANEL_CORR, j, f.get(), f.get(), ctl.get()->c);  // (:281-287)
            launch_arnoldi_decide(ctl.get(), H.get(), m, beta_thresh, 2, stream(), 0, hi);
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

    GenRestartOut run_restart_kernel(int selection, double tol, int do_restart)
    {
        {
            ScopedKernelTimer t(&prof, stream(), KC_SMALL);
            if (complex_h)
                // complex Hessenberg matrix: Q comes back as Q (real parts) and S (imaginary parts; S is free between eigenvector calls)
                launch_gen_restart_z(H.get(), Hi.get(), m, nev, ctl.get(), 0.0, 0, selection, tol, ritz_val.get(), ritz_est.get(), ritz_vec.get(), ritz_conv.get(),
                                     Q.get(), S.get(), rout.get(), do_restart, stream());
            else
                launch_gen_restart(H.get(), m, nev, ctl.get(), 0.0, 0, selection, tol, ritz_val.get(), ritz_est.get(), ritz_vec.get(), ritz_conv.get(), Q.get(),
                                   rout.get(), do_restart, stream());
        }
        SB200_CUDA_CHECK(cudaMemcpyAsync(hstat.get(), rout.get(), sizeof(GenRestartOut), cudaMemcpyDeviceToHost, stream()));
        SB200_CUDA_CHECK(cudaStreamSynchronize(stream()));
        stats.host_syncs++;
        GenRestartOut o = *reinterpret_cast<const GenRestartOut*>(hstat.get());
        if (o.info != 0)
            throw Error(SB200_RUNTIME, "UpperHessenbergSchur: Schur decomposition failed");
        return o;
    }

    // ---- Arnoldi::compress_V with a complex Q (Arnoldi.h:320-340): V Q = V Re(Q) + i V Im(Q) -- two runs of the real restart GEMM on
    // the interleaved basis (its rows are the real and imaginary parts), one combine pass, then the complex residual update ----
    DevBuf<double> Xz;
    void compress_v_z(int knew)
    {
        const int kk = knew + 1;
        stats.compress_launches++;
        stats.compress_cols += kk;
        if (Xz.n < (size_t) 2 * ld * m)
            Xz.alloc((size_t) 2 * ld * m);
        double* A = Xz.get();
        double* B = Xz.get() + (size_t) ld * m;
        {
            ScopedKernelTimer t(&prof, stream(), KC_COMPRESS, 4);
            launch_compress(V.get(), ld, nloc, m, Q.get(), kk, A, ld, nullptr, nullptr, nullptr, rs, stream());
            launch_compress(V.get(), ld, nloc, m, S.get(), kk, B, ld, nullptr, nullptr, nullptr, rs, stream());
            launch_zcombine(A, B, ld, V.get(), ld, nloc, kk, stream());
            launch_zf_update(f.get(), V.get() + (int64_t) knew * ld, Q.get(), S.get(), H.get(), Hi.get(), m, kk, nloc, stream());
        }
        launch_vec_reduce(VR_SUMSQ, f.get(), nullptr, nloc, ctl.get()->red_a + 2, rs, stream());
        launch_set_beta(ctl.get(), ctl.get()->red_a + 2, 1, stream());
        prof.launches += 2;
        h_beta = read_status()->beta;
        k = knew;
    }

    static void check_rule(int rule, const char* what)
    {
        switch (rule)
        {
            case SB200_LARGEST_MAGN:
            case SB200_LARGEST_REAL:
            case SB200_LARGEST_IMAG:
            case SB200_SMALLEST_MAGN:
            case SB200_SMALLEST_REAL:
            case SB200_SMALLEST_IMAG: return;
            default: throw Error(SB200_INVALID_ARGUMENT, what);
        }
    }

    // ---- GenEigsBase::compute (GenEigsBase.h:501-525) ----
    int64_t compute(int selection, int64_t maxit, double tol, int sorting)
    {
        // complex operators: GenEigsBase with Scalar = std::complex<double> (:111-140 restart, :204-277, :280-404) -- the same loop with the
        // complex restart kernel (dense_gen_z.cu) and the complex restart GEMM (compress_v_z)
        SB200_REQUIRE(initialised, SB200_LOGIC, "init() must be called before compute()");
        check_rule(selection, "unsupported selection rule");
        factorize_from(1, m);
        int64_t i, nconv = 0;
        for (i = 0; i < maxit; i++)
        {
            const GenRestartOut o = run_restart_kernel(selection, tol, 1);
            nconv = o.nconv;
            if (nconv >= nev)
                break;
            stats.restarts++;
            if (o.k < m)
            {
                if (complex_h)
                    compress_v_z(o.k);
                else
                    compress_v(o.k);
                factorize_from(o.k, m);
            }
        }
        if (i == maxit)
        {
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

    // ---- sort_ritzpair (GenEigsBase.h:345-404) ----
    void sort_ritzpair(int sort_rule)
    {
        check_rule(sort_rule, "unsupported sorting rule");
        std::vector<double> rv(2 * m), rvec((size_t) 2 * m * nev);
        SB200_CUDA_CHECK(cudaMemcpyAsync(rv.data(), ritz_val.get(), sizeof(double) * 2 * m, cudaMemcpyDeviceToHost, stream()));
        SB200_CUDA_CHECK(cudaMemcpyAsync(rvec.data(), ritz_vec.get(), sizeof(double) * 2 * m * nev, cudaMemcpyDeviceToHost, stream()));
        SB200_CUDA_CHECK(cudaMemcpyAsync(h_ritz_conv.data(), ritz_conv.get(), sizeof(int) * nev, cudaMemcpyDeviceToHost, stream()));
        SB200_CUDA_CHECK(cudaStreamSynchronize(stream()));
        stats.host_syncs++;
        std::vector<cplx> val(m), vec((size_t) m * nev);
        for (int q = 0; q < m; q++)
            val[q] = cplx(rv[2 * q], rv[2 * q + 1]);
        for (size_t q = 0; q < vec.size(); q++)
            vec[q] = cplx(rvec[2 * q], rvec[2 * q + 1]);
        auto key = [&](int q) {
            const cplx v = val[q];
            switch (sort_rule)
            {
                case SB200_LARGEST_MAGN: return -std::abs(v);
                case SB200_LARGEST_REAL: return -v.real();
                case SB200_LARGEST_IMAG: return -std::fabs(v.imag());
                case SB200_SMALLEST_MAGN: return std::abs(v);
                case SB200_SMALLEST_REAL: return v.real();
                default: return std::fabs(v.imag());
            }
        };
        std::vector<int> ind(nev);
        std::iota(ind.begin(), ind.end(), 0);
        std::sort(ind.begin(), ind.end(), [&](int a, int b) { return key(a) < key(b); });
        h_ritz_val.assign(m, cplx(0, 0));
        h_ritz_vec.assign((size_t) m * nev, cplx(0, 0));
        std::vector<int> nc(nev);
        for (int i = 0; i < nev; i++)
        {
            h_ritz_val[i] = val[ind[i]];
            std::copy(vec.begin() + (size_t) ind[i] * m, vec.begin() + (size_t) (ind[i] + 1) * m, h_ritz_vec.begin() + (size_t) i * m);
            nc[i] = h_ritz_conv[ind[i]];
        }
        h_ritz_conv.swap(nc);
    }

    int64_t count_conv() const
    {
        int64_t c = 0;
        for (int v : h_ritz_conv)
            c += v ? 1 : 0;
        return c;
    }

    // ---- eigenvectors (GenEigsBase.h:561-603): [Re X | Im X] = V * [Re S | Im S], one GEMM with 2*nvec columns ----
    int64_t eigenvectors_device(int64_t nvec)
    {
        const int64_t nconv = count_conv();
        nvec = std::min(nvec, nconv);
        if (nvec <= 0)
            return 0;
        std::vector<double> sel((size_t) 2 * m * m, 0.0);
        int64_t j = 0;
        for (int i = 0; i < nev && j < nvec; i++)
            if (h_ritz_conv[i])
            {
                for (int r = 0; r < m; r++)
                {
                    sel[(size_t) j * m + r] = h_ritz_vec[(size_t) i * m + r].real();
                    sel[(size_t) (nvec + j) * m + r] = h_ritz_vec[(size_t) i * m + r].imag();
                }
                j++;
            }
        SB200_CUDA_CHECK(cudaMemcpyAsync(S.get(), sel.data(), sizeof(double) * 2 * m * m, cudaMemcpyHostToDevice, stream()));
        SB200_CUDA_CHECK(cudaStreamSynchronize(stream()));
        stats.host_syncs++;
        if (X.n < (size_t) ld * 2 * nev)
            X.alloc((size_t) ld * 2 * nev);
        // two GEMMs (real and imaginary coefficient blocks) so that the output width stays <= m
        {
            ScopedKernelTimer t(&prof, stream(), KC_COMPRESS, 2);
            launch_compress(V.get(), ld, nloc, m, S.get(), (int) nvec, X.get(), ld, nullptr, nullptr, nullptr, rs, stream());
            launch_compress(V.get(), ld, nloc, m, S.get() + (size_t) nvec * m, (int) nvec, X.get() + (size_t) nvec * ld, ld, nullptr, nullptr, nullptr, rs,
                            stream());
        }
        return nvec;
    }
};

namespace sb200 {

sb200_gen_solver* gen_create(sb200_op* op, int64_t nev, int64_t ncv)
{
    device_info();
    SB200_REQUIRE(op != nullptr, SB200_INVALID_ARGUMENT, "null operator");
    const int64_t n = op->A.n;
    const int64_t m = ncv > n ? n : ncv;  // GenEigsBase.h:413
    if (nev < 1 || nev > n - 2)
        throw Error(SB200_INVALID_ARGUMENT, "nev must satisfy 1 <= nev <= n - 2, n is the size of matrix");
    if (ncv < nev + 2 || ncv > n)
        throw Error(SB200_INVALID_ARGUMENT, "ncv must satisfy nev + 2 <= ncv <= n, n is the size of matrix");
    SB200_REQUIRE(m <= kPanelMaxCols, SB200_INVALID_ARGUMENT, "this build supports ncv <= 64");
    SB200_REQUIRE(!op->cplx || m < kPanelMaxCols, SB200_INVALID_ARGUMENT, "this build supports ncv <= 63 for complex operators");
    std::unique_ptr<sb200_gen_solver> s(new sb200_gen_solver());
    s->complex_h = op->cplx;
    s->alloc_common(op, nev, m);
    s->ritz_val.alloc(2 * m);
    s->ritz_est.alloc(2 * m);
    s->ritz_vec.alloc((size_t) 2 * m * nev);
    s->ritz_conv.alloc(nev);
    s->rout.alloc(1);
    SB200_CUDA_CHECK(cudaStreamSynchronize(op->stream));
    return s.release();
}

void gen_init(sb200_gen_solver* s, const double* resid) { s->init(resid); }
int64_t gen_compute(sb200_gen_solver* s, int selection, int64_t maxit, double tol, int sorting) { return s->compute(selection, maxit, tol, sorting); }
void gen_factorize_from(sb200_gen_solver* s, int64_t from_k, int64_t to_m)
{
    SB200_REQUIRE(s->initialised, SB200_LOGIC, "init() must be called first");
    SB200_REQUIRE(to_m <= s->m, SB200_INVALID_ARGUMENT, "to_m exceeds ncv");
    s->factorize_from(from_k, to_m);
}
void gen_get_factorization(sb200_gen_solver* s, double* Vh, double* Hh, double* fh, double* beta, int64_t* kk) { s->get_factorization(Vh, Hh, fh, beta, kk); }

int64_t gen_eigenvalues(const sb200_gen_solver* s, double* out_ri)
{
    int64_t j = 0;
    for (int i = 0; i < s->nev; i++)
        if (s->h_ritz_conv.size() > (size_t) i && s->h_ritz_conv[i])
        {
            out_ri[2 * j] = s->h_ritz_val[i].real();
            out_ri[2 * j + 1] = s->h_ritz_val[i].imag();
            j++;
        }
    return j;
}

// full n x ncols complex matrix, interleaved (re, im), column-major
int64_t gen_eigenvectors(sb200_gen_solver* s, int64_t nvec, double* out_ri)
{
    const int64_t nc = s->eigenvectors_device(nvec);
    if (nc <= 0)
        return 0;
    const int64_t n = s->n, nloc = s->nloc, row0 = s->op->A.row0;
    std::vector<double> re((size_t) std::max<int64_t>(nloc, 1) * nc), im((size_t) std::max<int64_t>(nloc, 1) * nc);
    cudaStream_t st = s->stream();
    if (s->P() == 1)
    {
        SB200_CUDA_CHECK(cudaMemcpy2DAsync(re.data(), sizeof(double) * nloc, s->X.get(), sizeof(double) * s->ld, sizeof(double) * nloc, nc, cudaMemcpyDeviceToHost,
                                           st));
        SB200_CUDA_CHECK(cudaMemcpy2DAsync(im.data(), sizeof(double) * nloc, s->X.get() + (size_t) nc * s->ld, sizeof(double) * s->ld, sizeof(double) * nloc, nc,
                                           cudaMemcpyDeviceToHost, st));
        SB200_CUDA_CHECK(cudaStreamSynchronize(st));
        if (s->complex_h)
        {
            // complex basis: `re` holds V Re(S) and `im` holds V Im(S), both as interleaved complex columns of n = 2 n_c doubles;
            // X = V Re(S) + i V Im(S)
            for (int64_t c = 0; c < nc; c++)
                for (int64_t r = 0; r + 1 < n; r += 2)
                {
                    const double pr = re[(size_t) (r + c * nloc)], pi = re[(size_t) (r + 1 + c * nloc)];
                    const double qr = im[(size_t) (r + c * nloc)], qi = im[(size_t) (r + 1 + c * nloc)];
                    out_ri[r + c * n] = pr - qi;
                    out_ri[r + 1 + c * n] = pi + qr;
                }
            return nc;
        }
        for (int64_t c = 0; c < nc; c++)
            for (int64_t r = 0; r < n; r++)
            {
                out_ri[2 * (r + c * n)] = re[(size_t) (r + c * nloc)];
                out_ri[2 * (r + c * n) + 1] = im[(size_t) (r + c * nloc)];
            }
        (void) row0;
        return nc;
    }
    // sharded: all-gather each real / imaginary column
    std::vector<double> col((size_t) n);
    for (int part = 0; part < 2; part++)
        for (int64_t c = 0; c < nc; c++)
        {
            nccl_allgather(s->op->comm, s->X.get() + (size_t) (part * nc + c) * s->ld, s->op->x_full.get(), (size_t) s->op->slab, st);
            SB200_CUDA_CHECK(cudaMemcpyAsync(col.data(), s->op->x_full.get(), sizeof(double) * n, cudaMemcpyDeviceToHost, st));
            SB200_CUDA_CHECK(cudaStreamSynchronize(st));
            for (int64_t r = 0; r < n; r++)
                out_ri[2 * (r + c * n) + part] = col[(size_t) r];
        }
    return nc;
}

int gen_info(const sb200_gen_solver* s) { return s->info; }
int64_t gen_niter(const sb200_gen_solver* s) { return s->niter; }
int64_t gen_nops(const sb200_gen_solver* s) { return s->nmatop; }
const sb200_stats& gen_stats(const sb200_gen_solver* s) { return s->stats; }
void gen_destroy(sb200_gen_solver* s) { delete s; }

}  // namespace sb200

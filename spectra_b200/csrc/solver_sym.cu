// Host driver of the implicitly restarted Lanczos solver (SymEigsSolver / HermEigsBase).
//
// The host thread only sequences kernels and reads back small status words; V, H, f, the CSR
// operand and all Ritz data stay in HBM for the whole solve.  Control flow follows the reference
// line by line (citations relative to /root/reference/include/Spectra/):
//   init                Arnoldi.h:136-195 (via HermEigsBase.h:309-342)
//   factorize_from      Lanczos.h:62-187   -> fused kernels K-A (spmv.cu) and K-B/K-C (panel.cu)
//   expand_basis        Arnoldi.h:66-115   (rare path, host sequenced)
//   restart             HermEigsBase.h:105-155 -> dense_sym.cu + compress GEMM
//   compute             HermEigsBase.h:366-390
//   sort_ritzpair       HermEigsBase.h:229-251 (+ SymEigsShiftSolver.h:163-169)
//   eigenvalues/vectors HermEigsBase.h:417-470
#include <algorithm>
#include <cstring>
#include <numeric>

#include "host.h"

namespace sb200 {

void launch_sym_restart_beta(double* H, int m, int nev, double beta, int selection, double tol, double* ritz_val, double* ritz_est, double* ritz_vec,
                             int* ritz_conv, double* Q, SymRestartOut* out, cudaStream_t stream);

// Minimal LCG of Util/SimpleRandom.h:30-123 (default residual and expand_basis vectors)
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

struct sb200_sym_solver
{
    sb200_op* op = nullptr;
    int64_t n = 0, nloc = 0, ld = 0;
    int nev = 0, m = 0;
    bool shift_mode = false;
    double sigma = 0.0;

    DevBuf<double> V, f, w, t0, H, Q, ritz_val, ritz_est, ritz_vec, S, X;
    DevBuf<int> ritz_conv;
    DevBuf<FacCtl> ctl;
    DevBuf<SymRestartOut> rout;
    DevBuf<double> partials;
    DevBuf<unsigned int> ticket;
    RedScratch rs;
    PinnedBuf<char> hstat;  // status readback
    PinnedBuf<double> hred; // reduction readback (rare paths)

    // host-side algorithm state
    int64_t k = 0;  // m_k
    int64_t nmatop = 0, niter = 0;
    int info = SB200_NOT_COMPUTED;
    double h_beta = 0.0;
    bool initialised = false;
    std::vector<double> h_ritz_val, h_ritz_vec;
    std::vector<int> h_ritz_conv;

    Profiler prof;
    sb200_stats stats;
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr;

    cudaStream_t stream() const { return op->stream; }
    int P() const { return op->nranks(); }
    ~sb200_sym_solver()
    {
        if (ev_begin)
            cudaEventDestroy(ev_begin);
        if (ev_end)
            cudaEventDestroy(ev_end);
    }

    // ---- small helpers -----------------------------------------------------------------------
    const FacCtl* read_status()
    {
        SB200_CUDA_CHECK(cudaMemcpyAsync(hstat.get(), ctl.get(), kFacCtlStatusBytes, cudaMemcpyDeviceToHost, stream()));
        SB200_CUDA_CHECK(cudaStreamSynchronize(stream()));
        return reinterpret_cast<const FacCtl*>(hstat.get());
    }
    // all-reduce a device buffer across ranks (no-op on one GPU)
    void allreduce_sum(double* buf, size_t count)
    {
        if (P() > 1)
        {
            ScopedKernelTimer t(&prof, stream(), KC_COMM, 0);
            nccl_allreduce_sum(op->comm, buf, count, stream());
        }
    }
    void allreduce_max(double* buf, size_t count)
    {
        if (P() > 1)
        {
            ScopedKernelTimer t(&prof, stream(), KC_COMM, 0);
            nccl_allreduce_max(op->comm, buf, count, stream());
        }
    }
    // x_full <- all-gather of a local vector (sharded); returns the pointer SpMV must read
    const double* gather_full(const double* local)
    {
        if (P() == 1)
            return local;
        ScopedKernelTimer t(&prof, stream(), KC_COMM, 0);
        // the local vector has ld >= slab entries allocated; padding rows are zero
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
        if (opk == VR_MAXABS)
            allreduce_max(slot, 1);
        else
            allreduce_sum(slot, 1);
        SB200_CUDA_CHECK(cudaMemcpyAsync(hred.get(), slot, sizeof(double), cudaMemcpyDeviceToHost, stream()));
        SB200_CUDA_CHECK(cudaStreamSynchronize(stream()));
        return hred.get()[0];
    }
    void set_beta_host(double b)
    {
        h_beta = b;
        launch_set_scalar(&ctl.get()->beta, b, stream());
    }
    void panel(int mode, int j, const double* x, double* fo, const double* coef)
    {
        stats.panel_launches++;
        stats.panel_cols += j;
        {
            ScopedKernelTimer t(&prof, stream(), KC_PANEL);
            launch_panel_pass(mode, V.get(), ld, nloc, j, x, fo, coef, ctl.get()->red, rs, stream());
        }
        allreduce_sum(ctl.get()->red, kRedNrm + 1);
    }
    // host copy of red[0..j) and red[kRedNrm]
    void fetch_red(int j, double& ortho_err, double& nrm2)
    {
        SB200_CUDA_CHECK(cudaMemcpyAsync(hred.get(), ctl.get()->red, sizeof(double) * (kRedNrm + 1), cudaMemcpyDeviceToHost, stream()));
        SB200_CUDA_CHECK(cudaStreamSynchronize(stream()));
        ortho_err = 0.0;
        for (int q = 0; q < j; q++)
            ortho_err = std::max(ortho_err, std::fabs(hred.get()[q]));
        nrm2 = hred.get()[kRedNrm];
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
                // f = A * rand
                double* xfull = (P() > 1) ? op->x_full.get() : t0.get();
                SB200_CUDA_CHECK(cudaMemcpyAsync(xfull, rnd.data(), sizeof(double) * n, cudaMemcpyHostToDevice, stream()));
                {
                    ScopedKernelTimer t(&prof, stream(), KC_SPMV);
                    launch_spmv(op->A, op->plan, xfull, f.get(), stream());
                }
                stats.spmv_launches++;
                nmatop++;
            }
            else if (nloc > 0)
            {
                SB200_CUDA_CHECK(cudaMemcpyAsync(f.get(), rnd.data() + row0, sizeof(double) * nloc, cudaMemcpyHostToDevice, stream()));
            }
            // Vf = V^T f ; f -= V Vf ; fnorm ; Vf = V^T f   (:88-95)
            panel(PANEL_DOT, i, f.get(), nullptr, nullptr);
            SB200_CUDA_CHECK(cudaMemcpyAsync(ctl.get()->c, ctl.get()->red, sizeof(double) * i, cudaMemcpyDeviceToDevice, stream()));
            panel(PANEL_CORR, i, f.get(), f.get(), ctl.get()->c);
            double ortho_err, nrm2;
            fetch_red(i, ortho_err, nrm2);
            double fnorm = std::sqrt(nrm2);
            int count = 0;
            while (count < 3 && ortho_err >= kEps * fnorm)
            {
                SB200_CUDA_CHECK(cudaMemcpyAsync(ctl.get()->c, ctl.get()->red, sizeof(double) * i, cudaMemcpyDeviceToDevice, stream()));
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
    void init(const double* init_resid)
    {
        std::vector<double> gen;
        if (!init_resid)
        {
            // HermEigsBase.h:337-342: SimpleRandom<Scalar> rng(0); random_vec(n)
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
        ritz_val.zero(stream());
        ritz_est.zero(stream());
        ritz_vec.zero(stream());
        ritz_conv.zero(stream());
        ctl.zero(stream());
        nmatop = 0;
        niter = 0;
        k = 0;
        info = SB200_NOT_COMPUTED;
        h_ritz_val.assign(m, 0.0);
        h_ritz_vec.assign((size_t) m * nev, 0.0);
        h_ritz_conv.assign(nev, 0);

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
        {
            ScopedKernelTimer t(&prof, stream(), KC_SPMV);
            launch_spmv(op->A, op->plan, xfull, v, stream());  // v = A * v0
        }
        stats.spmv_launches++;
        nmatop++;
        const double vnorm = std::sqrt(reduce_scalar(VR_SUMSQ, v, nullptr));
        if (vnorm < kNear0)
            launch_vec_scale(t0.get(), v0norm, 1, v, nloc, stream());  // v = v0 / ||v0||
        else
            launch_vec_scale(v, vnorm, 1, v, nloc, stream());           // v /= ||v||
        prof.launches++;
        const double* vfull = gather_full(v);
        {
            ScopedKernelTimer t(&prof, stream(), KC_SPMV);
            launch_spmv(op->A, op->plan, vfull, w.get(), stream());  // w = A * v
        }
        stats.spmv_launches++;
        nmatop++;
        const double h00 = reduce_scalar(VR_DOT, v, w.get());
        launch_set_scalar(H.get(), h00, stream());
        launch_vec_axpy(w.get(), v, h00, f.get(), nloc, stream());  // f = w - v * H(0,0)
        prof.launches += 2;
        const double fmax = reduce_scalar(VR_MAXABS, f.get(), nullptr);
        if (fmax < kEps * std::fabs(h00))
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
    }

    // ---- Lanczos::factorize_from (Lanczos.h:62-187) ----
    void factorize_from(int64_t from_k, int64_t to_m)
    {
        if (to_m <= from_k)
            return;
        if (from_k > k)
            throw Error(SB200_INVALID_ARGUMENT, "Lanczos: from_k (= " + std::to_string(from_k) + ") is larger than the current subspace dimension (= " +
                                                    std::to_string(k) + ")");
        const double beta_thresh = kEps * std::sqrt(double(n));
        const double eps_sqrt = std::sqrt(kEps);
        launch_trim_h(H.get(), m, (int) from_k, stream());
        prof.launches++;

        for (int i = (int) from_k; i <= (int) to_m - 1; i++)
        {
            stats.lanczos_steps++;
            bool restart = (h_beta < kNear0);
            if (!restart && h_beta < eps_sqrt)
            {
                // (V_{i-1}^H) v with v = f / beta   (Lanczos.h:107-113)
                const double viv = reduce_scalar(VR_DOT, V.get() + (int64_t) (i - 1) * ld, f.get()) / h_beta;
                restart = (std::fabs(viv) > eps_sqrt);
            }
            if (restart)
                expand_basis(i, 2 * (int64_t) i);

            const double* xfull = gather_full(f.get());
            {
                ScopedKernelTimer t(&prof, stream(), KC_SPMV);
                launch_spmv_step(op->A, op->plan, xfull, f.get(), V.get(), ld, w.get(), ctl.get(), H.get(), m, i, restart ? 1 : 0, true, rs, stream());
            }
            stats.spmv_launches++;
            nmatop++;
            allreduce_sum(ctl.get()->red_a, 1);

            const int j = i + 1;
            panel(PANEL_FORM, j, w.get(), f.get(), ctl.get()->red_a);
            launch_lanczos_decide(ctl.get(), H.get(), m, beta_thresh, 1, stream());
            prof.launches++;
            const FacCtl* st = read_status();
            while (st->need_corr)
            {
                stats.reorth_passes++;
                panel(PANEL_CORR, j, f.get(), f.get(), ctl.get()->c);
                launch_lanczos_decide(ctl.get(), H.get(), m, beta_thresh, 0, stream());
                prof.launches++;
                st = read_status();
            }
            if (st->f_zeroed)
                f.zero(stream());
            h_beta = st->beta;
        }
        k = to_m;
    }

    // ---- HermEigsBase::restart tail: compress_V + factorize_from (HermEigsBase.h:148-152) ----
    void compress_and_expand(int knew)
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
        k = knew;  // compress_H decremented m_k once per shift (Lanczos.h:198-202)
        factorize_from(knew, m);
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
        SymRestartOut o;
        o.nconv = 0;
        for (i = 0; i < maxit; i++)
        {
            // retrieve_ritzpair + num_converged + (nev_adjusted, shifted QR chain) on the device
            o = run_restart_kernel(selection, tol, 1);
            nconv = o.nconv;
            if (nconv >= nev)
                break;
            stats.restarts++;
            if (o.k < m)
                compress_and_expand(o.k);
        }
        if (i == maxit && maxit > 0)
        {
            // the last restart() ended with retrieve_ritzpair(); convergence flags stay those of the
            // last num_converged() call (HermEigsBase.h:374-382)
            std::vector<int> conv_keep(nev);
            SB200_CUDA_CHECK(cudaMemcpyAsync(conv_keep.data(), ritz_conv.get(), sizeof(int) * nev, cudaMemcpyDeviceToHost, stream()));
            SB200_CUDA_CHECK(cudaStreamSynchronize(stream()));
            run_restart_kernel(selection, tol, 0);
            SB200_CUDA_CHECK(cudaMemcpyAsync(ritz_conv.get(), conv_keep.data(), sizeof(int) * nev, cudaMemcpyHostToDevice, stream()));
        }
        sort_ritzpair(sorting);
        niter += (i + 1);
        info = (nconv >= nev) ? SB200_SUCCESSFUL : SB200_NOT_CONVERGING;

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
        return std::min<int64_t>(nev, nconv);
    }

    // ---- sort_ritzpair (HermEigsBase.h:229-251; SymEigsShiftSolver.h:163-169) ----
    void sort_ritzpair(int sort_rule)
    {
        SB200_CUDA_CHECK(cudaMemcpyAsync(h_ritz_val.data(), ritz_val.get(), sizeof(double) * m, cudaMemcpyDeviceToHost, stream()));
        SB200_CUDA_CHECK(cudaMemcpyAsync(h_ritz_vec.data(), ritz_vec.get(), sizeof(double) * m * nev, cudaMemcpyDeviceToHost, stream()));
        SB200_CUDA_CHECK(cudaMemcpyAsync(h_ritz_conv.data(), ritz_conv.get(), sizeof(int) * nev, cudaMemcpyDeviceToHost, stream()));
        SB200_CUDA_CHECK(cudaStreamSynchronize(stream()));
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
        if (X.n < (size_t) ld * nvec)
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
    std::unique_ptr<sb200_sym_solver> s(new sb200_sym_solver());
    s->op = op;
    s->n = n;
    s->nloc = op->A.nrows;
    // leading dimension: multiple of 16 doubles (128 B) and at least the all-gather slab
    s->ld = round_up(std::max<int64_t>(std::max<int64_t>(s->nloc, op->slab), 2), 16);
    s->nev = (int) nev;
    s->m = (int) m;
    s->shift_mode = shift_mode;
    s->sigma = sigma;
    s->V.alloc((size_t) s->ld * m);
    s->f.alloc((size_t) s->ld);
    s->w.alloc((size_t) s->ld);
    s->t0.alloc((size_t) std::max<int64_t>(s->ld, n));
    s->H.alloc((size_t) m * m);
    s->Q.alloc((size_t) m * m);
    s->S.alloc((size_t) m * m);
    s->ritz_val.alloc(m);
    s->ritz_est.alloc(m);
    s->ritz_vec.alloc((size_t) m * nev);
    s->ritz_conv.alloc(nev);
    s->ctl.alloc(1);
    s->rout.alloc(1);
    const int max_grid = device_info().sm_count * 16;
    s->partials.alloc((size_t) max_grid * kRedStride);
    s->ticket.alloc(1);
    s->ticket.zero(op->stream);
    s->rs.partials = s->partials.get();
    s->rs.ticket = s->ticket.get();
    s->rs.max_grid = max_grid;
    s->hstat.alloc(256);
    s->hred.alloc(kRedStride + 8);
    SB200_CUDA_CHECK(cudaEventCreate(&s->ev_begin));
    SB200_CUDA_CHECK(cudaEventCreate(&s->ev_end));
    std::memset(&s->stats, 0, sizeof(s->stats));
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

void sym_get_factorization(sb200_sym_solver* s, double* Vh, double* Hh, double* fh, double* beta, int64_t* kk)
{
    cudaStream_t st = s->stream();
    if (Vh && s->nloc > 0)
        SB200_CUDA_CHECK(cudaMemcpy2DAsync(Vh, sizeof(double) * s->nloc, s->V.get(), sizeof(double) * s->ld, sizeof(double) * s->nloc, s->m,
                                           cudaMemcpyDeviceToHost, st));
    if (Hh)
        SB200_CUDA_CHECK(cudaMemcpyAsync(Hh, s->H.get(), sizeof(double) * s->m * s->m, cudaMemcpyDeviceToHost, st));
    if (fh && s->nloc > 0)
        SB200_CUDA_CHECK(cudaMemcpyAsync(fh, s->f.get(), sizeof(double) * s->nloc, cudaMemcpyDeviceToHost, st));
    SB200_CUDA_CHECK(cudaStreamSynchronize(st));
    if (beta)
        *beta = s->h_beta;
    if (kk)
        *kk = s->k;
}

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

}  // namespace sb200

// accessors used by abi.cu
namespace sb200 {
int sym_info(const sb200_sym_solver* s) { return s->info; }
int64_t sym_niter(const sb200_sym_solver* s) { return s->niter; }
int64_t sym_nops(const sb200_sym_solver* s) { return s->nmatop; }
int64_t sym_nloc(const sb200_sym_solver* s) { return s->nloc; }
int64_t sym_n(const sb200_sym_solver* s) { return s->n; }
int sym_nev(const sb200_sym_solver* s) { return s->nev; }
int sym_ncv(const sb200_sym_solver* s) { return s->m; }
const sb200_stats& sym_stats(const sb200_sym_solver* s) { return s->stats; }
void sym_destroy(sb200_sym_solver* s) { delete s; }
}  // namespace sb200

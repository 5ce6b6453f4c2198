// Single-CTA device kernels for the symmetric restart (K7, K8, K10 of SURVEY.md §2.1):
//   TridiagEigen::compute          LinAlg/TridiagEigen.h:121-210 (+ tridiagonal_qr_step :44-108)
//   TridiagQR::compute/matrix_QtHQ LinAlg/UpperHessenbergQR.h:515-598, :627-693 ; apply_YQ :383-417
//   retrieve_ritzpair / num_converged / nev_adjusted / shift loop of restart
//                                  HermEigsBase.h:205-224, :158-175, :178-202, :105-147
// so the whole restart decision stays on the device; the host reads back one 16-byte status.
//
// Parallelisation.  The scalar recurrences on the tridiagonal (rotation generation, deflation
// tests, shift strategy) are inherently sequential: thread 0 runs them back to back WITHOUT block
// synchronisation and appends every rotation to a log in shared memory.  Applying rotations to the
// m x m eigenvector / Q matrix is row-local, so when the log fills up (or the sequence ends) all
// threads replay it on their own row (matrix column-major in shared memory: a warp touches
// consecutive addresses).  The tridiagonal H is carried as (diag, subdiag) in shared memory.
//
// Arithmetic notes: rotations on the max-scaled iteration matrix use one reciprocal square root
// (c = p/r, s = -q/r, the closed form of Eigen's makeGivens); Givens<double>::compute_rotation keeps
// the reference's Taylor branch for tiny ratios and uses the same rsqrt form otherwise, falling back
// to the reference's hypot formulation outside the safe exponent range.
#include "dense_common.cuh"
#include "kernels.h"

namespace sb200 {

namespace {

using namespace dense;

constexpr int kDenseBlock = 128;
constexpr int kLogCap = 4096;   // rotations per log chunk
constexpr int kSweepCap = 512;  // sweeps per log chunk

struct TriShared
{
    double* d;    // working diag        [m]
    double* e;    // working subdiag     [m]
    double* t1;   // scratch             [m]
    double* t2;   // scratch             [m]
    double* ev;   // eigenvalues / keys  [m]
    double* aux;  // scratch             [m]
    int* idx;     // sort permutation    [m]
    double* Z;    // m x m
    double* lc;   // rotation log: cosines [kLogCap]
    double* ls;   // rotation log: sines   [kLogCap]
    int* hdr;     // sweep log: (first column, count, offset) triples [3 * kSweepCap]
};

__device__ __forceinline__ TriShared carve(double* smem, int m)
{
    TriShared s;
    s.d = smem;
    s.e = s.d + m;
    s.t1 = s.e + m;
    s.t2 = s.t1 + m;
    s.ev = s.t2 + m;
    s.aux = s.ev + m;
    s.Z = s.aux + m;
    s.lc = s.Z + m * m;
    s.ls = s.lc + kLogCap;
    s.idx = reinterpret_cast<int*>(s.ls + kLogCap);
    s.hdr = s.idx + 2 * ((m + 1) / 2);
    return s;
}
size_t tri_smem_bytes(int m)
{
    return sizeof(double) * (size_t) (6 * m + m * m + 2 * kLogCap) + sizeof(int) * (size_t) (2 * ((m + 1) / 2) + 3 * kSweepCap);
}

// makeGivens for the scaled QR iteration: entries are <= 1 in magnitude (TridiagEigen.h:139-152), so
// p^2 + q^2 cannot overflow; both branches of Eigen's makeGivens reduce to c = p/r, s = -q/r with
// r = +sqrt(p^2 + q^2).
__device__ __forceinline__ void make_givens_scaled(double p, double q, double& c, double& s)
{
    const double n2 = fma(p, p, q * q);
    if (q == 0.0 || p == 0.0 || n2 < 1e-280)
    {
        make_givens(p, q, c, s);
        return;
    }
    const double inv = rsqrt(n2);
    c = p * inv;
    s = -q * inv;
}

// Givens<double>::compute_rotation (Givens.h:166-205) with the rsqrt form for the regular case.
__device__ __forceinline__ void givens_rotation_fast(double x, double y, double& r, double& c, double& s)
{
    const double xabs = fabs(x), yabs = fabs(y);
    const double hi = fmax(xabs, yabs), lo = fmin(xabs, yabs);
    const double cutoff = 0.1 * 1.220703125e-4;
    if (x == 0.0 || y == 0.0 || lo < cutoff * hi || hi > 1e140 || hi < 1e-140)
    {
        givens_rotation(x, y, r, c, s);  // exact special cases, Taylor branch, extreme magnitudes
        return;
    }
    const double n2 = fma(x, x, y * y);
    const double inv = rsqrt(n2);
    r = n2 * inv;  // sqrt(x^2 + y^2) >= 0
    c = x * inv;   // c = x / r
    s = -y * inv;  // s = -y / r
}

// One implicit Wilkinson-shift QR step on rows start..end of the (scaled) tridiagonal (one thread; m = order of the matrix).
// Appends the rotations to (lc, ls) and returns their number.   TridiagEigen.h:44-108
// This loop is the serial critical path of the restart (about 4000 rotations for m = 60, one lane): the matrix entries the recurrence
// needs are carried in registers, the next column is prefetched one iteration ahead with clamped (branch-free) indices, and only the
// entries that are final after iteration k (diag[k], subdiag[k-1]) and the rotation log are stored inside the loop -- the three
// running entries are written once when the sweep ends.  Same operations in the same order as the straightforward form.
__device__ int tridiagonal_qr_step(double* diag, double* subdiag, int start, int end, int m, double* lc, double* ls)
{
    const double td = (diag[end - 1] - diag[end]) * 0.5;
    const double e = subdiag[end - 1];
    double mu = diag[end];
    if (td == 0.0)
        mu -= fabs(e);
    else if (e != 0.0)
    {
        const double e2 = e * e;
        const double h = eigen_hypot(td, e);
        if (e2 == 0.0)
            mu -= e / ((td + (td > 0.0 ? h : -h)) / e);
        else
            mu -= e2 / (td + (td > 0.0 ? h : -h));
    }
    double x = diag[start] - mu;
    double z = subdiag[start];
    double dk = diag[start], ek = subdiag[start];
    double dk1 = diag[start + 1];
    double ek1 = (start + 1 < end) ? subdiag[start + 1] : 0.0;
    double ekm1 = 0.0;
    double* pc = lc;
    double* ps = ls;
    int k = start;
    for (; k < end && z != 0.0; ++k)
    {
        // entries of column k + 2 (unused when the sweep stops before it needs them: the indices are clamped, not predicated)
        const double dk2 = diag[min(k + 2, m - 1)];
        const double ek2 = (k + 2 < end) ? subdiag[min(k + 2, m - 2)] : 0.0;
        double c, s;
        make_givens_scaled(x, z, c, s);
        const double sdk = s * dk + c * ek;
        const double dkp1 = s * ek + c * dk1;
        diag[k] = c * (c * dk - s * ek) - s * (c * ek - s * dk1);
        const double ndk1 = s * sdk + c * dkp1;
        const double nek = c * sdk - s * dkp1;
        if (k > start)
            subdiag[k - 1] = c * ekm1 - s * z;
        x = nek;
        if (k < end - 1)
        {
            z = -s * ek1;
            ek = c * ek1;
        }
        ekm1 = nek;
        dk = ndk1;
        dk1 = dk2;
        ek1 = ek2;
        *pc++ = c;
        *ps++ = s;
    }
    const int nrot = k - start;
    if (nrot > 0)
    {
        // running entries after the last rotation (k - 1): T(k, k-1) = nek, T(k, k) = ndk1, T(k+1, k) = c * ek1 when it exists
        subdiag[k - 1] = ekm1;
        diag[k] = dk;
        if (k < end)
            subdiag[k] = ek;
    }
    return nrot;
}

// All threads: replay the logged sweeps on their row of Z (q.applyOnTheRight, TridiagEigen.h:106).
__device__ void replay_sweeps(double* Z, int m, const double* lc, const double* ls, const int* hdr, int nsweeps)
{
    for (int t = threadIdx.x; t < m; t += kDenseBlock)
    {
        for (int sw = 0; sw < nsweeps; sw++)
        {
            const int st = hdr[3 * sw], nr = hdr[3 * sw + 1], off = hdr[3 * sw + 2];
            double xk = Z[t + st * m];
            for (int k = 0; k < nr; k++)
            {
                const double c = lc[off + k], s = ls[off + k];
                const double yk = Z[t + (st + k + 1) * m];
                Z[t + (st + k) * m] = c * xk - s * yk;
                xk = s * xk + c * yk;
            }
            Z[t + (st + nr) * m] = xk;
        }
    }
}

// TridiagEigen::compute on (hd, he) = diag / subdiag of H.  On exit sh.ev = eigenvalues (unsorted),
// sh.Z = eigenvectors.  Returns 0 on success, 1 when the 30*m sweep cap is hit.  Block-collective.
__device__ int tridiag_eigen_block(const double* hd, const double* he, int m, TriShared& sh)
{
    __shared__ int s_state[4];  // 0: more work after this chunk, 1: sweeps in chunk, 3: info
    __shared__ double s_scale;
    const int tid = threadIdx.x;
    for (int t = tid; t < m * m; t += kDenseBlock)
        sh.Z[t] = ((t % m) == (t / m)) ? 1.0 : 0.0;
    if (tid == 0)
    {
        double scale = 0.0;
        for (int i = 0; i < m; i++)
            scale = fmax(scale, fabs(hd[i]));
        for (int i = 0; i < m - 1; i++)
            scale = fmax(scale, fabs(he[i]));
        s_scale = scale;
        s_state[3] = 0;
    }
    __syncthreads();
    const double scale = s_scale;
    if (scale < kNear0)  // zero matrix: eigenvalues 0, vectors identity   (TridiagEigen.h:142-150)
    {
        for (int t = tid; t < m; t += kDenseBlock)
            sh.ev[t] = 0.0;
        __syncthreads();
        return 0;
    }
    for (int t = tid; t < m; t += kDenseBlock)
    {
        sh.d[t] = hd[t] / scale;
        if (t < m - 1)
            sh.e[t] = he[t] / scale;
    }
    __syncthreads();

    // Warp 0 runs the iteration: the deflation tests and the searches for the active block are spread over its lanes (votes), lane 0
    // generates the rotations; the loop state (start, end, iter, log fill) is kept identically in every lane.  The other warps follow s_state.
    int end = m - 1, start = 0, iter = 0;
    const double considerAsZero = kMin;
    const double precision_inv = 1.0 / kEps;
    const int lane = tid & 31;
    while (true)
    {
        if (tid < 32)
        {
            int nlog = 0, nsw = 0, more = 0;
            while (end > 0)
            {
                if (nlog + m > kLogCap || nsw >= kSweepCap)
                {
                    more = 1;  // log full: replay, then continue
                    break;
                }
                // deflation tests on the active block (TridiagEigen.h:165-178), one entry per lane
                for (int i = start + lane; i < end; i += 32)
                {
                    const double ei = sh.e[i];
                    if (fabs(ei) <= considerAsZero)
                        sh.e[i] = 0.0;
                    else
                    {
                        const double scaled = precision_inv * ei;
                        if (scaled * scaled <= (fabs(sh.d[i]) + fabs(sh.d[i + 1])))
                            sh.e[i] = 0.0;
                    }
                }
                __syncwarp();
                // while (end > 0 && e[end-1] == 0) end--;   -> end = 1 + (largest j < end with e[j] != 0), or 0
                {
                    int ne = 0;
                    for (int base = ((end - 1) >> 5) << 5; base >= 0; base -= 32)
                    {
                        const int j = base + lane;
                        const unsigned nz = __ballot_sync(0xffffffffu, j < end && sh.e[j] != 0.0);
                        if (nz)
                        {
                            ne = base + 32 - __clz((int) nz);
                            break;
                        }
                    }
                    end = ne;
                }
                if (end <= 0)
                    break;
                iter++;
                if (iter > 30 * m)
                {
                    if (lane == 0)
                        s_state[3] = 1;
                    break;
                }
                // start = end - 1; while (start > 0 && e[start-1] != 0) start--;   -> start = 1 + (largest j <= end-2 with e[j] == 0), or 0
                {
                    int ns = 0;
                    for (int base = ((end - 1) >> 5) << 5; base >= 0; base -= 32)
                    {
                        const int j = base + lane;
                        const unsigned zr = __ballot_sync(0xffffffffu, j <= end - 2 && sh.e[j] == 0.0);
                        if (zr)
                        {
                            ns = base + 32 - __clz((int) zr);
                            break;
                        }
                    }
                    start = ns;
                }
                int nrot = 0;
                if (lane == 0)
                {
                    nrot = tridiagonal_qr_step(sh.d, sh.e, start, end, m, sh.lc + nlog, sh.ls + nlog);
                    if (nrot > 0)
                    {
                        sh.hdr[3 * nsw] = start;
                        sh.hdr[3 * nsw + 1] = nrot;
                        sh.hdr[3 * nsw + 2] = nlog;
                    }
                }
                nrot = __shfl_sync(0xffffffffu, nrot, 0);  // also orders lane 0's shared-memory writes before the next tests
                if (nrot > 0)
                {
                    nsw++;
                    nlog += nrot;
                }
            }
            if (lane == 0)
            {
                s_state[0] = more;
                s_state[1] = nsw;
            }
        }
        __syncthreads();
        replay_sweeps(sh.Z, m, sh.lc, sh.ls, sh.hdr, s_state[1]);
        const int more = s_state[0];
        __syncthreads();
        if (!more)
            break;
    }
    const int info = s_state[3];
    for (int t = tid; t < m; t += kDenseBlock)
        sh.ev[t] = sh.d[t] * scale;
    __syncthreads();
    return info;
}

// TridiagQR::compute + matrix_QtHQ on the tridiagonal (d, e) with shift mu (thread 0).
// Rotations go to rc/rs[0..m-1); (d, e) are replaced by Q'TQ.   UpperHessenbergQR.h:515-598, :627-693
__device__ void tridiag_qr_step_scalar(double* d, double* e, int m, double mu, double* rc, double* rs, double* rdiag, double* rsupd)
{
    const int n1 = m - 1, n2 = m - 2;
    // deflation of small sub-diagonal elements (:533-539)
    for (int i = 0; i < n1; i++)
        if (fabs(e[i]) <= kEps * (fabs(d[i]) + fabs(d[i + 1])))
            e[i] = 0.0;
    // R = T - mu I, only what the rotation generation needs (:542-598)
    for (int i = 0; i < m; i++)
        rdiag[i] = d[i] - mu;
    for (int i = 0; i < n1; i++)
        rsupd[i] = e[i];
    double rd = rdiag[0];
    for (int i = 0; i < n1; i++)
    {
        double r, c, s;
        givens_rotation_fast(rd, e[i], r, c, s);
        rc[i] = c;
        rs[i] = s;
        const double Tii1 = rsupd[i];
        const double Ti1i1 = rdiag[i + 1];
        rd = s * Tii1 + c * Ti1i1;  // R[i+1, i+1]
        if (i < n2)
            rsupd[i + 1] *= c;
    }
    // Q'TQ applied to T directly (:627-693).  dest(i+1,i) lives in e[i]; o' needs the original
    // (deflated) T_subd(i+1), which is read before e[i+1] is overwritten.
    for (int i = 0; i < n1; i++)
    {
        const double c = rc[i], s = rs[i];
        const double cs = c * s, c2 = c * c, s2 = s * s;
        const double x = d[i], y = e[i], z = d[i + 1];
        const double c2x = c2 * x, s2x = s2 * x, c2z = c2 * z, s2z = s2 * z;
        const double csy2 = 2.0 * c * s * y;
        d[i] = c2x - csy2 + s2z;
        e[i] = cs * (x - z) + (c2 - s2) * y;
        d[i + 1] = s2x + csy2 + c2z;
        if (i < n2)
        {
            const double ci1 = rc[i + 1], si1 = rs[i + 1];
            const double tsub = e[i + 1];  // m_T_subd[i+1] (still original here)
            const double o = -s * tsub;    // o'
            e[i + 1] = tsub * c;           // w' = dest(i+2,i+1) *= c
            e[i] = ci1 * e[i] - si1 * o;   // y''
        }
    }
    for (int i = 0; i < n1; i++)
        if (fabs(e[i]) <= kEps * (fabs(d[i]) + fabs(d[i + 1])))
            e[i] = 0.0;
}

// The restart applies nshift shifted QR steps one after the other to the same tridiagonal (HermEigsBase.h:124-135).
// A QR step is a forward sweep whose work at position i only touches entries i-2 .. i+2, so consecutive shifts can
// chase each other down the band: lane q of warp 0 runs shift q, kLag positions behind lane q-1 (wavefront
// pipelining).  Every lane performs exactly the operations of TridiagQR::compute (:515-598) + matrix_QtHQ (:627-693)
// in the reference's order, on values that are final for the preceding shift, so the result is bit-identical to
// the sequential chain; the critical path shrinks from nshift*(m-1) to (m-1) + kLag*(nshift-1) rotations.
// Per shift, at local time t:  G(t) rotation t,  Q(t-1) similarity update of rows/cols t-1,t,  D(t-2) deflation test.
constexpr int kLag = 5;
__device__ void qr_chain_round(double* d, double* e, int m, const double* shifts, int nsh, double* lc, double* ls)
{
    const int lane = threadIdx.x;  // warp 0
    const int n1 = m - 1, n2 = m - 2;
    const bool has = lane < nsh;
    const double mu = has ? shifts[lane] : 0.0;
    double rd = 0.0, rsupd = 0.0, ed_cur = 0.0, x = 0.0, y = 0.0, c_prev = 1.0, s_prev = 0.0;
    const int total = (n1 + 2) + kLag * (nsh - 1);
    for (int g = 0; g < total; g++)
    {
        const int t = g - kLag * lane;
        if (has && t >= 0 && t <= n1 + 1)
        {
            if (t == 0)
            {
                const double e0 = e[0];
                ed_cur = (fabs(e0) <= kEps * (fabs(d[0]) + fabs(d[1]))) ? 0.0 : e0;  // pre-deflation (:533-539)
                rd = d[0] - mu;
                rsupd = ed_cur;
                x = d[0];
                y = ed_cur;
            }
            double c_t = 1.0, s_t = 0.0, ed_next = 0.0;
            if (t <= n1 - 1)
            {
                // G(t): rotation from (R[t,t], T[t+1,t])  (:557-590)
                if (t + 1 <= n1 - 1)
                {
                    const double a = e[t + 1];
                    ed_next = (fabs(a) <= kEps * (fabs(d[t + 1]) + fabs(d[t + 2]))) ? 0.0 : a;
                }
                double r;
                givens_rotation_fast(rd, ed_cur, r, c_t, s_t);
                lc[lane * n1 + t] = c_t;
                ls[lane * n1 + t] = s_t;
                rd = s_t * rsupd + c_t * (d[t + 1] - mu);  // R[t+1, t+1]
                rsupd = c_t * ed_next;                     // R[t+1, t+2]
            }
            if (t >= 1 && t <= n1)
            {
                // Q(i): Gi' T Gi on rows/cols i, i+1  (:650-679); (c_t, s_t) is rotation i+1
                const int i = t - 1;
                const double c = c_prev, s = s_prev;
                const double z = d[i + 1];
                const double cs = c * s, c2 = c * c, s2 = s * s;
                const double c2x = c2 * x, s2x = s2 * x, c2z = c2 * z, s2z = s2 * z;
                const double csy2 = 2.0 * c * s * y;
                const double xn = c2x - csy2 + s2z;
                double yn = cs * (x - z) + (c2 - s2) * y;
                const double zn = s2x + csy2 + c2z;
                double wn = 0.0;
                if (i < n2)
                {
                    const double tsub = ed_cur;  // deflated T[i+2, i+1]
                    const double o = -s * tsub;
                    wn = tsub * c;
                    yn = c_t * yn - s_t * o;
                }
                d[i] = xn;
                e[i] = yn;
                if (i == n1 - 1)
                    d[i + 1] = zn;
                x = zn;
                y = wn;
            }
            if (t >= 2)
            {
                // D(i): post-deflation (:682-689); d[i] and d[i+1] are final by now
                const int i = t - 2;
                if (fabs(e[i]) <= kEps * (fabs(d[i]) + fabs(d[i + 1])))
                    e[i] = 0.0;
            }
            c_prev = c_t;
            s_prev = s_t;
            ed_cur = ed_next;
        }
        __syncwarp();
    }
}

// Y <- Y Q for nseq consecutive rotation sequences of m-1 rotations each (UpperHessenbergQR.h:383-417)
__device__ void apply_yq_block(double* Y, int m, const double* lc, const double* ls, int nseq)
{
    for (int t = threadIdx.x; t < m; t += kDenseBlock)
    {
        for (int q = 0; q < nseq; q++)
        {
            const double* rc = lc + q * (m - 1);
            const double* rs = ls + q * (m - 1);
            double yi = Y[t];
            for (int i = 0; i < m - 1; i++)
            {
                const double c = rc[i], s = rs[i];
                const double yi1 = Y[t + (i + 1) * m];
                Y[t + i * m] = c * yi - s * yi1;
                yi = s * yi + c * yi1;
            }
            Y[t + (m - 1) * m] = yi;
        }
    }
}

__global__ void __launch_bounds__(kDenseBlock)
    sym_restart_kernel(double* H, int m, int nev, const FacCtl* ctl, double beta_override, int use_override, int selection, double tol, double* ritz_val,
                       double* ritz_est, double* ritz_vec, int* ritz_conv, double* Q, SymRestartOut* out, int do_restart)
{
    extern __shared__ double smem[];
    TriShared sh = carve(smem, m);
    __shared__ double s_hd[kMaxNcv], s_he[kMaxNcv];
    __shared__ int s_k, s_go;
    const int tid = threadIdx.x;
    for (int t = tid; t < m; t += kDenseBlock)
    {
        s_hd[t] = H[t + (int64_t) t * m];
        if (t < m - 1)
            s_he[t] = H[(t + 1) + (int64_t) t * m];
    }
    __syncthreads();

    // ---- retrieve_ritzpair (HermEigsBase.h:205-224) ----
    const int info = tridiag_eigen_block(s_hd, s_he, m, sh);
    if (tid == 0)
    {
        for (int i = 0; i < m; i++)
            sh.aux[i] = sort_key_real(selection, sh.ev[i]);
        argsort_keys(sh.aux, sh.idx, m);
        if (selection == SB200_BOTH_ENDS)  // SelectionRule.h:272-284
        {
            int* tmp = reinterpret_cast<int*>(sh.t1);
            for (int i = 0; i < m; i++)
                tmp[i] = sh.idx[i];
            for (int i = 0; i < m; i++)
                sh.idx[i] = (i % 2 == 0) ? tmp[i / 2] : tmp[m - 1 - i / 2];
        }
    }
    __syncthreads();
    for (int t = tid; t < m; t += kDenseBlock)
    {
        const int id = sh.idx[t];
        const double rv = sh.ev[id];
        const double re = sh.Z[(m - 1) + id * m];
        ritz_val[t] = rv;
        ritz_est[t] = re;
        sh.d[t] = rv;  // keep sorted Ritz values / estimates for the scalar logic below
        sh.e[t] = re;
    }
    for (int t = tid; t < m * nev; t += kDenseBlock)
    {
        const int r = t % m, c = t / m;
        ritz_vec[t] = sh.Z[r + sh.idx[c] * m];
    }
    __syncthreads();

    if (tid == 0)
    {
        // ---- num_converged (HermEigsBase.h:158-175) ----
        const double beta = use_override ? beta_override : ctl->beta;
        const double eps23 = 3.666852862501036e-11;  // eps^(2/3)
        int nconv = 0;
        for (int i = 0; i < nev; i++)
        {
            const double thresh = tol * fmax(fabs(sh.d[i]), eps23);
            const double resid = fabs(sh.e[i]) * beta;
            const int cv = resid < thresh;
            ritz_conv[i] = cv;
            nconv += cv;
        }
        // ---- nev_adjusted (HermEigsBase.h:178-202) ----
        int nev_new = nev;
        for (int i = nev; i < m; i++)
            if (fabs(sh.e[i]) < kNear0)
                nev_new++;
        nev_new += min(nconv, (m - nev_new) / 2);
        if (nev_new == 1 && m >= 6)
            nev_new = m / 2;
        else if (nev_new == 1 && m > 2)
            nev_new = 2;
        if (nev_new > m - 1)
            nev_new = m - 1;
        s_k = nev_new;
        s_go = (do_restart && nconv < nev && info == 0 && nev_new < m) ? 1 : 0;
        out->nconv = nconv;
        out->k = nev_new;
        out->info = info;
    }
    __syncthreads();
    if (!s_go)
        return;

    // ---- shift loop of restart() (HermEigsBase.h:112-147) ----
    const int k = s_k;
    const int nshift = m - k;
    double* shifts = sh.ev;  // reuse
    if (tid == 0)
    {
        // shifts = ritz_val.tail(nshift) sorted by |.| descending (:118-121)
        for (int i = 0; i < nshift; i++)
            sh.aux[i] = -fabs(sh.d[k + i]);
        argsort_keys(sh.aux, sh.idx, nshift);
        for (int i = 0; i < nshift; i++)
            shifts[i] = sh.d[k + sh.idx[i]];
    }
    // Q = I in the (now free) Z buffer; the working tridiagonal is (s_hd, s_he)
    for (int t = tid; t < m * m; t += kDenseBlock)
        sh.Z[t] = ((t % m) == (t / m)) ? 1.0 : 0.0;
    __syncthreads();
    for (int ish0 = 0; ish0 < nshift; ish0 += 32)
    {
        const int cnt = min(32, nshift - ish0);  // one warp-wide wavefront of shifts; 32 * (m - 1) <= kLogCap
        if (tid < 32)
            qr_chain_round(s_hd, s_he, m, shifts + ish0, cnt, sh.lc, sh.ls);
        __syncthreads();
        apply_yq_block(sh.Z, m, sh.lc, sh.ls, cnt);
        __syncthreads();
    }
    // write back Q and the (untrimmed) tridiagonal H
    for (int t = tid; t < m * m; t += kDenseBlock)
    {
        Q[t] = sh.Z[t];
        const int r = t % m, c = t / m;
        double h = 0.0;
        if (r == c)
            h = s_hd[r];
        else if (r == c + 1)
            h = s_he[c];
        else if (c == r + 1)
            h = s_he[r];
        H[t] = h;
    }
}

__global__ void __launch_bounds__(kDenseBlock) tridiag_eigen_kernel(const double* H, int m, double* evals, double* evecs, int* info)
{
    extern __shared__ double smem[];
    TriShared sh = carve(smem, m);
    __shared__ double s_hd[kMaxNcv], s_he[kMaxNcv];
    for (int t = threadIdx.x; t < m; t += kDenseBlock)
    {
        s_hd[t] = H[t + (int64_t) t * m];
        if (t < m - 1)
            s_he[t] = H[(t + 1) + (int64_t) t * m];
    }
    __syncthreads();
    const int rc = tridiag_eigen_block(s_hd, s_he, m, sh);
    for (int t = threadIdx.x; t < m; t += kDenseBlock)
        evals[t] = sh.ev[t];
    for (int t = threadIdx.x; t < m * m; t += kDenseBlock)
        evecs[t] = sh.Z[t];
    if (threadIdx.x == 0)
        *info = rc;
}

__global__ void __launch_bounds__(kDenseBlock) tridiag_qr_kernel(const double* H, int m, double shift, double* QtHQ, double* Q)
{
    extern __shared__ double smem[];
    TriShared sh = carve(smem, m);
    __shared__ double s_hd[kMaxNcv], s_he[kMaxNcv];
    for (int t = threadIdx.x; t < m; t += kDenseBlock)
    {
        s_hd[t] = H[t + (int64_t) t * m];
        if (t < m - 1)
            s_he[t] = H[(t + 1) + (int64_t) t * m];
    }
    for (int t = threadIdx.x; t < m * m; t += kDenseBlock)
        sh.Z[t] = ((t % m) == (t / m)) ? 1.0 : 0.0;
    __syncthreads();
    if (threadIdx.x == 0)
        tridiag_qr_step_scalar(s_hd, s_he, m, shift, sh.lc, sh.ls, sh.d, sh.e);
    __syncthreads();
    apply_yq_block(sh.Z, m, sh.lc, sh.ls, 1);
    __syncthreads();
    for (int t = threadIdx.x; t < m * m; t += kDenseBlock)
    {
        Q[t] = sh.Z[t];
        const int r = t % m, c = t / m;
        double h = 0.0;
        if (r == c)
            h = s_hd[r];
        else if (r == c + 1)
            h = s_he[c];
        else if (c == r + 1)
            h = s_he[r];
        QtHQ[t] = h;
    }
}

void ensure_smem(const void* fn, size_t bytes)
{
    if (bytes > 48 * 1024)
        SB200_CUDA_CHECK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) bytes));
}

// Unit-tier hook for Givens<double>::compute_rotation (Givens.h:166-205; StableScaling :28-86): one rotation per thread.
// variant 0 = givens_rotation (the reference's formulas, Taylor branch included), 1 = givens_rotation_fast (the form the QR
// kernels of this file call: rsqrt + fallback to variant 0), 2 = make_givens (Eigen's JacobiRotation::makeGivens as used by
// TridiagEigen.h:79-80, r returned through its third argument).
__global__ void givens_batch_kernel(int variant, int64_t count, const double* x, const double* y, double* r, double* c, double* s)
{
    const int64_t t = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count)
        return;
    double rr = 0.0, cc = 0.0, ss = 0.0;
    if (variant == 0)
        givens_rotation(x[t], y[t], rr, cc, ss);
    else if (variant == 1)
        givens_rotation_fast(x[t], y[t], rr, cc, ss);
    else
        make_givens(x[t], y[t], cc, ss, &rr);
    r[t] = rr;
    c[t] = cc;
    s[t] = ss;
}

}  // namespace

void launch_givens_batch(int variant, int64_t count, const double* x, const double* y, double* r, double* c, double* s, cudaStream_t stream)
{
    if (count <= 0)
        return;
    givens_batch_kernel<<<(unsigned) ((count + 127) / 128), 128, 0, stream>>>(variant, count, x, y, r, c, s);
    SB200_CUDA_CHECK(cudaGetLastError());
}

void launch_sym_restart(double* H, int m, int nev, const FacCtl* ctl, int selection, double tol, double* ritz_val, double* ritz_est, double* ritz_vec,
                        int* ritz_conv, double* Q, SymRestartOut* out, int do_restart, cudaStream_t stream)
{
    SB200_REQUIRE(m >= 2 && m <= kMaxNcv, SB200_INVALID_ARGUMENT, "ncv out of range for the device restart kernel");
    const size_t smem = tri_smem_bytes(m);
    ensure_smem((const void*) sym_restart_kernel, smem);
    sym_restart_kernel<<<1, kDenseBlock, smem, stream>>>(H, m, nev, ctl, 0.0, 0, selection, tol, ritz_val, ritz_est, ritz_vec, ritz_conv, Q, out, do_restart);
    SB200_CUDA_CHECK(cudaGetLastError());
}

// test hook: beta passed by value instead of through a control block
void launch_sym_restart_beta(double* H, int m, int nev, double beta, int selection, double tol, double* ritz_val, double* ritz_est, double* ritz_vec,
                             int* ritz_conv, double* Q, SymRestartOut* out, cudaStream_t stream)
{
    SB200_REQUIRE(m >= 2 && m <= kMaxNcv, SB200_INVALID_ARGUMENT, "ncv out of range for the device restart kernel");
    const size_t smem = tri_smem_bytes(m);
    ensure_smem((const void*) sym_restart_kernel, smem);
    sym_restart_kernel<<<1, kDenseBlock, smem, stream>>>(H, m, nev, nullptr, beta, 1, selection, tol, ritz_val, ritz_est, ritz_vec, ritz_conv, Q, out, 1);
    SB200_CUDA_CHECK(cudaGetLastError());
}

void launch_tridiag_eigen(const double* H, int m, double* evals, double* evecs, int* info, cudaStream_t stream)
{
    SB200_REQUIRE(m >= 1 && m <= kMaxNcv, SB200_INVALID_ARGUMENT, "matrix order out of range");
    const size_t smem = tri_smem_bytes(m);
    ensure_smem((const void*) tridiag_eigen_kernel, smem);
    tridiag_eigen_kernel<<<1, kDenseBlock, smem, stream>>>(H, m, evals, evecs, info);
    SB200_CUDA_CHECK(cudaGetLastError());
}

void launch_tridiag_qr(const double* H, int m, double shift, double* QtHQ, double* Q, cudaStream_t stream)
{
    SB200_REQUIRE(m >= 2 && m <= kMaxNcv, SB200_INVALID_ARGUMENT, "matrix order out of range");
    const size_t smem = tri_smem_bytes(m);
    ensure_smem((const void*) tridiag_qr_kernel, smem);
    tridiag_qr_kernel<<<1, kDenseBlock, smem, stream>>>(H, m, shift, QtHQ, Q);
    SB200_CUDA_CHECK(cudaGetLastError());
}

}  // namespace sb200

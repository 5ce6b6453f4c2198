// Small dense device kernels of the COMPLEX GenEigsSolver: one warp, the m x m complex matrices in shared memory.
//
// Replaces, for Scalar = std::complex<double>:
//   Givens<complex>::compute_rotation          LinAlg/Givens.h:218-335 (+ StableScaling :64-92)
//   UpperHessenbergQR<complex>                 LinAlg/UpperHessenbergQR.h:136-195 (compute), :219-255 (matrix_QtHQ), :383-417 (apply_YQ)
//   UpperHessenbergEigen<complex>              LinAlg/UpperHessenbergEigen.h:328-454: Eigen::ComplexSchur from a Hessenberg matrix
//                                              (single-shift QR with the shift strategy of ComplexSchur::computeShift, deflation test of
//                                              subdiagonalEntryIsNeglegible; restated from Eigen 3.4's published algorithm), unit-triangular
//                                              back-substitution :347-379, V = U X, column normalisation :383-387
//   GenEigsBase<complex>                       retrieve_ritzpair :280-340, num_converged :225-242, nev_adjusted :245-277,
//                                              RestartArnoldi<complex>::run :122-139
// The Schur sweep uses the reference's own complex Givens instead of Eigen's JacobiRotation::makeGivens: any rotation that annihilates
// the sub-diagonal entry gives a valid Schur form; the Ritz values are the same and the Ritz vectors differ by a unit phase, which
// neither the convergence test (|last component|) nor V s notices.
//
// Thread model: 32 lanes.  Scalars are recomputed by every lane from shared memory, row / column updates are split over the lanes, and
// every read-then-write of the same location is separated by __syncwarp() (checked in both fiber orders on the emulator).
#include "kernels.h"
#include "dense_common.cuh"

namespace sb200 {

struct GenRestartOut
{
    int nconv;
    int k;
    int info;  // 0 ok, 1 Schur iteration cap hit
    int pad;
};

namespace {

using namespace dense;

constexpr int kZBlock = 32;
#define ZLANE ((int) threadIdx.x)
#define ZFOR(var, lo, hi) for (int var = (lo) + ZLANE; var < (hi); var += kZBlock)
#define ZM(A, i, j) (A)[(i) + (j) * m]

struct Z
{
    double re, im;
};
__device__ __forceinline__ Z zmake(double a, double b)
{
    Z r;
    r.re = a;
    r.im = b;
    return r;
}
__device__ __forceinline__ Z zadd(Z a, Z b) { return zmake(a.re + b.re, a.im + b.im); }
__device__ __forceinline__ Z zsub(Z a, Z b) { return zmake(a.re - b.re, a.im - b.im); }
__device__ __forceinline__ Z zmul(Z a, Z b) { return zmake(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
__device__ __forceinline__ Z zconj(Z a) { return zmake(a.re, -a.im); }
__device__ __forceinline__ Z zscale(double c, Z a) { return zmake(c * a.re, c * a.im); }
__device__ __forceinline__ double zabs2(Z a) { return a.re * a.re + a.im * a.im; }
__device__ __forceinline__ double zabs(Z a) { return hypot(a.re, a.im); }
__device__ __forceinline__ double znorm1(Z a) { return fabs(a.re) + fabs(a.im); }
__device__ __forceinline__ bool zis0(Z a) { return a.re == 0.0 && a.im == 0.0; }
// Smith's algorithm (the scaling of libgcc's __divdc3 without inf/nan recovery)
__device__ __forceinline__ Z zdiv(Z a, Z b)
{
    Z r;
    if (fabs(b.re) < fabs(b.im))
    {
        const double ratio = b.re / b.im, denom = b.re * ratio + b.im;
        r.re = (a.re * ratio + a.im) / denom;
        r.im = (a.im * ratio - a.re) / denom;
    }
    else
    {
        const double ratio = b.im / b.re, denom = b.im * ratio + b.re;
        r.re = (a.im * ratio + a.re) / denom;
        r.im = (a.im - a.re * ratio) / denom;
    }
    return r;
}
// principal square root
__device__ __forceinline__ Z zsqrt(Z a)
{
    const double r = hypot(a.re, a.im);
    if (r == 0.0)
        return zmake(0.0, 0.0);
    const double t = sqrt(0.5 * (r + fabs(a.re)));
    if (a.re >= 0.0)
        return zmake(t, a.im / (2.0 * t));
    return zmake(fabs(a.im) / (2.0 * t), a.im >= 0.0 ? t : -t);
}

// Givens<complex>::compute_rotation (Givens.h:218-335): real c, complex s, r with  c x - s y = r,  conj(s) x + c y = 0
__device__ void zgivens(Z x, Z y, Z& r, double& c, Z& s)
{
    if (zis0(y))
    {
        c = 1.0;
        s = zmake(0.0, 0.0);
        r = x;
        return;
    }
    if (zis0(x))
    {
        // equivalent to the real rotation of (-Re y, -Im y)   (:231-243)
        double rr, sr, si;
        givens_rotation(-y.re, -y.im, rr, sr, si);
        c = 0.0;
        s = zmake(sr, si);
        r = zmake(rr, 0.0);
        return;
    }
    if (znorm1(x) > znorm1(y))
    {
        // StableScaling::run(Complex, Complex, a2, tc1, tc2)   (:64-92)
        const double b2 = zabs2(y), a2 = zabs2(x), t2 = b2 / a2;
        double tc1, tc2;
        if (t2 >= 0.1 * 1.4901161193847656e-8)  // 0.1 * sqrt(eps), sqrt(eps) = 2^-26
        {
            tc1 = sqrt(1.0 + t2);
            tc2 = sqrt(a2 / (a2 + b2));
        }
        else
        {
            tc1 = 1.0 + t2 * (0.5 - t2 * (0.125 - 0.0625 * t2));
            tc2 = 1.0 - t2 * (0.5 - t2 * (0.375 - 0.3125 * t2));
        }
        c = tc2;
        r = zscale(tc1, x);
        s = zscale(-(c / a2), zmul(x, zconj(y)));
        return;
    }
    const double rho = sqrt(zabs2(x) + zabs2(y));
    double xnorm, zr, zi;
    givens_rotation(x.re, -x.im, xnorm, zr, zi);
    const Z z = zmake(zr, zi);
    r = zscale(rho, z);
    c = xnorm / rho;
    s = zscale(-1.0 / rho, zmul(z, zconj(y)));
}

// rows p, p+1 of A, columns [c0, m):  [x; y] <- [c x - s y;  conj(s) x + c y]
__device__ void zrot_left(Z* A, int m, int p, int c0, double c, Z s)
{
    ZFOR(j, c0, m)
    {
        const Z x = ZM(A, p, j), y = ZM(A, p + 1, j);
        ZM(A, p, j) = zsub(zscale(c, x), zmul(s, y));
        ZM(A, p + 1, j) = zadd(zmul(zconj(s), x), zscale(c, y));
    }
}
// columns p, p+1 of A, rows [0, nrow):  [x y] <- [c x - conj(s) y,  s x + c y]      (apply_YQ, UpperHessenbergQR.h:406-413)
__device__ void zrot_right(Z* A, int m, int p, int nrow, double c, Z s)
{
    ZFOR(i, 0, nrow)
    {
        const Z x = ZM(A, i, p), y = ZM(A, i, p + 1);
        ZM(A, i, p) = zsub(zscale(c, x), zmul(zconj(s), y));
        ZM(A, i, p + 1) = zadd(zmul(s, x), zscale(c, y));
    }
}

// UpperHessenbergQR<complex> on H in place: H <- R Q + mu I (= Q^H H Q), Qacc <- Qacc Q.   rc: m doubles, rs: m complex.
__device__ void zhess_qr_shift(Z* H, Z* Qacc, int m, Z mu, double* rc, Z* rs)
{
    ZFOR(i, 0, m) ZM(H, i, i) = zsub(ZM(H, i, i), mu);
    __syncwarp();
    for (int i = 0; i < m - 1; i++)
    {
        // make sure R is upper Hessenberg: zero below R(i + 1, i)   (:163)
        ZFOR(r, i + 2, m) ZM(H, r, i) = zmake(0.0, 0.0);
        const Z xi = ZM(H, i, i), xj = ZM(H, i + 1, i);
        Z r, s;
        double c;
        zgivens(xi, xj, r, c, s);
        __syncwarp();
        if (ZLANE == 0)
        {
            rc[i] = c;
            rs[i] = s;
            ZM(H, i, i) = r;
            ZM(H, i + 1, i) = zmake(0.0, 0.0);
        }
        zrot_left(H, m, i, i + 1, c, s);  // columns i + 1 .. m - 1   (:178-184)
        __syncwarp();
    }
    // RQ (:231-249) and Q accumulation (:383-417)
    for (int i = 0; i < m - 1; i++)
    {
        const double c = rc[i];
        const Z s = rs[i];
        zrot_right(H, m, i, i + 2, c, s);
        zrot_right(Qacc, m, i, m, c, s);
        __syncwarp();
    }
    ZFOR(i, 0, m) ZM(H, i, i) = zadd(ZM(H, i, i), mu);
    __syncwarp();
}

// ComplexSchur::subdiagonalEntryIsNeglegible: |sd|_1 <= eps (|T(i,i)|_1 + |T(i+1,i+1)|_1) -> T(i+1, i) := 0
__device__ bool zsub_negligible(Z* T, int m, int i)
{
    const double d = znorm1(ZM(T, i, i)) + znorm1(ZM(T, i + 1, i + 1));
    const double sd = znorm1(ZM(T, i + 1, i));
    const bool neg = sd <= d * kEps;
    __syncwarp();
    if (neg && ZLANE == 0)
        ZM(T, i + 1, i) = zmake(0.0, 0.0);
    __syncwarp();
    return neg;
}

// ComplexSchur::computeShift
__device__ Z zschur_shift(const Z* T, int m, int iu, int iter)
{
    if (iter == 10 || iter == 20)
    {
        // exceptional shift (EISPACK comqr)
        const double a = fabs(ZM(T, iu, iu - 1).re), b = (iu >= 2) ? fabs(ZM(T, iu - 1, iu - 2).re) : 0.0;
        return zmake(a + b, 0.0);
    }
    Z t00 = ZM(T, iu - 1, iu - 1), t01 = ZM(T, iu - 1, iu), t10 = ZM(T, iu, iu - 1), t11 = ZM(T, iu, iu);
    const double normt = zabs(t00) + zabs(t01) + zabs(t10) + zabs(t11);
    t00 = zscale(1.0 / normt, t00);
    t01 = zscale(1.0 / normt, t01);
    t10 = zscale(1.0 / normt, t10);
    t11 = zscale(1.0 / normt, t11);
    const Z b = zmul(t01, t10);
    const Z c = zsub(t00, t11);
    const Z disc = zsqrt(zadd(zmul(c, c), zscale(4.0, b)));
    const Z det = zsub(zmul(t00, t11), b);
    const Z trace = zadd(t00, t11);
    Z e1 = zscale(0.5, zadd(trace, disc)), e2 = zscale(0.5, zsub(trace, disc));
    const double n1 = znorm1(e1), n2 = znorm1(e2);
    if (n1 > n2)
        e2 = zdiv(det, e1);
    else if (n2 != 0.0)
        e1 = zdiv(det, e2);
    if (znorm1(zsub(e1, t11)) < znorm1(zsub(e2, t11)))
        return zscale(normt, e1);
    return zscale(normt, e2);
}

// ComplexSchur::reduceToTriangularForm on the Hessenberg matrix T (in place), U <- Schur vectors.  Returns 0 / 1 (iteration cap).
__device__ int zschur(Z* T, Z* U, int m)
{
    ZFOR(q, 0, m * m) U[q] = zmake(((q % m) == (q / m)) ? 1.0 : 0.0, 0.0);
    __syncwarp();
    const int max_iter = 30 * m;
    int iu = m - 1, iter = 0, total = 0;
    while (true)
    {
        while (iu > 0)
        {
            if (!zsub_negligible(T, m, iu - 1))
                break;
            iter = 0;
            --iu;
        }
        if (iu == 0)
            break;
        iter++;
        total++;
        if (total > max_iter)
            break;
        int il = iu - 1;
        while (il > 0 && !zsub_negligible(T, m, il - 1))
            --il;
        const Z shift = zschur_shift(T, m, iu, iter);
        Z r, s;
        double c;
        zgivens(zsub(ZM(T, il, il), shift), ZM(T, il + 1, il), r, c, s);
        __syncwarp();
        zrot_left(T, m, il, max(il - 1, 0), c, s);
        __syncwarp();
        zrot_right(T, m, il, min(il + 2, iu) + 1, c, s);
        zrot_right(U, m, il, m, c, s);
        __syncwarp();
        for (int i = il + 1; i < iu; i++)
        {
            zgivens(ZM(T, i, i - 1), ZM(T, i + 1, i - 1), r, c, s);
            __syncwarp();
            if (ZLANE == 0)
            {
                ZM(T, i, i - 1) = r;
                ZM(T, i + 1, i - 1) = zmake(0.0, 0.0);
            }
            zrot_left(T, m, i, i, c, s);
            __syncwarp();
            zrot_right(T, m, i, min(i + 2, iu) + 1, c, s);
            zrot_right(U, m, i, m, c, s);
            __syncwarp();
        }
    }
    return (total > max_iter) ? 1 : 0;
}

// UpperHessenbergEigen<complex>::doComputeEigenvectors (:347-387): ev <- diag(T); X (unit upper triangular, in place of T) with
// T X = X D; U <- U X, columns normalised.  tmp: m complex.
__device__ void zeigenvectors(Z* T, Z* U, int m, Z* ev, Z* tmp)
{
    // Frobenius norm of T (m_schur.matrixT().norm())
    double nrm = 0.0;
    for (int q = 0; q < m * m; q++)
        nrm += zabs2(T[q]);
    nrm = fmax(sqrt(nrm), kMin);
    ZFOR(i, 0, m) ev[i] = ZM(T, i, i);
    __syncwarp();
    for (int k = m - 1; k >= 0; k--)
    {
        // column k of X overwrites column k of T; columns < k and the diagonal above row k are still T's
        for (int i = k - 1; i >= 0; i--)
        {
            // X(i,k) = -(T(i,k) + sum_{j=i+1}^{k-1} T(i,j) X(j,k)) / (T(i,i) - T(k,k))
            double sr = 0.0, si = 0.0;
            ZFOR(j, i + 1, k)
            {
                const Z p = zmul(ZM(T, i, j), ZM(T, j, k));
                sr += p.re;
                si += p.im;
            }
            sr = warp_sum(sr);
            si = warp_sum(si);
            Z z = zsub(ev[i], ev[k]);
            if (zis0(z))
                z.re = kEps * nrm;
            const Z num = zmake(-(ZM(T, i, k).re + sr), -(ZM(T, i, k).im + si));
            const Z x = zdiv(num, z);
            __syncwarp();
            if (ZLANE == 0)
                ZM(T, i, k) = x;
            __syncwarp();
        }
        if (ZLANE == 0)
            ZM(T, k, k) = zmake(1.0, 0.0);
        __syncwarp();
    }
    // V = U X, column j descending so that U(:, 0..j) is still the Schur basis when column j is formed
    for (int j = m - 1; j >= 0; j--)
    {
        ZFOR(i, 0, m)
        {
            Z acc = zmake(0.0, 0.0);
            for (int k = 0; k <= j; k++)
                acc = zadd(acc, zmul(ZM(U, i, k), ZM(T, k, j)));
            tmp[i] = acc;
        }
        __syncwarp();
        double sq = 0.0;
        ZFOR(i, 0, m) sq += zabs2(tmp[i]);
        sq = warp_sum(sq);
        const double inv = sq > 0.0 ? 1.0 / sqrt(sq) : 1.0;
        ZFOR(i, 0, m) ZM(U, i, j) = zscale(inv, tmp[i]);
        __syncwarp();
    }
}

// SortingTarget<complex, Rule>::get (SelectionRule.h:68-192)
__device__ __forceinline__ double zsort_key(int rule, Z v)
{
    switch (rule)
    {
        case SB200_LARGEST_MAGN: return -zabs(v);
        case SB200_LARGEST_REAL: return -v.re;
        case SB200_LARGEST_IMAG: return -fabs(v.im);
        case SB200_SMALLEST_MAGN: return zabs(v);
        case SB200_SMALLEST_REAL: return v.re;
        default: return fabs(v.im);  // SB200_SMALLEST_IMAG
    }
}

struct ZShared
{
    Z *H, *T, *U;        // m x m each
    Z *ev, *tmp, *rs;    // m each
    Z *rv;               // sorted Ritz values
    double *key, *rc, *re_abs;
    int* idx;
};
__device__ ZShared zcarve(double* smem, int m)
{
    ZShared g;
    Z* p = reinterpret_cast<Z*>(smem);
    g.H = p;
    p += m * m;
    g.T = p;
    p += m * m;
    g.U = p;
    p += m * m;
    g.ev = p;
    p += m;
    g.tmp = p;
    p += m;
    g.rs = p;
    p += m;
    g.rv = p;
    p += m;
    double* d = reinterpret_cast<double*>(p);
    g.key = d;
    d += m;
    g.rc = d;
    d += m;
    g.re_abs = d;
    d += m;
    g.idx = reinterpret_cast<int*>(d);
    return g;
}
size_t zsmem_bytes(int m) { return sizeof(double) * (size_t) (6 * m * m + 8 * m + 3 * m) + sizeof(int) * (size_t) (m + 2); }

// Ritz pairs of H (in g.H): g.ev (unsorted eigenvalues, ascending modulus as UpperHessenbergEigen::sortEigenvalues leaves them is not
// needed -- the selection sort below is total), g.U eigenvectors.  Returns info.
__device__ int zeigen_block(ZShared& g, int m)
{
    ZFOR(q, 0, m * m) g.T[q] = g.H[q];
    __syncwarp();
    const int info = zschur(g.T, g.U, m);
    __syncwarp();
    zeigenvectors(g.T, g.U, m, g.ev, g.tmp);
    __syncwarp();
    return info;
}

__global__ void __launch_bounds__(kZBlock)
    gen_restart_z_kernel(double* Hr, double* Hi, int m, int nev, const FacCtl* ctl, double beta_override, int use_override, int selection, double tol,
                         double* ritz_val_ri, double* ritz_est_ri, double* ritz_vec_ri, int* ritz_conv, double* Qr, double* Qi, GenRestartOut* out, int do_restart)
{
    extern __shared__ double smem[];
    ZShared g = zcarve(smem, m);
    ZFOR(q, 0, m * m) g.H[q] = zmake(Hr[q], Hi[q]);
    __syncwarp();
    const int info = zeigen_block(g, m);
    // ---- retrieve_ritzpair (:280-340): UpperHessenbergEigen::sortEigenvalues orders by ascending modulus first (:389-404) ----
    if (ZLANE == 0)
    {
        for (int i = 0; i < m; i++)
            g.key[i] = zabs(g.ev[i]);
        argsort_keys(g.key, g.idx, m);
        // second key: the selection rule, stable with respect to the modulus order
        for (int i = 0; i < m; i++)
            g.re_abs[i] = zsort_key(selection, g.ev[g.idx[i]]);
        // stable insertion sort of idx by re_abs (keys indexed by position)
        for (int i = 1; i < m; i++)
        {
            const int id = g.idx[i];
            const double k = g.re_abs[i];
            int q = i - 1;
            while (q >= 0 && g.re_abs[q] > k)
            {
                g.idx[q + 1] = g.idx[q];
                g.re_abs[q + 1] = g.re_abs[q];
                q--;
            }
            g.idx[q + 1] = id;
            g.re_abs[q + 1] = k;
        }
    }
    __syncwarp();
    ZFOR(t, 0, m)
    {
        const int id = g.idx[t];
        const Z v = g.ev[id];
        ritz_val_ri[2 * t] = v.re;
        ritz_val_ri[2 * t + 1] = v.im;
        g.rv[t] = v;
        const Z e = ZM(g.U, m - 1, id);
        ritz_est_ri[2 * t] = e.re;
        ritz_est_ri[2 * t + 1] = e.im;
        g.key[t] = zabs(e);
    }
    for (int c = 0; c < nev; c++)
    {
        const int id = g.idx[c];
        ZFOR(r, 0, m)
        {
            const Z e = ZM(g.U, r, id);
            ritz_vec_ri[2 * (r + c * m)] = e.re;
            ritz_vec_ri[2 * (r + c * m) + 1] = e.im;
        }
    }
    __syncwarp();
    // ---- num_converged (:225-242), nev_adjusted (:245-277) ----
    const double beta = use_override ? beta_override : ctl->beta;
    const double eps23 = 3.666852862501036e-11;
    int nconv = 0;
    for (int i = 0; i < nev; i++)
    {
        const double thresh = tol * fmax(zabs(g.rv[i]), eps23);
        const int cv = (g.key[i] * beta) < thresh;
        if (ZLANE == 0)
            ritz_conv[i] = cv;
        nconv += cv;
    }
    int nev_new = nev;
    for (int i = nev; i < m; i++)
        if (g.key[i] < kNear0)
            nev_new++;
    nev_new += min(nconv, (m - nev_new) / 2);
    if (nev_new == 1 && m >= 6)
        nev_new = m / 2;
    else if (nev_new == 1 && m > 3)
        nev_new = 2;
    if (nev_new > m - 2)
        nev_new = m - 2;
    if (nev_new >= 1 && g.rv[nev_new - 1].im != 0.0 && g.rv[nev_new - 1].re == g.rv[nev_new].re && g.rv[nev_new - 1].im == -g.rv[nev_new].im)
        nev_new++;
    if (ZLANE == 0)
    {
        out->nconv = nconv;
        out->k = nev_new;
        out->info = info;
    }
    if (!(do_restart && nconv < nev && info == 0 && nev_new < m))
        return;

    // ---- RestartArnoldi<complex>::run (:122-139): one complex shift per unwanted Ritz value; Q accumulates in the T buffer ----
    Z* Qs = g.T;
    __syncwarp();
    ZFOR(q, 0, m * m) Qs[q] = zmake(((q % m) == (q / m)) ? 1.0 : 0.0, 0.0);
    __syncwarp();
    for (int i = nev_new; i < m; i++)
        zhess_qr_shift(g.H, Qs, m, g.rv[i], g.rc, g.rs);
    __syncwarp();
    ZFOR(q, 0, m * m)
    {
        Hr[q] = g.H[q].re;
        Hi[q] = g.H[q].im;
        Qr[q] = Qs[q].re;
        Qi[q] = Qs[q].im;
    }
}

// ---- standalone kernels for the unit tier (test/QR.cpp:177-189, test/Eigen.cpp with a complex matrix) ----
__global__ void __launch_bounds__(kZBlock) hess_qr_z_kernel(const double* H_ri, int m, double mu_re, double mu_im, double* QtHQ_ri, double* Q_ri)
{
    extern __shared__ double smem[];
    ZShared g = zcarve(smem, m);
    ZFOR(q, 0, m * m)
    {
        g.H[q] = zmake(H_ri[2 * q], H_ri[2 * q + 1]);
        g.T[q] = zmake(((q % m) == (q / m)) ? 1.0 : 0.0, 0.0);
    }
    __syncwarp();
    zhess_qr_shift(g.H, g.T, m, zmake(mu_re, mu_im), g.rc, g.rs);
    ZFOR(q, 0, m * m)
    {
        QtHQ_ri[2 * q] = g.H[q].re;
        QtHQ_ri[2 * q + 1] = g.H[q].im;
        Q_ri[2 * q] = g.T[q].re;
        Q_ri[2 * q + 1] = g.T[q].im;
    }
}

__global__ void __launch_bounds__(kZBlock) hess_eigen_z_kernel(const double* H_ri, int m, double* evals_ri, double* evecs_ri, int* info)
{
    extern __shared__ double smem[];
    ZShared g = zcarve(smem, m);
    ZFOR(q, 0, m * m) g.H[q] = zmake(H_ri[2 * q], H_ri[2 * q + 1]);
    __syncwarp();
    const int rc = zeigen_block(g, m);
    ZFOR(j, 0, m)
    {
        evals_ri[2 * j] = g.ev[j].re;
        evals_ri[2 * j + 1] = g.ev[j].im;
    }
    ZFOR(q, 0, m * m)
    {
        evecs_ri[2 * q] = g.U[q].re;
        evecs_ri[2 * q + 1] = g.U[q].im;
    }
    if (ZLANE == 0)
        *info = rc;
}

void z_ensure_smem(const void* fn, size_t bytes)
{
    if (bytes > 48 * 1024)
        SB200_CUDA_CHECK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) bytes));
}

}  // namespace

void launch_gen_restart_z(double* Hr, double* Hi, int m, int nev, const FacCtl* ctl, double beta, int use_beta, int selection, double tol, double* ritz_val_ri,
                          double* ritz_est_ri, double* ritz_vec_ri, int* ritz_conv, double* Qr, double* Qi, GenRestartOut* out, int do_restart,
                          cudaStream_t stream)
{
    SB200_REQUIRE(m >= 3 && m < kPanelMaxCols, SB200_INVALID_ARGUMENT, "ncv out of range for the complex device restart kernel");
    const size_t smem = zsmem_bytes(m);
    z_ensure_smem((const void*) gen_restart_z_kernel, smem);
    gen_restart_z_kernel<<<1, kZBlock, smem, stream>>>(Hr, Hi, m, nev, ctl, beta, use_beta, selection, tol, ritz_val_ri, ritz_est_ri, ritz_vec_ri, ritz_conv, Qr,
                                                       Qi, out, do_restart);
    SB200_CUDA_CHECK(cudaGetLastError());
}

void dense_hess_qr_z_host(int64_t m, const double* H_ri, double mu_re, double mu_im, double* QtHQ_ri, double* Q_ri)
{
    device_info();
    SB200_REQUIRE(m >= 2 && m < kPanelMaxCols, SB200_INVALID_ARGUMENT, "matrix order out of range");
    DevBuf<double> dH(2 * m * m), dD(2 * m * m), dQ(2 * m * m);
    SB200_CUDA_CHECK(cudaMemcpy(dH.get(), H_ri, sizeof(double) * 2 * m * m, cudaMemcpyHostToDevice));
    const size_t smem = zsmem_bytes((int) m);
    z_ensure_smem((const void*) hess_qr_z_kernel, smem);
    hess_qr_z_kernel<<<1, kZBlock, smem>>>(dH.get(), (int) m, mu_re, mu_im, dD.get(), dQ.get());
    SB200_CUDA_CHECK(cudaGetLastError());
    SB200_CUDA_CHECK(cudaDeviceSynchronize());
    SB200_CUDA_CHECK(cudaMemcpy(QtHQ_ri, dD.get(), sizeof(double) * 2 * m * m, cudaMemcpyDeviceToHost));
    SB200_CUDA_CHECK(cudaMemcpy(Q_ri, dQ.get(), sizeof(double) * 2 * m * m, cudaMemcpyDeviceToHost));
}

void dense_hess_eigen_z_host(int64_t m, const double* H_ri, double* evals_ri, double* evecs_ri)
{
    device_info();
    SB200_REQUIRE(m >= 1 && m < kPanelMaxCols, SB200_INVALID_ARGUMENT, "matrix order out of range");
    DevBuf<double> dH(2 * m * m), dE(2 * m), dV(2 * m * m);
    DevBuf<int> dinfo(1);
    SB200_CUDA_CHECK(cudaMemcpy(dH.get(), H_ri, sizeof(double) * 2 * m * m, cudaMemcpyHostToDevice));
    const size_t smem = zsmem_bytes((int) m);
    z_ensure_smem((const void*) hess_eigen_z_kernel, smem);
    hess_eigen_z_kernel<<<1, kZBlock, smem>>>(dH.get(), (int) m, dE.get(), dV.get(), dinfo.get());
    SB200_CUDA_CHECK(cudaGetLastError());
    SB200_CUDA_CHECK(cudaDeviceSynchronize());
    int info = 0;
    SB200_CUDA_CHECK(cudaMemcpy(&info, dinfo.get(), sizeof(int), cudaMemcpyDeviceToHost));
    if (info != 0)
        throw Error(SB200_RUNTIME, "UpperHessenbergEigen: eigen decomposition failed");
    SB200_CUDA_CHECK(cudaMemcpy(evals_ri, dE.get(), sizeof(double) * 2 * m, cudaMemcpyDeviceToHost));
    SB200_CUDA_CHECK(cudaMemcpy(evecs_ri, dV.get(), sizeof(double) * 2 * m * m, cudaMemcpyDeviceToHost));
}

}  // namespace sb200

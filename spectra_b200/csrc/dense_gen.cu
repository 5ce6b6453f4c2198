// Single-warp device kernels for the nonsymmetric restart (K9, K10 of SURVEY.md §2.1):
//   UpperHessenbergQR      LinAlg/UpperHessenbergQR.h:136-195 (compute), :219-255 (RQ + sI), :383-417 (YQ)
//   DoubleShiftQR          LinAlg/DoubleShiftQR.h:106-214, :218-314, :334-398, :425-437
//   UpperHessenbergSchur   LinAlg/UpperHessenbergSchur.h:44-166, :287-341, :354-425
//   UpperHessenbergEigen   LinAlg/UpperHessenbergEigen.h:53-208, :221-277, :287-320
//   retrieve_ritzpair / num_converged / nev_adjusted / RestartArnoldi::run
//                          GenEigsBase.h:280-340, :225-242, :245-277, :60-107
// One warp owns the m x m problem in shared memory.  Scalar recurrences (rotation / reflector
// generation, shift strategy, deflation tests) are evaluated redundantly by all lanes from shared
// memory; the O(m) row / column updates are spread over the lanes; __syncwarp() separates phases.
#include "dense_common.cuh"
#include "kernels.h"

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

constexpr int kGenBlock = 32;
#define LANE ((int) threadIdx.x)
#define FOR_LANES(var, lo, hi) for (int var = (lo) + LANE; var < (hi); var += kGenBlock)

struct Cx
{
    double re, im;
};
// Smith's algorithm, the scaling strategy of libgcc's __divdc3 without the inf/nan recovery
__device__ __forceinline__ Cx cdiv(double a, double b, double c, double d)
{
    Cx r;
    if (fabs(c) < fabs(d))
    {
        const double ratio = c / d, denom = c * ratio + d;
        r.re = (a * ratio + b) / denom;
        r.im = (b * ratio - a) / denom;
    }
    else
    {
        const double ratio = d / c, denom = d * ratio + c;
        r.re = (b * ratio + a) / denom;
        r.im = (b - a * ratio) / denom;
    }
    return r;
}

#define M_(A, i, j) (A)[(i) + (j) * m]

// ------------------------------------------------------------------------------------------
// UpperHessenbergQR on H (in place): H <- R Q + s I, Q <- Q G1 G2 ...   (rc/rs: m doubles each)
// ------------------------------------------------------------------------------------------
__device__ void hess_qr_shift(double* H, double* Q, int m, double shift, double* rc, double* rs)
{
    FOR_LANES(i, 0, m) M_(H, i, i) -= shift;
    __syncwarp();
    const int n1 = m - 1;
    for (int i = 0; i < n1; i++)
    {
        // zero below the sub-diagonal of column i (:163), rotation from (R[i,i], R[i+1,i])
        const double xi = M_(H, i, i), xj = M_(H, i + 1, i);
        double r, c, s;
        givens_rotation(xi, xj, r, c, s);
        __syncwarp();
        FOR_LANES(t, i + 2, m) M_(H, t, i) = 0.0;
        if (LANE == 0)
        {
            rc[i] = c;
            rs[i] = s;
            M_(H, i, i) = r;
            M_(H, i + 1, i) = 0.0;
        }
        FOR_LANES(j, i + 1, m)
        {
            const double tmp = M_(H, i, j), t1 = M_(H, i + 1, j);
            M_(H, i, j) = c * tmp - s * t1;
            M_(H, i + 1, j) = s * tmp + c * t1;
        }
        __syncwarp();
    }
    // RQ (:219-255): column pair (i, i+1), rows 0..i+1
    for (int i = 0; i < n1; i++)
    {
        const double c = rc[i], s = rs[i];
        FOR_LANES(j, 0, i + 2)
        {
            const double tmp = M_(H, j, i), t1 = M_(H, j, i + 1);
            M_(H, j, i) = c * tmp - s * t1;
            M_(H, j, i + 1) = s * tmp + c * t1;
        }
        __syncwarp();
    }
    FOR_LANES(i, 0, m) M_(H, i, i) += shift;
    // Q <- Q * G (:383-417): row-parallel
    if (Q)
    {
        FOR_LANES(t, 0, m)
        {
            double yi = M_(Q, t, 0);
            for (int i = 0; i < n1; i++)
            {
                const double c = rc[i], s = rs[i];
                const double yi1 = M_(Q, t, i + 1);
                M_(Q, t, i) = c * yi - s * yi1;
                yi = s * yi + c * yi1;
            }
            M_(Q, t, m - 1) = yi;
        }
    }
    __syncwarp();
}

// ------------------------------------------------------------------------------------------
// DoubleShiftQR on H (in place) for H^2 - s H + t I; Q <- Q P0 P1 ...
// ref_u: 3*m doubles, ref_nr: m ints
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double stable_norm3(double x1, double x2, double x3)
{
    x1 = fabs(x1);
    x2 = fabs(x2);
    x3 = fabs(x3);
    if (x1 < x2)
    {
        const double t = x1;
        x1 = x2;
        x2 = t;
    }
    if (x1 < x3)
    {
        const double t = x1;
        x1 = x3;
        x3 = t;
    }
    if (x1 < kNear0)
        return 0.0;
    const double r2 = x2 / x1, r3 = x3 / x1;
    const double cutoff = 0.1 * 1.220703125e-4;
    double r = r2 * r2 + r3 * r3;
    r = (r2 >= cutoff || r3 >= cutoff) ? sqrt(1.0 + r) : (1.0 + r * (0.5 - 0.125 * r));
    return x1 * r;
}

__device__ __forceinline__ void stable_scaling3(double& x1, double& x2, double& x3)
{
    const double x1sign = (x1 > 0.0) ? 1.0 : -1.0;
    x1 = fabs(x1);
    const double r2 = x2 / x1, r3 = x3 / x1;
    const double cutoff = 0.1 * 1.220703125e-4;
    double r = r2 * r2 + r3 * r3;
    r = (fabs(r2) >= cutoff || fabs(r3) >= cutoff) ? 1.0 / sqrt(1.0 + r) : (1.0 - r * (0.5 - 0.375 * r));
    x1 = x1sign * r;
    x2 = r2 * r;
    x3 = r3 * r;
}

// DoubleShiftQR.h:106-145.  All lanes compute; lane 0 stores.  Returns nr.
__device__ int compute_reflector(double x1, double x2, double x3, int ind, double* ref_u, int* ref_nr)
{
    const double x2m = fabs(x2), x3m = fabs(x3);
    if (x2m < kNear0 && x3m < kNear0)
    {
        if (LANE == 0)
            ref_nr[ind] = 1;
        return 1;
    }
    const int nr = (x3m < kNear0) ? 2 : 3;
    const double x_norm = (x3m < kNear0) ? eigen_hypot(x1, x2) : stable_norm3(x1, x2, x3);
    const double rho = double(x1 <= 0.0) - double(x1 > 0.0);
    const double x1_new = x1 - rho * x_norm, x1m = fabs(x1_new);
    double u0 = x1_new, u1 = x2, u2 = x3;
    if (x1m >= x2m && x1m >= x3m)
        stable_scaling3(u0, u1, u2);
    else if (x2m >= x1m && x2m >= x3m)
        stable_scaling3(u1, u0, u2);
    else
        stable_scaling3(u2, u0, u1);
    if (LANE == 0)
    {
        ref_nr[ind] = nr;
        ref_u[3 * ind] = u0;
        ref_u[3 * ind + 1] = u1;
        ref_u[3 * ind + 2] = u2;
    }
    return nr;
}

// PX on the block rows r0..r0+nrow-1, cols c0..c0+ncol-1 of A (:218-253), column-parallel
__device__ void apply_PX(double* A, int m, int r0, int c0, int nrow, int ncol, int nr, double u0, double u1, double u2)
{
    if (nr == 1)
        return;
    const double u0_2 = 2.0 * u0, u1_2 = 2.0 * u1;
    if (nr == 2 || nrow == 2)
    {
        FOR_LANES(j, 0, ncol)
        {
            double* x = &M_(A, r0, c0 + j);
            const double tmp = u0_2 * x[0] + u1_2 * x[1];
            x[0] -= tmp * u0;
            x[1] -= tmp * u1;
        }
    }
    else
    {
        const double u2_2 = 2.0 * u2;
        FOR_LANES(j, 0, ncol)
        {
            double* x = &M_(A, r0, c0 + j);
            const double tmp = u0_2 * x[0] + u1_2 * x[1] + u2_2 * x[2];
            x[0] -= tmp * u0;
            x[1] -= tmp * u1;
            x[2] -= tmp * u2;
        }
    }
}

// XP on rows r0..r0+nrow-1, cols c0..c0+ncol-1 (:278-314), row-parallel
__device__ void apply_XP(double* A, int m, int r0, int c0, int nrow, int ncol, int nr, double u0, double u1, double u2)
{
    if (nr == 1)
        return;
    const double u0_2 = 2.0 * u0, u1_2 = 2.0 * u1;
    double* X0 = &M_(A, r0, c0);
    double* X1 = X0 + m;
    if (nr == 2 || ncol == 2)
    {
        FOR_LANES(i, 0, nrow)
        {
            const double tmp = u0_2 * X0[i] + u1_2 * X1[i];
            X0[i] -= tmp * u0;
            X1[i] -= tmp * u1;
        }
    }
    else
    {
        double* X2 = X1 + m;
        const double u2_2 = 2.0 * u2;
        FOR_LANES(i, 0, nrow)
        {
            const double tmp = u0_2 * X0[i] + u1_2 * X1[i] + u2_2 * X2[i];
            X0[i] -= tmp * u0;
            X1[i] -= tmp * u1;
            X2[i] -= tmp * u2;
        }
    }
}

// one reflector of update_block: generate from (x1,x2,x3), then PX and XP on H
__device__ void ds_reflect(double* H, int m, double x1, double x2, double x3, int ind, int px_r0, int px_c0, int px_nrow, int px_ncol, int xp_nrow, int xp_ncol,
                           double* ref_u, int* ref_nr)
{
    __syncwarp();
    const int nr = compute_reflector(x1, x2, x3, ind, ref_u, ref_nr);
    __syncwarp();
    const double u0 = ref_u[3 * ind], u1 = ref_u[3 * ind + 1], u2 = ref_u[3 * ind + 2];
    apply_PX(H, m, px_r0, px_c0, px_nrow, px_ncol, nr, u0, u1, u2);
    __syncwarp();
    apply_XP(H, m, 0, ind, xp_nrow, xp_ncol, nr, u0, u1, u2);
    __syncwarp();
}

// DoubleShiftQR.h:153-214
__device__ void ds_update_block(double* H, int m, int il, int iu, double shift_s, double shift_t, double* ref_u, int* ref_nr)
{
    const int bsize = iu - il + 1;
    if (bsize == 1)
    {
        if (LANE == 0)
            ref_nr[il] = 1;
        __syncwarp();
        return;
    }
    const double x00 = M_(H, il, il), x01 = M_(H, il, il + 1), x10 = M_(H, il + 1, il), x11 = M_(H, il + 1, il + 1);
    const double m00 = x00 * (x00 - shift_s) + x01 * x10 + shift_t;
    const double m10 = x10 * (x00 + x11 - shift_s);
    if (bsize == 2)
    {
        ds_reflect(H, m, m00, m10, 0.0, il, il, il, 2, m - il, il + 2, 2, ref_u, ref_nr);
        if (LANE == 0)
            ref_nr[il + 1] = 1;
        __syncwarp();
        return;
    }
    const double m20 = M_(H, il + 2, il + 1) * M_(H, il + 1, il);
    ds_reflect(H, m, m00, m10, m20, il, il, il, 3, m - il, il + min(bsize, 4), 3, ref_u, ref_nr);
    for (int i = 1; i < bsize - 2; i++)
    {
        const double y0 = M_(H, il + i, il + i - 1), y1 = M_(H, il + i + 1, il + i - 1), y2 = M_(H, il + i + 2, il + i - 1);
        ds_reflect(H, m, y0, y1, y2, il + i, il + i, il + i - 1, 3, m - il - i + 1, il + min(bsize, i + 4), 3, ref_u, ref_nr);
    }
    const double z0 = M_(H, iu - 1, iu - 2), z1 = M_(H, iu, iu - 2);
    ds_reflect(H, m, z0, z1, 0.0, iu - 1, iu - 1, iu - 2, 2, m - iu + 2, il + bsize, 2, ref_u, ref_nr);
    if (LANE == 0)
        ref_nr[iu] = 1;
    __syncwarp();
}

// DoubleShiftQR::compute + matrix_QtHQ (in place) + apply_YQ(Q)   (:334-398, :425-437)
__device__ void double_shift_qr(double* H, double* Q, int m, double s, double t, double* ref_u, int* ref_nr, int* zero_ind)
{
    const double eps_abs = kNear0 * (double(m) / kEps);
    const double eps_rel = kEps;
    FOR_LANES(q, 0, 3 * m) ref_u[q] = 0.0;
    FOR_LANES(q, 0, m) ref_nr[q] = 0;
    __syncwarp();
    // deflation scan (sequential bookkeeping by lane 0, zero-fill parallel)
    if (LANE == 0)
    {
        int cnt = 0;
        zero_ind[cnt++] = 0;
        for (int i = 0; i < m - 1; i++)
        {
            const double h = fabs(M_(H, i + 1, i));
            const double diag = fabs(M_(H, i, i)) + fabs(M_(H, i + 1, i + 1));
            if (h <= eps_abs || h <= eps_rel * diag)
            {
                M_(H, i + 1, i) = 0.0;
                zero_ind[cnt++] = i + 1;
            }
        }
        zero_ind[cnt++] = m;
        zero_ind[m + 1] = cnt;
    }
    for (int i = 0; i < m - 1; i++)
        FOR_LANES(r, i + 2, m) M_(H, r, i) = 0.0;
    __syncwarp();
    const int len = zero_ind[m + 1] - 1;
    for (int b = 0; b < len; b++)
    {
        const int start = zero_ind[b], end = zero_ind[b + 1] - 1;
        ds_update_block(H, m, start, end, s, t, ref_u, ref_nr);
    }
    __syncwarp();
    FOR_LANES(i, 0, m - 1)
    {
        const double h = fabs(M_(H, i + 1, i));
        const double diag = fabs(M_(H, i, i)) + fabs(M_(H, i + 1, i + 1));
        if (h <= eps_abs || h <= eps_rel * diag)
            M_(H, i + 1, i) = 0.0;
    }
    __syncwarp();
    if (Q)
    {
        // Y Q = Y P0 P1 ... (row-local, reflectors sequential)
        FOR_LANES(r, 0, m)
        {
            for (int i = 0; i < m - 1; i++)
            {
                const int nr = ref_nr[i];
                if (nr == 1)
                    continue;
                const double u0 = ref_u[3 * i], u1 = ref_u[3 * i + 1], u2 = ref_u[3 * i + 2];
                const bool two = (nr == 2) || (i == m - 2);
                if (two)
                {
                    const double a = M_(Q, r, i), b2 = M_(Q, r, i + 1);
                    const double tmp = 2.0 * u0 * a + 2.0 * u1 * b2;
                    M_(Q, r, i) = a - tmp * u0;
                    M_(Q, r, i + 1) = b2 - tmp * u1;
                }
                else
                {
                    const double a = M_(Q, r, i), b2 = M_(Q, r, i + 1), c2 = M_(Q, r, i + 2);
                    const double tmp = 2.0 * u0 * a + 2.0 * u1 * b2 + 2.0 * u2 * c2;
                    M_(Q, r, i) = a - tmp * u0;
                    M_(Q, r, i + 1) = b2 - tmp * u1;
                    M_(Q, r, i + 2) = c2 - tmp * u2;
                }
            }
        }
    }
    __syncwarp();
}

// ------------------------------------------------------------------------------------------
// Real Schur form of a Hessenberg matrix: T (in place), U accumulated   (UpperHessenbergSchur.h)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void make_householder3(double v0, double v1, double v2, double& e0, double& e1, double& tau, double& beta)
{
    const double tail_sq = v1 * v1 + v2 * v2;
    if (tail_sq <= kMin)
    {
        tau = 0.0;
        beta = v0;
        e0 = e1 = 0.0;
    }
    else
    {
        beta = sqrt(v0 * v0 + tail_sq);
        if (v0 >= 0.0)
            beta = -beta;
        e0 = v1 / (v0 - beta);
        e1 = v2 / (v0 - beta);
        tau = (beta - v0) / beta;
    }
}

// rows p,q of A for columns col_from..m-1:  applyOnTheLeft(p, q, rot.adjoint())
__device__ void rot_left(double* A, int m, int col_from, int p, int q, double c, double s)
{
    FOR_LANES(j, col_from, m)
    {
        const double x = M_(A, p, j), y = M_(A, q, j);
        M_(A, p, j) = c * x - s * y;
        M_(A, q, j) = s * x + c * y;
    }
}
// columns p,q of A for rows 0..nrow-1: applyOnTheRight(p, q, rot)
__device__ void rot_right(double* A, int m, int nrow, int p, int q, double c, double s)
{
    FOR_LANES(i, 0, nrow)
    {
        const double x = M_(A, i, p), y = M_(A, i, q);
        M_(A, i, p) = c * x - s * y;
        M_(A, i, q) = s * x + c * y;
    }
}

// returns 0 on success, 1 when the 40*m iteration cap is hit   (:354-425)
__device__ int hess_schur(double* T, double* U, int m)
{
    FOR_LANES(q, 0, m * m) U[q] = ((q % m) == (q / m)) ? 1.0 : 0.0;
    __syncwarp();
    const int max_iter = m * 40;
    int iu = m - 1, iter = 0, total_iter = 0;
    double ex_shift = 0.0;
    // L1 norm of the Hessenberg part (:44-51), sequential order
    double norm = 0.0;
    for (int j = 0; j < m; j++)
    {
        const int len = min(m, j + 2);
        double sc = 0.0;
        for (int i = 0; i < len; i++)
            sc += fabs(M_(T, i, j));
        norm += sc;
    }
    const double near_0 = fmax(norm * kEps * kEps, kMin);
    if (norm == 0.0)
        return 0;
    while (iu >= 0)
    {
        // find_small_subdiag (:54-73)
        int il = iu;
        while (il > 0)
        {
            double sc = fabs(M_(T, il - 1, il - 1)) + fabs(M_(T, il, il));
            sc = fmax(sc * kEps, near_0);
            if (fabs(M_(T, il, il - 1)) <= sc)
                break;
            il--;
        }
        __syncwarp();
        if (il == iu)
        {
            if (LANE == 0)
            {
                M_(T, iu, iu) += ex_shift;
                if (iu > 0)
                    M_(T, iu, iu - 1) = 0.0;
            }
            iu--;
            iter = 0;
            __syncwarp();
        }
        else if (il == iu - 1)
        {
            // split_off_two_rows (:76-100)
            const double p = 0.5 * (M_(T, iu - 1, iu - 1) - M_(T, iu, iu));
            const double q = p * p + M_(T, iu, iu - 1) * M_(T, iu - 1, iu);
            const double tsub = M_(T, iu, iu - 1);
            __syncwarp();
            if (LANE == 0)
            {
                M_(T, iu, iu) += ex_shift;
                M_(T, iu - 1, iu - 1) += ex_shift;
            }
            __syncwarp();
            if (q >= 0.0)
            {
                const double z = sqrt(fabs(q));
                double c, s;
                make_givens((p >= 0.0) ? (p + z) : (p - z), tsub, c, s);
                rot_left(T, m, iu - 1, iu - 1, iu, c, s);
                __syncwarp();
                rot_right(T, m, iu + 1, iu - 1, iu, c, s);
                __syncwarp();
                if (LANE == 0)
                    M_(T, iu, iu - 1) = 0.0;
                rot_right(U, m, m, iu - 1, iu, c, s);
            }
            __syncwarp();
            if (iu > 1 && LANE == 0)
                M_(T, iu - 1, iu - 2) = 0.0;
            iu -= 2;
            iter = 0;
            __syncwarp();
        }
        else
        {
            // compute_shift (:103-142)
            double sh0 = M_(T, iu, iu), sh1 = M_(T, iu - 1, iu - 1), sh2 = M_(T, iu, iu - 1) * M_(T, iu - 1, iu);
            if (iter == 10)
            {
                ex_shift += sh0;
                __syncwarp();
                FOR_LANES(i, 0, iu + 1) M_(T, i, i) -= sh0;
                __syncwarp();
                const double sc = fabs(M_(T, iu, iu - 1)) + fabs(M_(T, iu - 1, iu - 2));
                sh0 = 0.75 * sc;
                sh1 = 0.75 * sc;
                sh2 = -0.4375 * sc * sc;
            }
            if (iter == 30)
            {
                double sc = (sh1 - sh0) / 2.0;
                sc = sc * sc + sh2;
                if (sc > 0.0)
                {
                    sc = sqrt(sc);
                    if (sh1 < sh0)
                        sc = -sc;
                    sc = sc + (sh1 - sh0) / 2.0;
                    sc = sh0 - sh2 / sc;
                    ex_shift += sc;
                    __syncwarp();
                    FOR_LANES(i, 0, iu + 1) M_(T, i, i) -= sc;
                    __syncwarp();
                    sh0 = sh1 = sh2 = 0.964;
                }
            }
            iter++;
            total_iter++;
            if (total_iter > max_iter)
                break;
            // init_francis_qr_step (:145-166)
            int im;
            double v0 = 0, v1 = 0, v2 = 0;
            for (im = iu - 2; im >= il; --im)
            {
                const double Tmm = M_(T, im, im);
                const double r = sh0 - Tmm;
                const double sc = sh1 - Tmm;
                v0 = (r * sc - sh2) / M_(T, im + 1, im) + M_(T, im, im + 1);
                v1 = M_(T, im + 1, im + 1) - Tmm - r - sc;
                v2 = M_(T, im + 2, im + 1);
                if (im == il)
                    break;
                const double lhs = M_(T, im, im - 1) * (fabs(v1) + fabs(v2));
                const double rhs = v0 * (fabs(M_(T, im - 1, im - 1)) + fabs(Tmm) + fabs(M_(T, im + 1, im + 1)));
                if (fabs(lhs) < kEps * rhs)
                    break;
            }
            // perform_francis_qr_step (:287-341)
            for (int k = im; k <= iu - 2; ++k)
            {
                const bool first_iter = (k == im);
                double w0, w1, w2;
                if (first_iter)
                {
                    w0 = v0;
                    w1 = v1;
                    w2 = v2;
                }
                else
                {
                    w0 = M_(T, k, k - 1);
                    w1 = M_(T, k + 1, k - 1);
                    w2 = M_(T, k + 2, k - 1);
                }
                double e0, e1, tau, beta;
                make_householder3(w0, w1, w2, e0, e1, tau, beta);
                __syncwarp();
                if (fabs(beta) > near_0)
                {
                    if (LANE == 0)
                    {
                        if (first_iter && k > il)
                            M_(T, k, k - 1) = -M_(T, k, k - 1);
                        else if (!first_iter)
                            M_(T, k, k - 1) = beta;
                    }
                    // left: rows k..k+2, columns k..m-1
                    FOR_LANES(j, k, m)
                    {
                        double* x = &M_(T, k, j);
                        const double tvx = tau * (x[0] + e0 * x[1] + e1 * x[2]);
                        x[0] -= tvx;
                        x[1] -= tvx * e0;
                        x[2] -= tvx * e1;
                    }
                    __syncwarp();
                    // right on T: rows 0..min(iu, k+3), columns k..k+2
                    const int nrow = min(iu, k + 3) + 1;
                    FOR_LANES(i, 0, nrow)
                    {
                        const double x0 = M_(T, i, k), x1 = M_(T, i, k + 1), x2 = M_(T, i, k + 2);
                        const double txv = tau * (x0 + e0 * x1 + e1 * x2);
                        M_(T, i, k) = x0 - txv;
                        M_(T, i, k + 1) = x1 - txv * e0;
                        M_(T, i, k + 2) = x2 - txv * e1;
                    }
                    // right on U: all rows
                    FOR_LANES(i, 0, m)
                    {
                        const double x0 = M_(U, i, k), x1 = M_(U, i, k + 1), x2 = M_(U, i, k + 2);
                        const double txv = tau * (x0 + e0 * x1 + e1 * x2);
                        M_(U, i, k) = x0 - txv;
                        M_(U, i, k + 1) = x1 - txv * e0;
                        M_(U, i, k + 2) = x2 - txv * e1;
                    }
                }
                __syncwarp();
            }
            {
                double c, s, beta;
                make_givens(M_(T, iu - 1, iu - 2), M_(T, iu, iu - 2), c, s, &beta);
                __syncwarp();
                if (fabs(beta) > near_0)
                {
                    if (LANE == 0)
                        M_(T, iu - 1, iu - 2) = beta;
                    rot_left(T, m, iu - 1, iu - 1, iu, c, s);
                    __syncwarp();
                    rot_right(T, m, iu + 1, iu - 1, iu, c, s);
                    rot_right(U, m, m, iu - 1, iu, c, s);
                }
                __syncwarp();
            }
            // clean up pollution (:334-340)
            FOR_LANES(i, im + 2, iu + 1)
            {
                M_(T, i, i - 2) = 0.0;
                if (i > im + 2)
                    M_(T, i, i - 3) = 0.0;
            }
            __syncwarp();
        }
    }
    return (total_iter > max_iter) ? 1 : 0;
}

// warp-parallel dot  row(i)[l..l+len) . col(n)[l..l+len)  of T
__device__ __forceinline__ double row_col_dot(const double* T, int m, int i, int n, int l, int len)
{
    double s = 0.0;
    FOR_LANES(k, 0, len) s += M_(T, i, l + k) * M_(T, l + k, n);
    return warp_sum(s);
}

// UpperHessenbergEigen::compute without the final complex assembly (:221-277, :53-208).
// T holds the quasi-triangular Schur factor (overwritten by the back-substituted vectors),
// U the Schur vectors (overwritten by the real eigenvector basis), ev_re/ev_im the eigenvalues.
__device__ void hess_eigen_real(double* T, double* U, int m, double* ev_re, double* ev_im, double* tmp)
{
    // eigenvalues from the diagonal blocks (:239-268), sequential scan by every lane, lane 0 stores
    {
        int i = 0;
        while (i < m)
        {
            if (i == m - 1 || M_(T, i + 1, i) == 0.0)
            {
                if (LANE == 0)
                {
                    ev_re[i] = M_(T, i, i);
                    ev_im[i] = 0.0;
                }
                ++i;
            }
            else
            {
                const double p = 0.5 * (M_(T, i, i) - M_(T, i + 1, i + 1));
                double t0 = M_(T, i + 1, i), t1 = M_(T, i, i + 1);
                const double maxval = fmax(fabs(p), fmax(fabs(t0), fabs(t1)));
                t0 /= maxval;
                t1 /= maxval;
                const double p0 = p / maxval;
                const double z = maxval * sqrt(fabs(p0 * p0 + t0 * t1));
                if (LANE == 0)
                {
                    ev_re[i] = M_(T, i + 1, i + 1) + p;
                    ev_im[i] = z;
                    ev_re[i + 1] = M_(T, i + 1, i + 1) + p;
                    ev_im[i + 1] = -z;
                }
                i += 2;
            }
        }
    }
    __syncwarp();
    // doComputeEigenvectors (:53-208)
    double norm = 0.0;
    for (int j = 0; j < m; ++j)
    {
        const int from = max(j - 1, 0);
        double sc = 0.0;
        for (int k = from; k < m; k++)
            sc += fabs(M_(T, j, k));
        norm += sc;
    }
    if (norm == 0.0)
        return;
    const double eps = kEps;
    for (int n = m - 1; n >= 0; n--)
    {
        const double p = ev_re[n], q = ev_im[n];
        if (q == 0.0)
        {
            double lastr = 0.0, lastw = 0.0;
            int l = n;
            __syncwarp();
            if (LANE == 0)
                M_(T, n, n) = 1.0;
            __syncwarp();
            for (int i = n - 1; i >= 0; i--)
            {
                const double w = M_(T, i, i) - p;
                const double r = row_col_dot(T, m, i, n, l, n - l + 1);
                if (ev_im[i] < 0.0)
                {
                    lastw = w;
                    lastr = r;
                }
                else
                {
                    l = i;
                    __syncwarp();
                    if (ev_im[i] == 0.0)
                    {
                        const double val = (w != 0.0) ? (-r / w) : (-r / (eps * norm));
                        if (LANE == 0)
                            M_(T, i, n) = val;
                    }
                    else
                    {
                        const double x = M_(T, i, i + 1), y = M_(T, i + 1, i);
                        const double denom = (ev_re[i] - p) * (ev_re[i] - p) + ev_im[i] * ev_im[i];
                        const double t = (x * lastr - lastw * r) / denom;
                        const double t2 = (fabs(x) > fabs(lastw)) ? ((-r - w * t) / x) : ((-lastr - y * t) / lastw);
                        if (LANE == 0)
                        {
                            M_(T, i, n) = t;
                            M_(T, i + 1, n) = t2;
                        }
                    }
                    __syncwarp();
                    const double t = fabs(M_(T, i, n));
                    __syncwarp();  // every lane must have read t before the lane that owns row i rescales T(i, n)
                    if ((eps * t) * t > 1.0)
                    {
                        FOR_LANES(k, i, m) M_(T, k, n) /= t;
                    }
                    __syncwarp();
                }
            }
        }
        else if (q < 0.0 && n > 0)
        {
            double lastra = 0.0, lastsa = 0.0, lastw = 0.0;
            int l = n - 1;
            __syncwarp();
            {
                double a, b;
                if (fabs(M_(T, n, n - 1)) > fabs(M_(T, n - 1, n)))
                {
                    a = q / M_(T, n, n - 1);
                    b = -(M_(T, n, n) - p) / M_(T, n, n - 1);
                }
                else
                {
                    const Cx cc = cdiv(0.0, -M_(T, n - 1, n), M_(T, n - 1, n - 1) - p, q);
                    a = cc.re;
                    b = cc.im;
                }
                __syncwarp();
                if (LANE == 0)
                {
                    M_(T, n - 1, n - 1) = a;
                    M_(T, n - 1, n) = b;
                    M_(T, n, n - 1) = 0.0;
                    M_(T, n, n) = 1.0;
                }
            }
            __syncwarp();
            for (int i = n - 2; i >= 0; i--)
            {
                const double ra = row_col_dot(T, m, i, n - 1, l, n - l + 1);
                const double sa = row_col_dot(T, m, i, n, l, n - l + 1);
                const double w = M_(T, i, i) - p;
                if (ev_im[i] < 0.0)
                {
                    lastw = w;
                    lastra = ra;
                    lastsa = sa;
                }
                else
                {
                    l = i;
                    __syncwarp();
                    if (ev_im[i] == 0.0)
                    {
                        const Cx cc = cdiv(-ra, -sa, w, q);
                        if (LANE == 0)
                        {
                            M_(T, i, n - 1) = cc.re;
                            M_(T, i, n) = cc.im;
                        }
                    }
                    else
                    {
                        const double x = M_(T, i, i + 1), y = M_(T, i + 1, i);
                        double vr = (ev_re[i] - p) * (ev_re[i] - p) + ev_im[i] * ev_im[i] - q * q;
                        const double vi = (ev_re[i] - p) * 2.0 * q;
                        if ((vr == 0.0) && (vi == 0.0))
                            vr = eps * norm * (fabs(w) + fabs(q) + fabs(x) + fabs(y) + fabs(lastw));
                        const Cx cc = cdiv(x * lastra - lastw * ra + q * sa, x * lastsa - lastw * sa - q * ra, vr, vi);
                        double b0, b1;
                        if (fabs(x) > (fabs(lastw) + fabs(q)))
                        {
                            b0 = (-ra - w * cc.re + q * cc.im) / x;
                            b1 = (-sa - w * cc.im - q * cc.re) / x;
                        }
                        else
                        {
                            const Cx c2 = cdiv(-lastra - y * cc.re, -lastsa - y * cc.im, lastw, q);
                            b0 = c2.re;
                            b1 = c2.im;
                        }
                        if (LANE == 0)
                        {
                            M_(T, i, n - 1) = cc.re;
                            M_(T, i, n) = cc.im;
                            M_(T, i + 1, n - 1) = b0;
                            M_(T, i + 1, n) = b1;
                        }
                    }
                    __syncwarp();
                    const double t = fmax(fabs(M_(T, i, n - 1)), fabs(M_(T, i, n)));
                    __syncwarp();  // as above: t is read by all lanes before row i is rescaled
                    if ((eps * t) * t > 1.0)
                    {
                        FOR_LANES(k, i, m)
                        {
                            M_(T, k, n - 1) /= t;
                            M_(T, k, n) /= t;
                        }
                    }
                    __syncwarp();
                }
            }
            n--;
        }
    }
    __syncwarp();
    // back transformation (:202-207): U(:, j) = U(:, 0..j) * T(0..j, j), j descending, row-parallel
    for (int j = m - 1; j >= 0; j--)
    {
        FOR_LANES(i, 0, m)
        {
            double s = 0.0;
            for (int k = 0; k <= j; k++)
                s += M_(U, i, k) * M_(T, k, j);
            tmp[i] = s;
        }
        __syncwarp();
        FOR_LANES(i, 0, m) M_(U, i, j) = tmp[i];
        __syncwarp();
    }
}

// complex eigenvector j of the assembled matV (:287-320), element i, before normalisation
__device__ __forceinline__ Cx eigvec_elem(const double* U, int m, const double* ev_im, int i, int j)
{
    Cx r;
    if (ev_im[j] == 0.0 || (ev_im[j] > 0.0 && j + 1 == m))
    {
        r.re = M_(U, i, j);
        r.im = 0.0;
    }
    else if (ev_im[j] > 0.0)
    {
        r.re = M_(U, i, j);
        r.im = M_(U, i, j + 1);
    }
    else
    {
        r.re = M_(U, i, j - 1);
        r.im = -M_(U, i, j);
    }
    return r;
}

__device__ __forceinline__ double sort_key_complex(int rule, double re, double im)
{
    switch (rule)
    {
        case SB200_LARGEST_MAGN: return -hypot(re, im);
        case SB200_LARGEST_REAL: return -re;
        case SB200_LARGEST_IMAG: return -fabs(im);
        case SB200_SMALLEST_MAGN: return hypot(re, im);
        case SB200_SMALLEST_REAL: return re;
        default: return fabs(im);  // SB200_SMALLEST_IMAG
    }
}

struct GenShared
{
    double *H, *T, *U;         // m x m each
    double *ev_re, *ev_im;     // m
    double *key, *cn, *tmp;    // m (sort keys, column norms, scratch)
    double *rc, *rs;           // m
    double *ref_u;             // 3 m
    double *rv_re, *rv_im;     // sorted Ritz values, m
    double *re_abs;            // |ritz_est|, m
    int *idx, *ref_nr, *zero_ind;  // m, m, m + 2
};

__device__ GenShared carve_gen(double* smem, int m)
{
    GenShared g;
    double* p = smem;
    g.H = p;
    p += m * m;
    g.T = p;
    p += m * m;
    g.U = p;
    p += m * m;
    g.ev_re = p;
    p += m;
    g.ev_im = p;
    p += m;
    g.key = p;
    p += m;
    g.cn = p;
    p += m;
    g.tmp = p;
    p += m;
    g.rc = p;
    p += m;
    g.rs = p;
    p += m;
    g.ref_u = p;
    p += 3 * m;
    g.rv_re = p;
    p += m;
    g.rv_im = p;
    p += m;
    g.re_abs = p;
    p += m;
    g.idx = reinterpret_cast<int*>(p);
    g.ref_nr = g.idx + m;
    g.zero_ind = g.ref_nr + m;
    return g;
}
size_t gen_smem_bytes(int m) { return sizeof(double) * (size_t) (3 * m * m + 13 * m) + sizeof(int) * (size_t) (3 * m + 4); }

// scale (:231-234), Schur, eigenvalues, eigenvectors; returns info.  Column norms in g.cn.
__device__ int hess_eigen_block(GenShared& g, int m)
{
    double scale = 0.0;
    for (int q = 0; q < m * m; q++)
        scale = fmax(scale, fabs(g.H[q]));
    FOR_LANES(q, 0, m * m) g.T[q] = g.H[q] / scale;
    __syncwarp();
    const int info = hess_schur(g.T, g.U, m);
    __syncwarp();
    hess_eigen_real(g.T, g.U, m, g.ev_re, g.ev_im, g.tmp);
    __syncwarp();
    // column norms of the complex vectors (:299-300, :310-311) and eigenvalue un-scaling (:274)
    FOR_LANES(j, 0, m)
    {
        double sq = 0.0;
        for (int i = 0; i < m; i++)
        {
            const Cx e = eigvec_elem(g.U, m, g.ev_im, i, j);
            sq += e.re * e.re + e.im * e.im;
        }
        g.cn[j] = sq > 0.0 ? sqrt(sq) : 1.0;
    }
    __syncwarp();
    FOR_LANES(j, 0, m)
    {
        g.ev_re[j] *= scale;
        g.ev_im[j] *= scale;
    }
    __syncwarp();
    return info;
}

__global__ void __launch_bounds__(kGenBlock)
    gen_restart_kernel(double* H, int m, int nev, const FacCtl* ctl, double beta_override, int use_override, int selection, double tol, double* ritz_val_ri,
                       double* ritz_est_ri, double* ritz_vec_ri, int* ritz_conv, double* Q, GenRestartOut* out, int do_restart)
{
    extern __shared__ double smem[];
    GenShared g = carve_gen(smem, m);
    FOR_LANES(q, 0, m * m) g.H[q] = H[q];
    __syncwarp();
    // ---- retrieve_ritzpair (GenEigsBase.h:280-340) ----
    const int info = hess_eigen_block(g, m);
    if (LANE == 0)
    {
        for (int i = 0; i < m; i++)
            g.key[i] = sort_key_complex(selection, g.ev_re[i], g.ev_im[i]);
        argsort_keys(g.key, g.idx, m);
    }
    __syncwarp();
    FOR_LANES(t, 0, m)
    {
        const int id = g.idx[t];
        const double re = g.ev_re[id], im = g.ev_im[id];
        ritz_val_ri[2 * t] = re;
        ritz_val_ri[2 * t + 1] = im;
        g.rv_re[t] = re;
        g.rv_im[t] = im;
        const Cx e = eigvec_elem(g.U, m, g.ev_im, m - 1, id);
        const double cn = g.cn[id];
        ritz_est_ri[2 * t] = e.re / cn;
        ritz_est_ri[2 * t + 1] = e.im / cn;
        g.re_abs[t] = hypot(e.re / cn, e.im / cn);
    }
    for (int c = 0; c < nev; c++)
    {
        const int id = g.idx[c];
        const double cn = g.cn[id];
        FOR_LANES(r, 0, m)
        {
            const Cx e = eigvec_elem(g.U, m, g.ev_im, r, id);
            ritz_vec_ri[2 * (r + c * m)] = e.re / cn;
            ritz_vec_ri[2 * (r + c * m) + 1] = e.im / cn;
        }
    }
    __syncwarp();
    // ---- num_converged (:225-242), nev_adjusted (:245-277) ----
    const double beta = use_override ? beta_override : ctl->beta;
    const double eps23 = 3.666852862501036e-11;
    int nconv = 0;
    for (int i = 0; i < nev; i++)
    {
        const double thresh = tol * fmax(hypot(g.rv_re[i], g.rv_im[i]), eps23);
        const double resid = g.re_abs[i] * beta;
        const int cv = resid < thresh;
        if (LANE == 0)
            ritz_conv[i] = cv;
        nconv += cv;
    }
    int nev_new = nev;
    for (int i = nev; i < m; i++)
        if (g.re_abs[i] < kNear0)
            nev_new++;
    nev_new += min(nconv, (m - nev_new) / 2);
    if (nev_new == 1 && m >= 6)
        nev_new = m / 2;
    else if (nev_new == 1 && m > 3)
        nev_new = 2;
    if (nev_new > m - 2)
        nev_new = m - 2;
    if (nev_new >= 1 && g.rv_im[nev_new - 1] != 0.0 && g.rv_re[nev_new - 1] == g.rv_re[nev_new] && g.rv_im[nev_new - 1] == -g.rv_im[nev_new])
        nev_new++;
    if (LANE == 0)
    {
        out->nconv = nconv;
        out->k = nev_new;
        out->info = info;
    }
    if (!(do_restart && nconv < nev && info == 0 && nev_new < m))
        return;

    // ---- RestartArnoldi::run (:60-107): Q accumulates in the T buffer ----
    double* Qs = g.T;
    FOR_LANES(q, 0, m * m) Qs[q] = ((q % m) == (q / m)) ? 1.0 : 0.0;
    __syncwarp();
    for (int i = nev_new; i < m; i++)
    {
        const bool cplx = g.rv_im[i] != 0.0;
        if (cplx && i + 1 < m && g.rv_re[i] == g.rv_re[i + 1] && g.rv_im[i] == -g.rv_im[i + 1])
        {
            const double s = 2.0 * g.rv_re[i];
            const double t = g.rv_re[i] * g.rv_re[i] + g.rv_im[i] * g.rv_im[i];
            double_shift_qr(g.H, Qs, m, s, t, g.ref_u, g.ref_nr, g.zero_ind);
            i++;
        }
        else
        {
            hess_qr_shift(g.H, Qs, m, g.rv_re[i], g.rc, g.rs);
        }
    }
    __syncwarp();
    FOR_LANES(q, 0, m * m)
    {
        H[q] = g.H[q];
        Q[q] = Qs[q];
    }
}

// ---- standalone kernels for the unit tier ----
__global__ void __launch_bounds__(kGenBlock) hess_qr_kernel(const double* H, int m, double shift, double* QtHQ, double* Q)
{
    extern __shared__ double smem[];
    GenShared g = carve_gen(smem, m);
    FOR_LANES(q, 0, m * m)
    {
        g.H[q] = H[q];
        g.T[q] = ((q % m) == (q / m)) ? 1.0 : 0.0;
    }
    __syncwarp();
    hess_qr_shift(g.H, g.T, m, shift, g.rc, g.rs);
    FOR_LANES(q, 0, m * m)
    {
        QtHQ[q] = g.H[q];
        Q[q] = g.T[q];
    }
}

__global__ void __launch_bounds__(kGenBlock) double_shift_qr_kernel(const double* H, int m, double s, double t, double* QtHQ, double* Q)
{
    extern __shared__ double smem[];
    GenShared g = carve_gen(smem, m);
    FOR_LANES(q, 0, m * m)
    {
        g.H[q] = H[q];
        g.T[q] = ((q % m) == (q / m)) ? 1.0 : 0.0;
    }
    __syncwarp();
    double_shift_qr(g.H, g.T, m, s, t, g.ref_u, g.ref_nr, g.zero_ind);
    FOR_LANES(q, 0, m * m)
    {
        QtHQ[q] = g.H[q];
        Q[q] = g.T[q];
    }
}

__global__ void __launch_bounds__(kGenBlock) hess_eigen_kernel(const double* H, int m, double* evals_ri, double* evecs_ri, int* info)
{
    extern __shared__ double smem[];
    GenShared g = carve_gen(smem, m);
    FOR_LANES(q, 0, m * m) g.H[q] = H[q];
    __syncwarp();
    const int rc = hess_eigen_block(g, m);
    FOR_LANES(j, 0, m)
    {
        evals_ri[2 * j] = g.ev_re[j];
        evals_ri[2 * j + 1] = g.ev_im[j];
    }
    for (int j = 0; j < m; j++)
    {
        const double cn = g.cn[j];
        FOR_LANES(i, 0, m)
        {
            const Cx e = eigvec_elem(g.U, m, g.ev_im, i, j);
            evecs_ri[2 * (i + j * m)] = e.re / cn;
            evecs_ri[2 * (i + j * m) + 1] = e.im / cn;
        }
    }
    if (LANE == 0)
        *info = rc;
}

void gen_ensure_smem(const void* fn, size_t bytes)
{
    if (bytes > 48 * 1024)
        SB200_CUDA_CHECK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) bytes));
}

}  // namespace

void launch_gen_restart(double* H, int m, int nev, const FacCtl* ctl, double beta, int use_beta, int selection, double tol, double* ritz_val_ri,
                        double* ritz_est_ri, double* ritz_vec_ri, int* ritz_conv, double* Q, GenRestartOut* out, int do_restart, cudaStream_t stream)
{
    SB200_REQUIRE(m >= 3 && m <= kPanelMaxCols, SB200_INVALID_ARGUMENT, "ncv out of range for the device restart kernel");
    const size_t smem = gen_smem_bytes(m);
    gen_ensure_smem((const void*) gen_restart_kernel, smem);
    gen_restart_kernel<<<1, kGenBlock, smem, stream>>>(H, m, nev, ctl, beta, use_beta, selection, tol, ritz_val_ri, ritz_est_ri, ritz_vec_ri, ritz_conv, Q, out,
                                                       do_restart);
    SB200_CUDA_CHECK(cudaGetLastError());
}

void dense_hess_qr_host(int64_t m, const double* H, double shift, double* QtHQ, double* Q)
{
    device_info();
    SB200_REQUIRE(m >= 2 && m <= kPanelMaxCols, SB200_INVALID_ARGUMENT, "matrix order out of range");
    DevBuf<double> dH(m * m), dD(m * m), dQ(m * m);
    SB200_CUDA_CHECK(cudaMemcpy(dH.get(), H, sizeof(double) * m * m, cudaMemcpyHostToDevice));
    const size_t smem = gen_smem_bytes((int) m);
    gen_ensure_smem((const void*) hess_qr_kernel, smem);
    hess_qr_kernel<<<1, kGenBlock, smem>>>(dH.get(), (int) m, shift, dD.get(), dQ.get());
    SB200_CUDA_CHECK(cudaGetLastError());
    SB200_CUDA_CHECK(cudaDeviceSynchronize());
    SB200_CUDA_CHECK(cudaMemcpy(QtHQ, dD.get(), sizeof(double) * m * m, cudaMemcpyDeviceToHost));
    SB200_CUDA_CHECK(cudaMemcpy(Q, dQ.get(), sizeof(double) * m * m, cudaMemcpyDeviceToHost));
}

void dense_double_shift_qr_host(int64_t m, const double* H, double s, double t, double* QtHQ, double* Q)
{
    device_info();
    SB200_REQUIRE(m >= 3 && m <= kPanelMaxCols, SB200_INVALID_ARGUMENT, "matrix order out of range");
    DevBuf<double> dH(m * m), dD(m * m), dQ(m * m);
    SB200_CUDA_CHECK(cudaMemcpy(dH.get(), H, sizeof(double) * m * m, cudaMemcpyHostToDevice));
    const size_t smem = gen_smem_bytes((int) m);
    gen_ensure_smem((const void*) double_shift_qr_kernel, smem);
    double_shift_qr_kernel<<<1, kGenBlock, smem>>>(dH.get(), (int) m, s, t, dD.get(), dQ.get());
    SB200_CUDA_CHECK(cudaGetLastError());
    SB200_CUDA_CHECK(cudaDeviceSynchronize());
    SB200_CUDA_CHECK(cudaMemcpy(QtHQ, dD.get(), sizeof(double) * m * m, cudaMemcpyDeviceToHost));
    SB200_CUDA_CHECK(cudaMemcpy(Q, dQ.get(), sizeof(double) * m * m, cudaMemcpyDeviceToHost));
}

void dense_hess_eigen_host(int64_t m, const double* H, double* evals_ri, double* evecs_ri)
{
    device_info();
    SB200_REQUIRE(m >= 1 && m <= kPanelMaxCols, SB200_INVALID_ARGUMENT, "matrix order out of range");
    DevBuf<double> dH(m * m), dE(2 * m), dV(2 * m * m);
    DevBuf<int> dinfo(1);
    SB200_CUDA_CHECK(cudaMemcpy(dH.get(), H, sizeof(double) * m * m, cudaMemcpyHostToDevice));
    const size_t smem = gen_smem_bytes((int) m);
    gen_ensure_smem((const void*) hess_eigen_kernel, smem);
    hess_eigen_kernel<<<1, kGenBlock, smem>>>(dH.get(), (int) m, dE.get(), dV.get(), dinfo.get());
    SB200_CUDA_CHECK(cudaGetLastError());
    SB200_CUDA_CHECK(cudaDeviceSynchronize());
    int info = 0;
    SB200_CUDA_CHECK(cudaMemcpy(&info, dinfo.get(), sizeof(int), cudaMemcpyDeviceToHost));
    if (info != 0)
        throw Error(SB200_RUNTIME, "UpperHessenbergSchur: Schur decomposition failed");
    SB200_CUDA_CHECK(cudaMemcpy(evals_ri, dE.get(), sizeof(double) * 2 * m, cudaMemcpyDeviceToHost));
    SB200_CUDA_CHECK(cudaMemcpy(evecs_ri, dV.get(), sizeof(double) * 2 * m * m, cudaMemcpyDeviceToHost));
}

}  // namespace sb200

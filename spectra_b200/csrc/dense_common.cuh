// Scalar building blocks of the small dense device kernels (K7-K10 of SURVEY.md §2.1).
// Restated for the GPU from the reference formulas; every function names the lines it follows.
#pragma once

#include "common.cuh"

namespace sb200 {
namespace dense {

// Eigen::numext::hypot as used at TridiagEigen.h:65 / DoubleShiftQR.h:123
__device__ __forceinline__ double eigen_hypot(double x, double y)
{
    x = fabs(x);
    y = fabs(y);
    const double p = fmax(x, y);
    if (p == 0.0)
        return 0.0;
    const double qp = fmin(y, x) / p;
    return p * sqrt(1.0 + qp * qp);
}

// Eigen::JacobiRotation::makeGivens (TridiagEigen.h:79-80, UpperHessenbergSchur.h:92,324):
// c*p - s*q = r, s*p + c*q = 0
__device__ __forceinline__ void make_givens(double p, double q, double& c, double& s, double* r = nullptr)
{
    if (q == 0.0)
    {
        c = p < 0.0 ? -1.0 : 1.0;
        s = 0.0;
        if (r)
            *r = fabs(p);
    }
    else if (p == 0.0)
    {
        c = 0.0;
        s = q < 0.0 ? 1.0 : -1.0;
        if (r)
            *r = fabs(q);
    }
    else if (fabs(p) > fabs(q))
    {
        const double t = q / p;
        double u = sqrt(1.0 + t * t);
        if (p < 0.0)
            u = -u;
        c = 1.0 / u;
        s = -t * c;
        if (r)
            *r = p * u;
    }
    else
    {
        const double t = p / q;
        double u = sqrt(1.0 + t * t);
        if (q < 0.0)
            u = -u;
        s = -1.0 / u;
        c = -t * s;
        if (r)
            *r = q * u;
    }
}

// StableScaling<double>::run (Givens.h:28-86): a >= b > 0
__device__ __forceinline__ void stable_scaling(double a, double b, double& r, double& c, double& s)
{
    const double t = b / a;
    const double cutoff = 0.1 * 1.220703125e-4;  // 0.1 * eps^(1/4), eps^(1/4) = 2^-13 exactly
    if (t >= cutoff)
    {
        r = hypot(a, b);
        c = a / r;
        s = b / r;
    }
    else
    {
        const double t2 = t * t;
        c = 1.0 - t2 * (0.5 - t2 * (0.375 - 0.3125 * t2));
        s = t * c;
        r = a + 0.5 * b * t * (1.0 - t2 * (0.25 - 0.125 * t2));
    }
}

// Givens<double>::compute_rotation (Givens.h:166-205): c*x - s*y = r, s*x + c*y = 0
__device__ __forceinline__ void givens_rotation(double x, double y, double& r, double& c, double& s)
{
    const double xsign = (x > 0.0) ? 1.0 : -1.0;
    const double xabs = fabs(x);
    if (y == 0.0)
    {
        c = (x == 0.0) ? 1.0 : xsign;
        s = 0.0;
        r = xabs;
        return;
    }
    const double ysign = (y > 0.0) ? 1.0 : -1.0;
    const double yabs = fabs(y);
    if (x == 0.0)
    {
        c = 0.0;
        s = -ysign;
        r = yabs;
        return;
    }
    if (xabs >= yabs)
    {
        stable_scaling(xabs, yabs, r, c, s);
        c = xsign * c;
        s = -ysign * s;
    }
    else
    {
        stable_scaling(yabs, xabs, r, s, c);
        c = xsign * c;
        s = -ysign * s;
    }
}

// SortingTarget<double, Rule>::get (SelectionRule.h:68-192); BothEnds sorts like LargestAlge.
__device__ __forceinline__ double sort_key_real(int rule, double v)
{
    switch (rule)
    {
        case SB200_LARGEST_MAGN: return -fabs(v);
        case SB200_LARGEST_ALGE:
        case SB200_BOTH_ENDS: return -v;
        case SB200_SMALLEST_MAGN: return fabs(v);
        default: return v;  // SB200_SMALLEST_ALGE
    }
}

// Index sort by key, ascending (SortEigenvalue, SelectionRule.h:195-224).  The reference uses
// std::sort (unspecified order of exactly equal keys); this is a stable insertion sort.
__device__ inline void argsort_keys(const double* key, int* idx, int len)
{
    for (int i = 0; i < len; i++)
        idx[i] = i;
    for (int i = 1; i < len; i++)
    {
        const int id = idx[i];
        const double k = key[id];
        int q = i - 1;
        while (q >= 0 && key[idx[q]] > k)
        {
            idx[q + 1] = idx[q];
            q--;
        }
        idx[q + 1] = id;
    }
}

}  // namespace dense
}  // namespace sb200

// ORACLE — TEST INFRASTRUCTURE ONLY.
// Shift-solve operator y = (A - sigma I)^{-1} x of SparseSymShiftSolve (MatOp/SparseSymShiftSolve.h:85-109).
//
// The reference delegates to Eigen::SparseLU (Eigen 3.4.0, a supernodal left-looking LU with partial pivoting and COLAMD
// ordering; third-party, NOT under /root/reference and not installed here).  Its factors and pivot order are not
// observable through the reference's API or tests -- only the solve result is (test/SymEigsShift.cpp:72-76, 1e-9 at
// solver level; SURVEY.md 8c item 4: "LU factors/pivot order: parity unpinned").  The oracle therefore restates the
// textbook algorithm for the matrix class of BASELINE config 5 (banded): LU with partial (row) pivoting in LAPACK band
// storage, i.e. the published DGBTF2 / DGBTRS algorithms, which computes the same mathematical result
// P (A - sigma I) = L U.  Cross-checked in tests/test_oracle.py against scipy.sparse.linalg.splu (SuperLU).
#pragma once

#include <algorithm>
#include <cmath>
#include <stdexcept>
#include <vector>

#include "solver.hpp"

namespace oracle {

struct BandLuOp
{
    Index n = 0, kl = 0, ku = 0, ldab = 0;
    std::vector<double> ab;   // AB(kl + ku + i - j, j) = A(i, j), column-major ldab x n, ldab = 2 kl + ku + 1
    std::vector<Index> ipiv;

    Index rows() const { return n; }
    double& AB(Index r, Index j) { return ab[size_t(r) + size_t(j) * size_t(ldab)]; }
    double AB(Index r, Index j) const { return ab[size_t(r) + size_t(j) * size_t(ldab)]; }

    // set_shift (SparseSymShiftSolve.h:85-95): mat = A - sigma I, factorise, throw invalid_argument on failure
    void set_shift(const CsrOp& A, double sigma)
    {
        n = A.n;
        kl = ku = 0;
        for (Index i = 0; i < n; i++)
            for (int64_t p = A.rowptr[i]; p < A.rowptr[i + 1]; p++)
            {
                const Index j = A.col[p];
                kl = std::max(kl, i - j);
                ku = std::max(ku, j - i);
            }
        ldab = 2 * kl + ku + 1;
        ab.assign(size_t(ldab) * size_t(n), 0.0);
        ipiv.assign(size_t(n), 0);
        const Index kv = kl + ku;
        for (Index i = 0; i < n; i++)
        {
            for (int64_t p = A.rowptr[i]; p < A.rowptr[i + 1]; p++)
                AB(kv + i - A.col[p], A.col[p]) += A.val[p];
            AB(kv, i) -= sigma;
        }
        // DGBTF2: unblocked band LU with partial pivoting
        Index ju = 0;
        for (Index j = 0; j < n; j++)
        {
            const Index km = std::min(kl, n - 1 - j);
            Index jp = 0;
            double best = std::fabs(AB(kv, j));
            for (Index r = 1; r <= km; r++)
                if (std::fabs(AB(kv + r, j)) > best)
                {
                    best = std::fabs(AB(kv + r, j));
                    jp = r;
                }
            ipiv[j] = j + jp;
            if (best == 0.0)
                throw std::invalid_argument("SparseSymShiftSolve: factorization failed with the given shift");
            ju = std::max(ju, std::min(j + ku + jp, n - 1));
            if (jp != 0)
                for (Index c = j; c <= ju; c++)
                    std::swap(AB(kv + jp - (c - j), c), AB(kv - (c - j), c));
            const double piv = 1.0 / AB(kv, j);
            for (Index r = 1; r <= km; r++)
                AB(kv + r, j) *= piv;
            for (Index c = j + 1; c <= ju; c++)
            {
                const double u = AB(kv - (c - j), c);
                if (u != 0.0)
                    for (Index r = 1; r <= km; r++)
                        AB(kv + r - (c - j), c) -= AB(kv + r, j) * u;
            }
        }
    }

    // perform_op (SparseSymShiftSolve.h:104-109): y = solve(x)   (DGBTRS, no transpose)
    void perform_op(const double* x, double* y) const
    {
        const Index kv = kl + ku;
        std::copy(x, x + n, y);
        for (Index j = 0; j < n; j++)
        {
            const Index lm = std::min(kl, n - 1 - j);
            const Index l = ipiv[j];
            if (l != j)
                std::swap(y[l], y[j]);
            const double bj = y[j];
            for (Index r = 1; r <= lm; r++)
                y[j + r] -= bj * AB(kv + r, j);
        }
        for (Index j = n - 1; j >= 0; j--)
        {
            y[j] /= AB(kv, j);
            const double bj = y[j];
            const Index lo = std::max<Index>(0, j - kv);
            for (Index i = lo; i < j; i++)
                y[i] -= bj * AB(kv - (j - i), j);
        }
    }
};

}  // namespace oracle

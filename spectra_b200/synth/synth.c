/*
 * Synthetic workload generator (host, C + OpenMP) for the benchmark configurations of
 * BASELINE.json / SURVEY.md §8(d).  Counter-based: every entry is a pure function of
 * (seed, row, slot), so any rank can generate any row slab and the CPU baseline sees exactly the
 * matrix the GPU path sees.
 *
 *   G_sym(n, d, seed):  A = sum_{t < d/2} ( P_t D_t + D_t P_t^T ),  P_t a pseudo-random permutation
 *       matrix (4-round Feistel network with cycle walking), D_t = diag(u_t), u_t ~ U(-0.5, 0.5)
 *       (the value distribution of the reference's gen_sparse_data, test/SymEigs.cpp:32,38).
 *       Row i holds (pi_t(i), u_t(i)) and (pi_t^{-1}(i), u_t(pi_t^{-1}(i))), i.e. exactly d
 *       uniformly scattered columns per row before duplicate merging; A is exactly symmetric.
 *   G_gen(n, d, seed):  row i holds (pi_t(i), u_t(i)) for t < d (no symmetrisation,
 *       as test/GenEigs.cpp:21-36).
 * Rows come out as CSR with ascending columns and duplicates summed.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

static inline uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

typedef struct
{
    uint64_t n;
    int half_bits;
    uint64_t half_mask;
    uint64_t key[4];
} perm_t;

static void perm_init(perm_t* p, uint64_t n, uint64_t seed, int t)
{
    int bits = 2;
    while ((1ULL << bits) < n)
        bits += 2; /* even number of bits, 2^bits >= n */
    p->n = n;
    p->half_bits = bits / 2;
    p->half_mask = (1ULL << p->half_bits) - 1;
    for (int r = 0; r < 4; r++)
        p->key[r] = splitmix64(seed * 0x100000001B3ULL + (uint64_t) t * 8191ULL + (uint64_t) r + 1);
}

static inline uint64_t feistel_fwd(const perm_t* p, uint64_t x)
{
    uint64_t l = x >> p->half_bits, r = x & p->half_mask;
    for (int k = 0; k < 4; k++)
    {
        const uint64_t f = splitmix64(r ^ p->key[k]) & p->half_mask;
        const uint64_t nl = r, nr = l ^ f;
        l = nl;
        r = nr;
    }
    return (l << p->half_bits) | r;
}

static inline uint64_t feistel_inv(const perm_t* p, uint64_t y)
{
    uint64_t l = y >> p->half_bits, r = y & p->half_mask;
    for (int k = 3; k >= 0; k--)
    {
        const uint64_t pr = l;
        const uint64_t f = splitmix64(pr ^ p->key[k]) & p->half_mask;
        const uint64_t pl = r ^ f;
        l = pl;
        r = pr;
    }
    return (l << p->half_bits) | r;
}

static inline uint64_t perm_fwd(const perm_t* p, uint64_t i)
{
    uint64_t y = feistel_fwd(p, i);
    while (y >= p->n)
        y = feistel_fwd(p, y);
    return y;
}

static inline uint64_t perm_inv(const perm_t* p, uint64_t i)
{
    uint64_t y = feistel_inv(p, i);
    while (y >= p->n)
        y = feistel_inv(p, y);
    return y;
}

static inline double uval(uint64_t seed, int t, uint64_t i)
{
    const uint64_t h = splitmix64(splitmix64(seed ^ 0xA5A5A5A5DEADBEEFULL) + (uint64_t) t * 0x9E3779B97F4A7C15ULL + i * 0xD1B54A32D192ED03ULL);
    return (double) (h >> 11) * (1.0 / 9007199254740992.0) - 0.5;
}

#define MAXD 256

/* generates row i into (c, v), sorted by column with duplicates summed; returns the length */
static int gen_row(const perm_t* perms, int d, int sym, uint64_t seed, uint64_t i, int32_t* c, double* v)
{
    int len = 0;
    if (sym)
    {
        const int h = d / 2;
        for (int t = 0; t < h; t++)
        {
            const uint64_t j = perm_fwd(&perms[t], i);
            c[len] = (int32_t) j;
            v[len] = uval(seed, t, i);
            len++;
            const uint64_t q = perm_inv(&perms[t], i);
            c[len] = (int32_t) q;
            v[len] = uval(seed, t, q);
            len++;
        }
    }
    else
    {
        for (int t = 0; t < d; t++)
        {
            c[len] = (int32_t) perm_fwd(&perms[t], i);
            v[len] = uval(seed, t, i);
            len++;
        }
    }
    /* insertion sort + merge */
    for (int a = 1; a < len; a++)
    {
        const int32_t cc = c[a];
        const double vv = v[a];
        int b = a - 1;
        while (b >= 0 && c[b] > cc)
        {
            c[b + 1] = c[b];
            v[b + 1] = v[b];
            b--;
        }
        c[b + 1] = cc;
        v[b + 1] = vv;
    }
    int w = 0;
    for (int a = 0; a < len; a++)
    {
        if (w > 0 && c[w - 1] == c[a])
            v[w - 1] += v[a];
        else
        {
            c[w] = c[a];
            v[w] = v[a];
            w++;
        }
    }
    return w;
}

/* Pass 1: rowptr[0..nrows] (offsets local to the slab).  Returns the slab's nnz, or -1 on bad input. */
int64_t synth_csr_count(int64_t n, int d, int sym, uint64_t seed, int64_t row0, int64_t nrows, int64_t* rowptr)
{
    if (d < 1 || d > MAXD || (sym && (d % 2)) || n < 2 || row0 < 0 || row0 + nrows > n)
        return -1;
    const int np = sym ? d / 2 : d;
    perm_t* perms = (perm_t*) malloc(sizeof(perm_t) * np);
    for (int t = 0; t < np; t++)
        perm_init(&perms[t], (uint64_t) n, seed, t);
    rowptr[0] = 0;
#pragma omp parallel
    {
        int32_t c[MAXD];
        double v[MAXD];
#pragma omp for schedule(static)
        for (int64_t r = 0; r < nrows; r++)
            rowptr[r + 1] = gen_row(perms, d, sym, seed, (uint64_t) (row0 + r), c, v);
    }
    for (int64_t r = 0; r < nrows; r++)
        rowptr[r + 1] += rowptr[r];
    free(perms);
    return rowptr[nrows];
}

/* Pass 2: fills col / val for the slab given pass 1's rowptr. */
int synth_csr_fill(int64_t n, int d, int sym, uint64_t seed, int64_t row0, int64_t nrows, const int64_t* rowptr, int32_t* col, double* val)
{
    if (d < 1 || d > MAXD || (sym && (d % 2)) || n < 2 || row0 < 0 || row0 + nrows > n)
        return -1;
    const int np = sym ? d / 2 : d;
    perm_t* perms = (perm_t*) malloc(sizeof(perm_t) * np);
    for (int t = 0; t < np; t++)
        perm_init(&perms[t], (uint64_t) n, seed, t);
#pragma omp parallel
    {
        int32_t c[MAXD];
        double v[MAXD];
#pragma omp for schedule(static)
        for (int64_t r = 0; r < nrows; r++)
        {
            const int len = gen_row(perms, d, sym, seed, (uint64_t) (row0 + r), c, v);
            memcpy(col + rowptr[r], c, sizeof(int32_t) * len);
            memcpy(val + rowptr[r], v, sizeof(double) * len);
        }
    }
    free(perms);
    return 0;
}

/* G_band(n, b, seed, diag_add): symmetric band matrix, A(i,j) = u(min(i,j), |i-j|) in U(-0.5, 0.5) for |i-j| <= b, plus diag_add on the
 * diagonal (SURVEY.md 8d "G_shift": symmetric banded, half-bandwidth 15, A - sigma I nonsingular).  Also the locality case of the SpMV
 * roofline study: every gathered x entry lies within b rows of the diagonal. */
int64_t synth_band_count(int64_t n, int b, int64_t row0, int64_t nrows, int64_t* rowptr)
{
    if (b < 0 || b > 127 || n < 2 || row0 < 0 || row0 + nrows > n)
        return -1;
    rowptr[0] = 0;
    for (int64_t r = 0; r < nrows; r++)
    {
        const int64_t i = row0 + r;
        const int64_t lo = i - b < 0 ? 0 : i - b, hi = i + b > n - 1 ? n - 1 : i + b;
        rowptr[r + 1] = rowptr[r] + (hi - lo + 1);
    }
    return rowptr[nrows];
}

int synth_band_fill(int64_t n, int b, uint64_t seed, double diag_add, int64_t row0, int64_t nrows, const int64_t* rowptr, int32_t* col, double* val)
{
    if (b < 0 || b > 127 || n < 2 || row0 < 0 || row0 + nrows > n)
        return -1;
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < nrows; r++)
    {
        const int64_t i = row0 + r;
        const int64_t lo = i - b < 0 ? 0 : i - b, hi = i + b > n - 1 ? n - 1 : i + b;
        int64_t p = rowptr[r];
        for (int64_t j = lo; j <= hi; j++, p++)
        {
            const int64_t mn = i < j ? i : j, off = i < j ? j - i : i - j;
            col[p] = (int32_t) j;
            val[p] = uval(seed, 1000 + (int) off, (uint64_t) mn) + (off == 0 ? diag_add : 0.0);
        }
    }
    return 0;
}

int synth_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// GenEigsSolver host driver — placeholder until the Arnoldi path lands (see solver_sym.cu).
#include "host.h"

struct sb200_gen_solver
{
    int dummy;
};

namespace sb200 {
static void nyi() { throw Error(SB200_RUNTIME, "GenEigsSolver device path is not built yet"); }
static sb200_stats g_empty_stats;
sb200_gen_solver* gen_create(sb200_op*, int64_t, int64_t) { nyi(); return nullptr; }
void gen_init(sb200_gen_solver*, const double*) { nyi(); }
int64_t gen_compute(sb200_gen_solver*, int, int64_t, double, int) { nyi(); return 0; }
void gen_factorize_from(sb200_gen_solver*, int64_t, int64_t) { nyi(); }
void gen_get_factorization(sb200_gen_solver*, double*, double*, double*, double*, int64_t*) { nyi(); }
int64_t gen_eigenvalues(const sb200_gen_solver*, double*) { nyi(); return 0; }
int64_t gen_eigenvectors(sb200_gen_solver*, int64_t, double*) { nyi(); return 0; }
int gen_info(const sb200_gen_solver*) { return SB200_NOT_COMPUTED; }
int64_t gen_niter(const sb200_gen_solver*) { return 0; }
int64_t gen_nops(const sb200_gen_solver*) { return 0; }
const sb200_stats& gen_stats(const sb200_gen_solver*) { return g_empty_stats; }
void gen_destroy(sb200_gen_solver* s) { delete s; }
void dense_hess_qr_host(int64_t, const double*, double, double*, double*) { nyi(); }
void dense_double_shift_qr_host(int64_t, const double*, double, double, double*, double*) { nyi(); }
void dense_hess_eigen_host(int64_t, const double*, double*, double*) { nyi(); }
}  // namespace sb200

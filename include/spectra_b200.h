/*
 * spectra_b200.h — C ABI of the B200-native implicitly restarted Lanczos/Arnoldi eigensolver.
 *
 * This is the drop-in boundary for the ONE hot path of yixuan/spectra:
 *   SymEigsSolver / GenEigsSolver restart loop + Sparse{Sym,Gen}MatProd::perform_op.
 * Plain pointers and sizes only; no C++/torch types.  All matrices are column-major
 * (Eigen default), all indices 0-based.  Every entry point returns an sb200_status; the text of
 * the last error on the calling thread is available from sb200_last_error().
 *
 * The C++ shim in include/Spectra/ (same class / method names as the reference) and the Python
 * mirror in spectra_b200/ are thin wrappers over these functions.  Citations (file:line) are
 * relative to /root/reference/include/Spectra/ and name the reference interface each entry
 * point replaces.
 *
 * There is NO CPU fallback: every compute entry point needs a CUDA device (sm_100a) and fails
 * with SB200_CUDA otherwise.
 */
#ifndef SPECTRA_B200_H
#define SPECTRA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes; the shim maps them onto the reference's exception types ---- */
typedef enum
{
    SB200_OK = 0,
    SB200_INVALID_ARGUMENT = 1, /* std::invalid_argument (HermEigsBase.h:267-271, Arnoldi.h:147-148, ...) */
    SB200_LOGIC = 2,            /* std::logic_error   (e.g. UpperHessenbergQR.h:206-207) */
    SB200_RUNTIME = 3,          /* std::runtime_error (TridiagEigen.h:203-204, UpperHessenbergSchur.h:421-422) */
    SB200_CUDA = 4,             /* CUDA runtime failure / no device */
    SB200_NCCL = 5              /* NCCL failure */
} sb200_status;

/* Util/SelectionRule.h:33-58 — same order as the reference's enum class SortRule */
typedef enum
{
    SB200_LARGEST_MAGN = 0,
    SB200_LARGEST_REAL,
    SB200_LARGEST_IMAG,
    SB200_LARGEST_ALGE,
    SB200_SMALLEST_MAGN,
    SB200_SMALLEST_REAL,
    SB200_SMALLEST_IMAG,
    SB200_SMALLEST_ALGE,
    SB200_BOTH_ENDS
} sb200_sort_rule;

/* Util/CompInfo.h:17-30 */
typedef enum
{
    SB200_SUCCESSFUL = 0,
    SB200_NOT_COMPUTED,
    SB200_NOT_CONVERGING,
    SB200_NUMERICAL_ISSUE
} sb200_comp_info;

/* How the stored sparse matrix is interpreted (template parameters of the MatProd wrappers). */
typedef enum
{
    SB200_COL_MAJOR = 0, /* Eigen::ColMajor: outer = column (CSC)  — the reference default */
    SB200_ROW_MAJOR = 1  /* Eigen::RowMajor: outer = row    (CSR) */
} sb200_storage_order;

typedef enum
{
    SB200_GENERAL = 0,   /* SparseGenMatProd: every stored entry is used          (SparseGenMatProd.h:82-87) */
    SB200_SYM_LOWER = 1, /* SparseSymMatProd<.., Eigen::Lower>: selfadjointView   (SparseSymMatProd.h:83-88) */
    SB200_SYM_UPPER = 2, /* SparseSymMatProd<.., Eigen::Upper> */
    SB200_HERM_LOWER = 3, /* SparseHermMatProd<std::complex<double>, Eigen::Lower> (SparseHermMatProd.h:83-88); complex operators only */
    SB200_HERM_UPPER = 4  /* SparseHermMatProd<std::complex<double>, Eigen::Upper> */
} sb200_matrix_mode;

typedef struct sb200_comm sb200_comm;
typedef struct sb200_op sb200_op;
typedef struct sb200_sym_solver sb200_sym_solver;
typedef struct sb200_gen_solver sb200_gen_solver;

const char* sb200_last_error(void);
/* Library / device facts: returns SB200_CUDA when no sm_100-class device is usable. */
int sb200_device_info(int* device, int* sm_count, int* cc_major, int* cc_minor, int64_t* hbm_bytes);
const char* sb200_version(void);
/* Selects the CUDA device of the calling process (one process per GPU); must precede every other call. */
int sb200_set_device(int device);

/* ------------------------------------------------------------------------------------------
 * Communicator — row-sharded multi-GPU runs, one process per GPU.  The reference
 * has no distributed layer; this is new surface.  The launcher creates the 128-byte id on rank
 * 0 (sb200_comm_unique_id), distributes it by any means (torch.distributed / MPI / file) and
 * every rank calls sb200_comm_create.
 * ------------------------------------------------------------------------------------------ */
int sb200_comm_unique_id(void* id128);
int sb200_comm_create(int rank, int nranks, const void* id128, sb200_comm** out);
int sb200_comm_rank(const sb200_comm* c, int* rank, int* nranks);
int sb200_comm_destroy(sb200_comm* c);

/* ------------------------------------------------------------------------------------------
 * Operator — replaces MatOp/SparseSymMatProd.h:30-105 and MatOp/SparseGenMatProd.h:29-104.
 * The compressed arrays (outer[n+1], inner[nnz], values[nnz]) are host pointers with the layout
 * of Eigen::SparseMatrix<double, Flags, int> in compressed mode; they are uploaded ONCE and
 * turned into a full device-resident CSR (symmetric modes read only the named triangle and
 * mirror it, exactly like selfadjointView<Uplo>).  outer may be int32 (outer_is_64 = 0, Eigen's
 * StorageIndex=int) or int64.
 * With comm != NULL the operator keeps only this rank's contiguous block of rows.
 * ------------------------------------------------------------------------------------------ */
int sb200_op_create_sparse(int64_t n, const void* outer, int outer_is_64, const int32_t* inner, const double* values, int storage_order, int matrix_mode,
                           sb200_comm* comm, sb200_op** out);
/* Pre-partitioned form: this rank's rows [row0, row0 + nrows) of a full n x n CSR (general). */
int sb200_op_create_csr_slab(int64_t n, int64_t row0, int64_t nrows, const int64_t* rowptr_local, const int32_t* col, const double* values, sb200_comm* comm,
                             sb200_op** out);
/* User-defined operator: the reference's OpType concept (SymEigsSolver.h:99-114, MIGRATION.md:19-37) — any
 * host function computing y_out = A * x_in on n-vectors.  The Krylov basis, the re-orthogonalisation and the
 * restart still run on the GPU; each matrix operation copies v to pinned host memory, calls fn, and copies
 * the product back.  Single-GPU only. */
typedef void (*sb200_matvec_fn)(const double* x_in, double* y_out, void* user);
int sb200_op_create_callback(int64_t n, sb200_matvec_fn fn, void* user, sb200_op** out);
/* The same for a user-defined COMPLEX operator (OpType::Scalar = std::complex<double>, for sb200_herm_create): fn receives and fills
 * interleaved (re, im) vectors of 2 n doubles. */
int sb200_op_create_callback_z(int64_t n, sb200_matvec_fn fn, void* user, sb200_op** out);
/* Complex Hermitian operator: replaces MatOp/SparseHermMatProd.h:21-89 (Scalar = std::complex<double>).  values_ri holds the
 * nnz complex values interleaved (re, im) -- the memory layout of std::complex<double> --, matrix_mode is SB200_HERM_LOWER /
 * SB200_HERM_UPPER (the Uplo template argument: only that triangle is read, mirrored conjugated, the diagonal taken as real) or
 * SB200_GENERAL.  Every vector the operator or its solver exchanges is interleaved complex too: perform_op / apply_matrix take
 * and return 2 n doubles per column.  Single GPU.  (device-verified in round 2: tests/test_gpu_layouts_complex.py.) */
int sb200_op_create_sparse_herm(int64_t n, const void* outer, int outer_is_64, const int32_t* inner, const double* values_ri, int storage_order, int matrix_mode,
                                sb200_op** out);
/* Shift-solve operator: replaces MatOp/SparseSymShiftSolve.h:30-110.  Same matrix arguments as sb200_op_create_sparse
 * (matrix_mode SB200_SYM_LOWER / SB200_SYM_UPPER = the Uplo template argument).  perform_op then computes
 * y = (A - sigma I)^{-1} x.  Device implementation, chosen from the half-bandwidth b of the pattern: block cyclic reduction on the
 * block-tridiagonal form for b <= 32 (BASELINE config 5's class); sequential block elimination with grid-wide block kernels for wider
 * bands and mesh-like patterns (2-D / 3-D stencils in natural ordering), as long as the 3 n B doubles of block factors fit in device
 * memory; an explicit inverse for n <= 2048.  A large pattern without band structure returns SB200_INVALID_ARGUMENT.  Single-GPU. */
int sb200_op_create_shift_solve(int64_t n, const void* outer, int outer_is_64, const int32_t* inner, const double* values, int storage_order, int matrix_mode,
                                sb200_op** out);
/* set_shift(sigma) (SparseSymShiftSolve.h:85-95): factorises A - sigma I on the device.  SB200_INVALID_ARGUMENT
 * ("factorization failed with the given shift", :93-94) when a pivot block is singular or the verification solve fails. */
int sb200_op_set_shift(sb200_op* op, double sigma);
/* layout chosen for the factorisation: half-bandwidth found, block size B, number of block rows, reduction levels (block cyclic
 * reduction, half-bandwidth <= 32) or -1 (sequential block elimination with grid-wide block kernels: wider bands, mesh-like patterns) */
int sb200_op_shift_solve_info(const sb200_op* op, int* half_bandwidth, int* block, int64_t* block_rows, int* levels);
/* iterative-refinement steps per solve: 0 or 1 fixes it; a negative value (the default) lets set_shift() decide -- the refinement sweep is
 * dropped when the plain solve of its verification right-hand side already has a relative residual <= 5e-14 */
int sb200_op_shift_solve_refine(sb200_op* op, int steps);
/* outcome of the last set_shift(): refinement steps in use, relative residual of the verification solve as the solves now run, and of
 * the unrefined solve (-1 when it was not measured) */
int sb200_op_shift_solve_status(const sb200_op* op, int* refine_steps, double* verify_residual, double* unrefined_residual);
int sb200_op_rows(const sb200_op* op, int64_t* rows);      /* rows()  SparseSymMatProd.h:70 */
int sb200_op_cols(const sb200_op* op, int64_t* cols);      /* cols()  SparseSymMatProd.h:74 */
int sb200_op_local_rows(const sb200_op* op, int64_t* row0, int64_t* nrows);
int sb200_op_nnz(const sb200_op* op, int64_t* nnz_local);
/* perform_op(x_in, y_out) with the reference's HOST-pointer semantics (SparseSymMatProd.h:83-88):
 * x (n) is copied to the device, the CSR SpMV kernel runs, y (n, or this rank's rows when sharded)
 * is copied back. */
int sb200_op_perform_op(sb200_op* op, const double* x_host, double* y_host);
/* operator*(Matrix) (SparseSymMatProd.h:93-96): Y = A * X for an n x k column-major X. */
int sb200_op_apply_matrix(sb200_op* op, const double* X_host, int64_t k, double* Y_host);
/* Device-pointer form used internally and by benchmarks: x_dev holds n doubles (full vector),
 * y_dev the local rows.  Runs on the operator's stream; *elapsed_ms (optional) gets the CUDA-event
 * time of `repeat` back-to-back launches.  x_dev / y_dev may be NULL: the call then uses an internal
 * vector of ones / scratch output (benchmark convenience). */
int sb200_op_spmv_device(sb200_op* op, const double* x_dev, double* y_dev, int repeat, float* elapsed_ms);
int sb200_op_destroy(sb200_op* op);
/* Device layout of a sparse operator (new surface, for tests and benchmarks): *format = 0 CSR (sub-warp per row kernels), 1 sliced
 * CSR (one lane per row; the default whenever its padding is below SB200_SELL_MAX_FILL, SB200_SPMV_FORMAT=csr forces the CSR-vector kernels); *col_blocks = column blocks the operand is
 * split into; *stored_entries = matrix entries held on the device including padding. */
int sb200_op_layout_info(const sb200_op* op, int* format, int* col_blocks, int64_t* stored_entries);
/* Row-sharded operators (new surface): *peer = 1 when the ranks exchange the SpMV operand and the dot products through NVLink-mapped
 * peer memory (CUDA IPC windows: residual rows written by the correction pass, one-shot mailbox all-reduce), 0 when they use the NCCL
 * collectives (single rank, SB200_PEER=0, or peers that cannot be mapped). */
int sb200_op_peer_mode(const sb200_op* op, int* peer);
/* Roofline microbenchmark (tools/gather_roof.py; not on the product path): average time of `gathers` independent, uniformly random
 * 8-byte read-only loads from a device vector of n doubles -- the operand access of a CSR SpMV with random column ids, without the
 * matrix stream.  *checksum = mean of the loaded values (1.0). */
int sb200_bench_gather(int64_t n, int64_t gathers, int repeat, float* elapsed_ms, double* checksum);
/* The same with the indices and one coefficient per gather STREAMED from HBM (12 B per gather, coalesced, four steps in flight) and
 * each gather dependent on its index load: acc += val[k] * x[col[k]] without rows, padding or output -- the floor of any SpMV on this
 * access pattern.  band = 0: uniformly random columns; band = 1: neighbouring columns (coalesced gathers). */
int sb200_bench_stream_gather(int64_t n, int64_t gathers, int band, int repeat, float* elapsed_ms, double* checksum);

/* ------------------------------------------------------------------------------------------
 * SymEigsSolver — replaces SymEigsSolver.h:133-160 + HermEigsBase.h:43-479 (+ the
 * back-transform of SymEigsShiftSolver.h:163-169 when created with sb200_sym_create_shift on a
 * shift-solve operator).
 * ------------------------------------------------------------------------------------------ */
int sb200_sym_create(sb200_op* op, int64_t nev, int64_t ncv, sb200_sym_solver** out);      /* ctor, HermEigsBase.h:257-272 */
/* SymEigsShiftSolver(op, nev, ncv, sigma) (SymEigsShiftSolver.h:190-195): calls set_shift(sigma) on a device shift-solve
 * operator (a callback operator must already apply (A - sigma I)^{-1}); eigenvalues are mapped back by
 * lambda = 1/nu + sigma before sorting (:163-169).  All other calls are the sb200_sym_* functions. */
int sb200_sym_create_shift(sb200_op* op, int64_t nev, int64_t ncv, double sigma, sb200_sym_solver** out);
/* HermEigsSolver(op, nev, ncv) (HermEigsSolver.h:121-122 + HermEigsBase.h with a complex Scalar) on an operator made by
 * sb200_op_create_sparse_herm; ncv <= 63.  The handle is used with the sb200_sym_* calls: eigenvalues are real, init() takes and
 * eigenvectors() returns interleaved complex data (2 n doubles per vector / column). */
int sb200_herm_create(sb200_op* op, int64_t nev, int64_t ncv, sb200_sym_solver** out);
int sb200_sym_init(sb200_sym_solver* s, const double* init_resid_or_null);                  /* init(), init(const Scalar*) :309-342 */
int sb200_sym_compute(sb200_sym_solver* s, int selection, int64_t maxit, double tol, int sorting, int64_t* nconv); /* compute() :366-390 */
int sb200_sym_info(const sb200_sym_solver* s, int* info);                                   /* info() :396 */
int sb200_sym_num_iterations(const sb200_sym_solver* s, int64_t* niter);                    /* :401 */
int sb200_sym_num_operations(const sb200_sym_solver* s, int64_t* nops);                     /* :406 */
/* eigenvalues() :417-436 — out must hold nev doubles; *count = number of converged values */
int sb200_sym_eigenvalues(const sb200_sym_solver* s, double* out, int64_t* count);
/* eigenvectors(nvec) :447-470 — out is n x min(nvec, nconv) column-major host memory (full n rows;
 * sharded runs gather the slabs).  *ncols = columns written. */
int sb200_sym_eigenvectors(sb200_sym_solver* s, int64_t nvec, double* out, int64_t* ncols);
/* Sharded runs: only this rank's rows (nrows_local x ncols). */
int sb200_sym_eigenvectors_local(sb200_sym_solver* s, int64_t nvec, double* out, int64_t* ncols);
int sb200_sym_destroy(sb200_sym_solver* s);

/* Instrumentation (new surface): device time split and algorithm counters of the last compute(). */
typedef struct
{
    int64_t lanczos_steps;  /* factorize_from loop bodies */
    int64_t reorth_passes;  /* correction passes (Lanczos.h:156-182) */
    int64_t restarts;       /* restart() calls (HermEigsBase.h:105-155) */
    int64_t expand_calls;   /* expand_basis() calls (Arnoldi.h:66-115) */
    int64_t kernel_launches;/* kernels launched by init()+compute() */
    int64_t spmv_launches;  /* CSR SpMV kernels (plain + fused step head) */
    int64_t panel_launches; /* fused Krylov-panel passes */
    int64_t panel_cols;     /* sum of the panel widths j over those passes (algorithmic bytes = 8 n (panel_cols + 2 panel_launches)) */
    int64_t compress_launches; /* restart GEMMs */
    int64_t compress_cols;  /* sum of output widths k+1 over those GEMMs */
    double ms_total;        /* device time of init()+compute() (CUDA events) */
    double ms_spmv;         /* of which: CSR SpMV kernels (only when profiling enabled) */
    double ms_panel;        /* fused re-orthogonalisation panel kernels */
    double ms_compress;     /* compress_V GEMM */
    double ms_small;        /* small dense restart kernels */
    double ms_comm;         /* collectives (NCCL or peer-memory kernels) */
    int64_t fused_dot_launches; /* operator applications whose last kernel also carried the first panel pass (sell_step_dot_kernel) */
    int64_t fused_dot_cols;     /* sum of the panel widths i streamed by those kernels: their extra algorithmic bytes are 8 n fused_dot_cols */
    int64_t host_syncs;         /* stream synchronisations issued by init()+compute() */
} sb200_stats;
int sb200_sym_stats(const sb200_sym_solver* s, sb200_stats* out);
/* 0 = off (default, no extra events), 1 = per-kernel-class CUDA-event timing (adds syncs). */
int sb200_set_profiling(int level);

/* ---- test hooks: the factorisation tier of the reference's tests (test/Arnoldi.cpp:19-85) ---- */
/* Lanczos::factorize_from(from_k, to_m) on the device state (Lanczos.h:62-187). */
int sb200_sym_factorize_from(sb200_sym_solver* s, int64_t from_k, int64_t to_m);
/* Copies V (n_local x ncv), H (ncv x ncv), f (n_local) and beta to host buffers (any may be NULL). */
int sb200_sym_get_factorization(sb200_sym_solver* s, double* V, double* H, double* f, double* beta, int64_t* k);

/* ------------------------------------------------------------------------------------------
 * GenEigsSolver — replaces GenEigsSolver.h:158-186 + GenEigsBase.h:43-612 (real double).
 * Complex results are returned as interleaved (re, im) pairs.
 * ------------------------------------------------------------------------------------------ */
/* With a COMPLEX operator (sb200_op_create_sparse_herm in SB200_GENERAL mode, sb200_op_create_callback_z; ncv <= 63) this is
 * GenEigsSolver with Scalar = std::complex<double> (GenEigsBase.h:111-140, test/ComplexEigs.cpp): the initial residual, the
 * factorisation (sb200_gen_get_factorization: V, H, f) and the eigenvectors are interleaved complex.  Experimental in round 1:
 * verified on the kernel-logic emulator, not yet on a device. */
int sb200_gen_create(sb200_op* op, int64_t nev, int64_t ncv, sb200_gen_solver** out);      /* GenEigsBase.h:409-424 */
int sb200_gen_init(sb200_gen_solver* s, const double* init_resid_or_null);                  /* :442-475 */
int sb200_gen_compute(sb200_gen_solver* s, int selection, int64_t maxit, double tol, int sorting, int64_t* nconv); /* :501-525 */
int sb200_gen_info(const sb200_gen_solver* s, int* info);
int sb200_gen_num_iterations(const sb200_gen_solver* s, int64_t* niter);
int sb200_gen_num_operations(const sb200_gen_solver* s, int64_t* nops);
int sb200_gen_eigenvalues(const sb200_gen_solver* s, double* out_ri, int64_t* count);        /* :531-551 */
int sb200_gen_eigenvectors(sb200_gen_solver* s, int64_t nvec, double* out_ri, int64_t* ncols); /* :561-603 */
int sb200_gen_stats(const sb200_gen_solver* s, sb200_stats* out);
int sb200_gen_factorize_from(sb200_gen_solver* s, int64_t from_k, int64_t to_m);             /* Arnoldi.h:198-295 */
int sb200_gen_get_factorization(sb200_gen_solver* s, double* V, double* H, double* f, double* beta, int64_t* k);
int sb200_gen_destroy(sb200_gen_solver* s);

/* ------------------------------------------------------------------------------------------
 * Small dense device kernels, exposed for the unit tier of the reference's tests
 * (test/QR.cpp, test/Eigen.cpp, test/Schur.cpp).  Host buffers in / out; each call runs the
 * single-CTA device kernel the solvers use.
 * ------------------------------------------------------------------------------------------ */
/* TridiagEigen::compute (TridiagEigen.h:121-210): evals (m, unsorted) and evecs (m x m). */
int sb200_dense_tridiag_eigen(int64_t m, const double* H, double* evals, double* evecs);
/* Givens<double>::compute_rotation (Givens.h:166-205, StableScaling :28-86; sign convention test/Givens.cpp:82-95): `count` independent
 * rotations, (r, c, s) with c*x - s*y = r, s*x + c*y = 0.  variant 0 = the reference's formulas (Taylor branch included),
 * 1 = the rsqrt form the device QR kernels call, 2 = Eigen's JacobiRotation::makeGivens (TridiagEigen.h:79-80). */
int sb200_dense_givens(int variant, int64_t count, const double* x, const double* y, double* r, double* c, double* s);
/* TridiagQR (kind 0, UpperHessenbergQR.h:459-709) / UpperHessenbergQR (kind 1, :46-447):
 * QtHQ = Q'HQ and Q = G1*G2*... for H - shift*I = QR. */
int sb200_dense_shifted_qr(int kind, int64_t m, const double* H, double shift, double* QtHQ, double* Q);
/* DoubleShiftQR (DoubleShiftQR.h:20-438) for H^2 - s*H + t*I. */
int sb200_dense_double_shift_qr(int64_t m, const double* H, double s, double t, double* QtHQ, double* Q);
/* UpperHessenbergEigen (UpperHessenbergEigen.h:32-321): interleaved complex evals (m) and evecs (m x m). */
int sb200_dense_hess_eigen(int64_t m, const double* H, double* evals_ri, double* evecs_ri);
/* Complex counterparts (interleaved (re, im) m x m matrices, m <= 63) for the complex GenEigsSolver:
 * UpperHessenbergQR<std::complex<double>> (UpperHessenbergQR.h:136-255, 383-417 with the complex Givens of Givens.h:218-335) and
 * UpperHessenbergEigen<std::complex<double>> (UpperHessenbergEigen.h:328-454; unit-norm eigenvectors, unsorted). */
int sb200_dense_shifted_qr_z(int64_t m, const double* H_ri, double shift_re, double shift_im, double* QtHQ_ri, double* Q_ri);
int sb200_dense_hess_eigen_z(int64_t m, const double* H_ri, double* evals_ri, double* evecs_ri);
/* One restart "prepare" step of HermEigsBase (retrieve_ritzpair :205-224, num_converged :158-175,
 * nev_adjusted :178-202, shift loop :118-147) on a tridiagonal H and beta. */
int sb200_dense_sym_restart(int64_t m, const double* H, double beta, int64_t nev, int selection, double tol, double* ritz_val, double* ritz_est,
                            int32_t* conv, int64_t* nconv, int64_t* k, double* Q, double* Hnew);
/* Restart GEMM of Arnoldi::compress_V (Arnoldi.h:320-340) on host buffers: Vout (n x kk) = V (n x m) * Q[:, :kk] (Q m x m,
 * col-major).  If f != NULL (kk >= 2): f <- f*Q(m-1,kk-2) + Vout[:,kk-1]*H(kk-1,kk-2) (:337) and *fnorm2 = ||f||^2.
 * impl 0 = DMMA/TMA kernel, 1 = FMA kernel (both are product kernels; the solvers use 0). */
int sb200_dense_compress(int64_t n, int64_t m, int64_t kk, const double* V, const double* Q, const double* H, double* Vout, double* f, double* fnorm2,
                         int impl);

#ifdef __cplusplus
}
#endif
#endif /* SPECTRA_B200_H */

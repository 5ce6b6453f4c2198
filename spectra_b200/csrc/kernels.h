// Host-side launch interface of the sm_100a kernels (implemented in spmv.cu, panel.cu,
// vecops.cu, csr_build.cu, dense_sym.cu, dense_gen.cu).
#pragma once

#include "common.cuh"

namespace sb200 {

constexpr int kPanelMaxCols = 64;   // widest Krylov panel one fused pass handles
constexpr int kRedStride = 128;     // slots per CTA in the reduction scratch
constexpr int kRedNrm = 64;         // red[kRedNrm] = ||f||^2 of the last panel pass

// Device-resident control block of one running Lanczos/Arnoldi factorisation.  All data-dependent
// scalars of the inner loop live here so that kernels chain on the stream without host round trips.
struct FacCtl
{
    // ---- status block (first 64 bytes; the host reads exactly this part after a decide kernel) ----
    double beta;                  // ||f||                                   (Arnoldi.h:61 m_beta)
    double ortho_err;             // max |V^T f|                              (Lanczos.h:153)
    double hnorm;                 // ||h|| of the Arnoldi step                (Arnoldi.h:257)
    double hsub;                  // H(i, i-1) of the step being built (beta, or 0 after a restart)
    int i;                        // column being built
    int count;                    // correction passes done in this step      (Lanczos.h:155)
    int need_corr;                // another correction pass is required      (Lanczos.h:156)
    int f_zeroed;                 // beta < eps*sqrt(n): f forced to zero     (Lanczos.h:163-168)
    int dgks_skip;                // Arnoldi: beta > 0.717 ||h||, no re-orth  (Arnoldi.h:257)
    int abort;                    // sweep mode: a step needs the host (extra correction, zeroed residual, tiny beta) -- every later
                                  // kernel of the enqueued sweep returns at once; `i` names the step that raised it
    int acc_count;                // Arnoldi sweeps: correction passes of the steps completed on the device since the sweep began
    int acc_skipped;              //   speculatively enqueued passes that turned out unnecessary (returned in their prologue)
    int acc_skipped_cols;         //   and the sum of their panel widths (host-side traffic accounting)
    int pad[3];
    // ---- reduction outputs / coefficients -------------------------------------------------------
    double red_a[8];              // SpMV-epilogue reduction: [0] = <v_i, w>  (Lanczos.h:142)
    double c[kRedStride];         // coefficients applied by the next correction pass  (Vf, Lanczos.h:152)
    double red[kRedStride];       // output of the last panel reduction: [0..j) = V^T f, [kRedNrm] = ||f||^2
    double red2[kRedStride];      // the same for the first row range of a correction pass run in two parts (added by the decide kernel)
};
constexpr size_t kFacCtlStatusBytes = 80;

// Largest grid a solver's reduction scratch serves: the persistent kernels use at most sm_count * 16 CTAs; the sliced-layout step
// kernel runs one CTA per 1024-row window up to this bound (16384 windows = 16.7 M rows per rank) and turns persistent beyond it.
inline int reduction_max_grid(int sm_count) { return sm_count * 16 > 16384 ? sm_count * 16 : 16384; }

// Reduction scratch shared by all kernels of one solver (one stream => no overlap).
struct RedScratch
{
    double* partials;       // max_grid * kRedStride doubles
    unsigned int* ticket;   // zero-initialised
    int max_grid;
};

// ---- CSR build (csr_build.cu) -----------------------------------------------------------------
constexpr int kMaxColBlocks = 16;

// Sliced layout of one CSR (block) for the lane-per-row SpMV kernels (spmv.cu, "sliced CSR" / SELL-32-sigma):
// rows are grouped in windows of kSellWindow consecutive rows; inside a window the rows are ordered by
// descending length (stable) and cut into slices of 32 rows; a slice stores its entries step-major
// (entry t of the 32 rows side by side), padded to the longest row of the slice with column id -1.
// Lane r of a warp therefore owns one row and reads col / val with fully coalesced 128 B / 256 B warp
// loads, sums its row sequentially in ascending column order (the order a scalar CSR loop uses) and needs
// no cross-lane reduction.  perm maps (window, position after sorting) -> row inside the window, so that the
// kernels can hand their results back in natural row order through shared memory.
// Experimental in round 1 (SB200_SPMV_FORMAT=sell): built next to the CSR arrays, which stay authoritative.
constexpr int kSellWindow = 1024;
constexpr int kSellSlice = 32;
struct SellBlock
{
    int64_t nwin = 0;              // windows = ceil(nrows / kSellWindow)
    int64_t padded = 0;            // stored entries including padding (multiple of 32)
    DevBuf<int> slice_ptr;         // nwin * 32 + 1 entry offsets
    DevBuf<int> col;               // padded; -1 marks padding
    DevBuf<double> val;            // padded
    DevBuf<unsigned short> perm;   // nwin * kSellWindow
    bool built() const { return nwin > 0; }
    void release()
    {
        nwin = 0;
        padded = 0;
        slice_ptr.release();
        col.release();
        val.release();
        perm.release();
    }
};

// One column block of the operand: a CSR of its own over columns [c0, c1).
struct CsrBlock
{
    int64_t nnz = 0;
    DevBuf<int> rowptr;   // nrows + 1
    DevBuf<int> col;      // nnz (global column ids)
    DevBuf<double> val;   // nnz
    SellBlock sell;       // optional sliced copy (build_sell_layout)
};
struct DeviceCsr
{
    int64_t n = 0;        // global order
    int64_t row0 = 0;     // first local row
    int64_t nrows = 0;    // local rows
    int64_t nnz = 0;      // local nnz
    DevBuf<int> rowptr;   // nrows + 1
    DevBuf<int> col;      // nnz (global column ids)
    DevBuf<double> val;   // nnz
    // Column blocking (large n): when the gathered operand x (8n bytes) does not fit in L2, the matrix is
    // re-laid out as `blocks.size()` CSR sub-matrices of `col_block_width` columns each, processed one after
    // the other so that each pass gathers from an L2-resident slice of x.  Empty => single block above.
    std::vector<CsrBlock> blocks;
    int64_t col_block_width = 0;
    SellBlock sell;       // optional sliced copy of the unblocked CSR (build_sell_layout)
    // Chunked column layout (row-sharded operators; chunk_len == 0 means the natural layout above).  Column block c holds, for
    // every rank r, the columns [r*chunk_slab + c*chunk_len, +chunk_len) with their ids REMAPPED to r*chunk_len + i, i.e. to
    // positions inside the result of the c-th partial all-gather (chunk_ranks*chunk_len doubles).  The SpMV of block c can
    // therefore start as soon as all-gather c has landed while all-gather c+1 is still in flight.
    int64_t chunk_slab = 0, chunk_len = 0;
    int chunk_ranks = 0;
    int64_t chunk_stride() const { return (int64_t) chunk_ranks * chunk_len; }
    const double* x_of_block(const double* x, int c) const { return chunk_len ? x + (int64_t) c * chunk_stride() : x; }
};
// Complex operand of the Hermitian path: full CSR with interleaved complex values, single GPU, one column block.
struct DeviceCsrZ
{
    int64_t n = 0;
    int64_t nnz = 0;
    DevBuf<int> rowptr;      // n + 1
    DevBuf<int> col;         // nnz
    DevBuf<double2> val;     // nnz, (re, im)
};
void build_device_csr_z(int64_t n, const void* outer, bool outer64, const int32_t* inner, const double* values_ri, int order, int mode, cudaStream_t stream,
                        DeviceCsrZ& out);
// y (n complex, interleaved) = A x   (MatOp/SparseHermMatProd.h:83-88)
void launch_spmv_z(const DeviceCsrZ& A, const double* x_ri, double* y_ri, cudaStream_t stream);
// Re-lays the CSR out in nblocks column blocks (frees the unblocked arrays).  No-op for nblocks <= 1.
void split_column_blocks(DeviceCsr& A, int nblocks, cudaStream_t stream);
// Chunked variant for row-sharded operators (see DeviceCsr::chunk_len): nchunks blocks, chunk_len = ceil(slab / nchunks) rounded
// up to 16 entries.
void split_column_chunks(DeviceCsr& A, int nchunks, int64_t slab, int nranks, cudaStream_t stream);
// x_chunked (nchunks * nranks * chunk_len) <- natural-layout x_nat (A.n entries); padding positions are zeroed
void launch_permute_to_chunks(const DeviceCsr& A, const double* x_nat, double* x_chunked, cudaStream_t stream);
// Adds the sliced layout (SellBlock) to the CSR / to every column block.  Returns false (and builds nothing) when padding
// would store more than max_fill times the nonzeros (very uneven row lengths: the CSR-vector kernels stay in charge).
bool build_sell_layout(DeviceCsr& A, double max_fill, cudaStream_t stream);
// Builds the full CSR (columns ascending in each row, duplicates summed) from host compressed
// arrays.  mode/order as in sb200_matrix_mode / sb200_storage_order.  Keeps rows [row0,row0+nrows).
void build_device_csr(int64_t n, const void* outer, bool outer64, const int32_t* inner, const double* values, int order, int mode, int64_t row0, int64_t nrows,
                      cudaStream_t stream, DeviceCsr& out);
void upload_csr_slab(int64_t n, int64_t row0, int64_t nrows, const int64_t* rowptr_local, const int32_t* col, const double* values, cudaStream_t stream,
                     DeviceCsr& out);

// ---- SpMV (spmv.cu) -----------------------------------------------------------------------------
struct SpmvPlan
{
    int lanes = 8;  // lanes per row
    int grid = 0;
    int sell_threads = 0;  // > 0: the operand carries the sliced layout and the lane-per-row kernels run (CTA size)
    int sell_grid = 0;        // fused step kernel (bounded by the reduction scratch)
    int sell_grid_plain = 0;  // plain / accumulate kernels: one CTA per window
    int sell_grid_fused = 0;  // fused step head + first panel pass (sell_step_dot_kernel); 0 = not used
};
SpmvPlan make_spmv_plan(const DeviceCsr& A);
// column blocks needed so that one slice of the gathered operand stays L2-resident (env SB200_XSLICE_MB)
int choose_col_blocks(int64_t n);
// y = A x   (x: full vector of A.n entries, y: A.nrows)
void launch_spmv(const DeviceCsr& A, const SpmvPlan& plan, const double* x, double* y, cudaStream_t stream);
// Fused Lanczos/Arnoldi step head (K-A):  V[:,i] = f_loc/beta;  w = (A x)/beta - hsub*V[:,i-1]
// (sym only);  red_a[0] = <V[:,i], w> (sym only).   x is the full (un-normalised) residual.
// V points at column 0 (ld = ldv).  Reads i, beta, hsub from ctl.
// Also performs the step bookkeeping (block 0): ctl->i = i, count = 0, hsub = restarted ? 0 : beta,
// H(i,i-1) = hsub (and H(i-1,i) when symmetric)      (Lanczos.h:127-128, Arnoldi.h:239).
// One column block of the fused step: c < nblocks-1 accumulates the raw product into w, the last block applies the step head.
// x_block points at the operand slice of block c (DeviceCsr::x_of_block).
int spmv_num_blocks(const DeviceCsr& A);
// dot_out != nullptr asks the last block to also deliver dot_out[0..i] = V[:, :i+1]^T w (the first panel pass of the step, fused into the
// operator kernel -- sliced layout only); the return value says whether it did (false: the caller runs the panel pass itself).
bool launch_spmv_step_block(const DeviceCsr& A, const SpmvPlan& plan, int c, const double* x_block, const double* f_loc, double* V, int64_t ldv, double* w,
                            FacCtl* ctl, double* H, int m, int i, int restarted, bool symmetric, const RedScratch& rs, cudaStream_t stream,
                            double* dot_out = nullptr);
// (every kernel of a step returns immediately when ctl->abort is set: sweeps enqueued without host round trips, see FacCtl::abort)
bool launch_spmv_step(const DeviceCsr& A, const SpmvPlan& plan, const double* x_full, const double* f_loc, double* V, int64_t ldv, double* w, FacCtl* ctl,
                      double* H, int m, int i, int restarted, bool symmetric, const RedScratch& rs, cudaStream_t stream, double* dot_out = nullptr);

// ---- peer-memory kernels (peer.cu; row-sharded runs over NVLink) --------------------------------------------------------
constexpr int kPeerMax = 16;
// Mailboxes of the one-shot all-reduce inside each rank's control window: 2 (round parity) x nranks x kRedStride doubles of data
// followed by 2 x nranks 64-bit flags.
struct PeerCtl
{
    double* slots[kPeerMax];                 // slots[r] = rank r's mailbox area (mapped peer pointer; slots[rank] is local)
    unsigned long long* flags[kPeerMax];     // flags[r] = rank r's flag area
    unsigned long long* seq;                 // local round counter
    int rank, nranks;
};
inline size_t peer_ctl_bytes(int nranks) { return sizeof(double) * 2 * (size_t) nranks * kRedStride + sizeof(unsigned long long) * 2 * (size_t) nranks; }
// in-place all-reduce (op 0 = sum in rank order, 1 = max) of count <= kRedStride doubles in buf: every rank writes its values into every
// peer's mailbox, releases a flag per peer, waits for the flags of all peers and combines the mailboxes in rank order (bitwise identical
// results on every rank).  One launch, no host involvement; doubles as a barrier for peer writes issued by earlier kernels of the stream.
// abort (optional): device flag; the kernel is a no-op when it is set (sweep mode; identical on every rank).
void launch_peer_allreduce(const PeerCtl& pc, double* buf, int count, int op, cudaStream_t stream, const int* abort = nullptr);
// Destination of this rank's residual rows in the chunk-major operand buffer of every rank (DeviceCsr::chunk_len): local row r goes to
// x[(r / len) * stride + rank * len + r % len] on all ranks.
struct PeerX
{
    double* dst[kPeerMax];   // operand buffers of all ranks (own included); np == 0: no peer writes
    int np;
    int rank;
    int64_t len, stride, rows;  // rows = chunk rows (len * number of chunks): rows at or beyond it have no destination
    int64_t row0;               // local row of the first entry the kernel sees (row-range launches of the correction pass)
};
// x (all ranks) <- f_loc rows of this rank (cold paths: after init / restart / expand_basis; the hot path writes from the panel pass)
void launch_peer_push(const PeerX& px, const double* f_loc, int64_t nrows_ld, cudaStream_t stream);

// ---- fused Krylov-panel passes (panel.cu) ---------------------------------------------------------
enum PanelMode
{
    PANEL_DOT = 0,   // red = V^T x, ||x||^2          (x unchanged)
    PANEL_FORM = 1,  // f = w - alpha*V[:,j-1]; red = V^T f, ||f||^2       (Lanczos.h:145-152)
    PANEL_CORR = 2,  // f = x - V c;            red = V^T f, ||f||^2       (Lanczos.h:171-179, Arnoldi.h:254)
};
// j = number of panel columns (<= kPanelMaxCols) is read from ctl->i + 1 when j_host < 0.
// x: input vector (w for FORM, f or w for CORR/DOT), f_out: output residual (may alias x).
// coef: device pointer to the coefficients (CORR) or to alpha (FORM).
// pred (optional): device flag; the kernel is a no-op when *pred == 0 (speculatively enqueued correction pass).
// cplx: the vectors hold interleaved complex entries and nrows / ldv count doubles (Hermitian path); red_out / coef then use the
// layout [0, j) = Re, [kRedNrm] = ||f||^2, [kRedNrm + 1, kRedNrm + 1 + j) = Im and j <= 63.
// push (optional, CORR mode, real path): also write the new residual rows into every rank's SpMV operand buffer (peer memory).
void launch_panel_pass(int mode, const double* V, int64_t ldv, int64_t nrows, int j, const double* x, double* f_out, const double* coef, double* red_out,
                       const RedScratch& rs, cudaStream_t stream, const int* pred = nullptr, bool cplx = false, const PeerX* push = nullptr,
                       const int* abort = nullptr, int64_t row_limit = -1);
// (row_limit: rows at or beyond it -- relative to the pointers passed -- are treated as absent; default ldv, i.e. the zero padding rows of a
//  full-length pass are read like real ones.  Row-range launches pass V / x / f_out already offset and the length of their range.)

// Decide kernels: consume ctl->red after a panel pass (and after the all-reduce when sharded).
//  stage 0: after c = V^T w (fills H(i,i), H(i-1,i), the coefficients of the first pass f = w - V c); stage 1: after that pass;
//  stage 2: after a further correction pass (also applies the H update of Lanczos.h:172-175 with the coefficients just used).
//  sweep != 0 (stage 1 only): the step was enqueued without a host round trip; raise ctl->abort when the host has to take over.
//  two_part != 0 (stage 1, real path): the pass ran in two row ranges; ctl->red2 is added to ctl->red first.
void launch_lanczos_decide(FacCtl* ctl, double* H, int m, double beta_thresh, int stage, cudaStream_t stream, int predicated = 0, bool cplx = false,
                           int sweep = 0, int two_part = 0);
// Peer-memory runs: all-reduce of ctl->red[0..count) over the ranks and the decisions of `stage` in one launch (real Lanczos path).
void launch_lanczos_decide_peer(const PeerCtl& pc, int count, FacCtl* ctl, double* H, int m, double beta_thresh, int stage, cudaStream_t stream, int sweep = 0,
                                int two_part = 0);
// Arnoldi flavours (Arnoldi.h:242-290): stage 0 = after h = V^T w (copies h into H(:,i) and c),
// stage 1 = after f = w - V h (DGKS test), stage 2 = after a correction pass.
// Hi != nullptr: complex Arnoldi (Hermitian-path layout of ctl->red / ctl->c); H receives the real and Hi the imaginary parts of H(:, i)
// sweep != 0 (real path): the step was enqueued without a host round trip -- completed steps add their counters to ctl->acc_*, a step that
// needs the host (a third pass, a zeroed residual, beta below near_0) raises ctl->abort.
void launch_arnoldi_decide(FacCtl* ctl, double* H, int m, double beta_thresh, int stage, cudaStream_t stream, int predicated = 0, double* Hi = nullptr,
                           int sweep = 0);

// ---- restart GEMM (panel.cu) ------------------------------------------------------------------------
// Vout[:, c] = sum_j V[:, j] * Q[j, c]  for c < kk  (Q: m x m column-major on device, ldq = m).
// In place when Vout == V.  When f != nullptr also  f = f*Q(m-1,kk-2) + Vout[:,kk-1]*H(kk-1,kk-2)
// and red_out[0] = ||f||^2   (Arnoldi.h:320-340).
void launch_compress_fma(const double* V, int64_t ldv, int64_t nrows, int m, const double* Q, int kk, double* Vout, int64_t ldo, double* f, const double* H,
                         double* red_out, const RedScratch& rs, cudaStream_t stream);
void dense_compress_host(int64_t n, int m, int kk, const double* V, const double* Q, const double* H, double* Vout, double* f, double* fnorm2, int impl);
void launch_compress_dmma(const double* V, int64_t ldv, int64_t nrows, int m, const double* Q, int kk, double* Vout, int64_t ldo, double* f, const double* H,
                          double* red_out, const RedScratch& rs, cudaStream_t stream);
void launch_compress(const double* V, int64_t ldv, int64_t nrows, int m, const double* Q, int kk, double* Vout, int64_t ldo, double* f, const double* H,
                     double* red_out, const RedScratch& rs, cudaStream_t stream);

// ---- BLAS-1 helpers for init / expand_basis (vecops.cu) ------------------------------------------------
enum VecReduceOp
{
    VR_SUMSQ = 0,
    VR_DOT = 1,
    VR_MAXABS = 2,
    // interleaved complex vectors of n doubles (Hermitian path): Im <x, y> = sum x_re y_im - x_im y_re, max |x_k| (modulus)
    VR_CDOT_IM = 3,
    VR_CMAXABS = 4
};
void launch_vec_reduce(int op, const double* x, const double* y, int64_t n, double* out, const RedScratch& rs, cudaStream_t stream);
void launch_vec_scale(const double* x, double s, int divide, double* y, int64_t n, cudaStream_t stream);          // y = x*s or x/s
void launch_vec_axpy(const double* w, const double* v, double a, double* f, int64_t n, cudaStream_t stream);      // f = w - a*v
void launch_vec_caxpy(const double* w, const double* v, double ar, double ai, double* f, int64_t n, cudaStream_t stream);  // complex f = w - (ar + i ai) v, n doubles
// complex restart (complex Q): V(:, c) = A(:, c) + i B(:, c) for c < kk on interleaved data; f <- f Q(m-1,kk-2) + V(:,kk-1) H(kk-1,kk-2)
void launch_zcombine(const double* A, const double* B, int64_t ldab, double* V, int64_t ldv, int64_t n, int kk, cudaStream_t stream);
void launch_zf_update(double* f, const double* vk, const double* Qr, const double* Qi, const double* Hr, const double* Hi, int m, int kk, int64_t n,
                      cudaStream_t stream);
void launch_set_beta(FacCtl* ctl, const double* red_slot, int take_sqrt, cudaStream_t stream);                      // ctl->beta = (sqrt) *red_slot
void launch_set_scalar(double* dst, double v, cudaStream_t stream);
// Host-operator path (user OpType): the two halves of the fused step head around the host call.
//   scale:    V[:,i] = f / beta                                              (Lanczos.h:106, Arnoldi.h:236)
//   epilogue: w -= hsub * V[:,i-1] (sym), red_a[0] = <V[:,i], w> (sym), step bookkeeping as in launch_spmv_step
void launch_step_scale(const double* f, const FacCtl* ctl, double* vi, int64_t n, cudaStream_t stream);
void launch_step_epilogue(double* w, const double* V, int64_t ldv, int64_t n, FacCtl* ctl, double* H, int m, int i, int restarted, bool symmetric,
                          const RedScratch& rs, cudaStream_t stream);

// ---- small dense restart kernels (dense_sym.cu) ----------------------------------------------------------
struct SymRestartOut
{
    int nconv;
    int k;      // nev_adjusted
    int info;   // 0 ok, 1 TridiagEigen failed
    int pad;
};
// retrieve_ritzpair + num_converged (+ nev_adjusted + shifted-QR chain when not converged), all on
// the device in one single-CTA kernel.  H (m x m) is replaced by Q'HQ, Q receives the accumulated
// rotations.  ritz_vec: m x nev.  do_restart = 0 skips the QR chain (used after the last step).
void launch_sym_restart(double* H, int m, int nev, const FacCtl* ctl, int selection, double tol, double* ritz_val, double* ritz_est, double* ritz_vec,
                        int* ritz_conv, double* Q, SymRestartOut* out, int do_restart, cudaStream_t stream);
// standalone pieces for the unit tier
void launch_tridiag_eigen(const double* H, int m, double* evals, double* evecs, int* info, cudaStream_t stream);
void launch_givens_batch(int variant, int64_t count, const double* x, const double* y, double* r, double* c, double* s, cudaStream_t stream);
void launch_tridiag_qr(const double* H, int m, double shift, double* QtHQ, double* Q, cudaStream_t stream);

}  // namespace sb200

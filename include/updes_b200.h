/*
 * updes_b200.h -- C-ABI of the B200-native Updes hot path (libupdes_b200.so).
 *
 * The reference (ddrous/Updes) has no FFI: its boundary is the Python call surface
 *   pde_solver / pde_solver_jit          updes/operators.py:559-683
 * whose arithmetic lives in
 *   assemble_Phi / assemble_P / assemble_A        updes/assembly.py:10-85
 *   assemble_op_Phi_P / assemble_bd_Phi_P         updes/assembly.py:93-362
 *   assemble_invert_A / assemble_B                updes/assembly.py:87-90, :366-401
 *   core_compute_coefficients                     updes/assembly.py:404-410
 *   lx.linear_solve(..., lx.QR())                 updes/operators.py:612-613
 *   value / gradient / laplacian / divergence     updes/operators.py:118-351
 * Each entry point below names the reference code it replaces.  INTEGRATION.md shows the
 * jax.ffi / ctypes stubs a maintainer would add on the reference side.
 *
 * Conventions
 *   - plain pointers and sizes only; every buffer is caller-allocated DEVICE memory unless a
 *     parameter is documented as host; the library never frees caller memory;
 *   - all work is enqueued on the caller's cudaStream_t (passed as void*), no hidden threads;
 *   - return value: 0 ok; < 0 bad argument (-k = k-th argument); > 0 CUDA runtime error code
 *     (cudaError_t) raised while enqueuing.  Numerical status (LAPACK-style "zero pivot at
 *     column k", 1-based) is written to a device int32 `info` so no call forces a sync;
 *   - matrices are ROW-MAJOR with leading dimension `ld` (elements); `ld` must be a multiple
 *     of 16 (128-byte rows: TMA tiles, 16-byte vector stores) and the base 128-byte aligned;
 *   - all arithmetic is FP64 (reference default, updes/config.py:15-16).
 */
#ifndef UPDES_B200_H
#define UPDES_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* rbf kinds, updes/utils.py:30-69 (param = a for the first two, eps for the others) */
enum {
  UPDES_RBF_POLYHARMONIC = 0,         /* r^(2a+1)            utils.py:50-55 */
  UPDES_RBF_THIN_PLATE = 1,           /* r^(2a) log r        utils.py:63-69 */
  UPDES_RBF_GAUSSIAN = 2,             /* exp(-(eps r)^2)     utils.py:44-48 */
  UPDES_RBF_MULTIQUADRIC = 3,         /* sqrt(1+(eps r)^2)   utils.py:30-35 */
  UPDES_RBF_INVERSE_MULTIQUADRIC = 4  /* 1/sqrt(1+(eps r)^2) utils.py:37-42 */
};

/*
 * Row descriptors of the collocation system (one per collocation row r < N).
 * A row is  sum over up to two evaluation points p of  c_p . jet(x_p, centre_j)  with
 * jet = (phi, phi_x, phi_y, phi_xx, phi_yy) taken w.r.t. the evaluation point, derivatives
 * forced to 0 at r == 0 (the reference's nan_to_num, operators.py:58,:83,:109).  This one form
 * covers internal operator rows (assembly.py:93-137) and Dirichlet / Neumann / Robin /
 * periodic rows (assembly.py:141-362); the host layer fills it (updes_b200/assembly.py).
 */
typedef struct UpdesRows {
  const double *pts;        /* [npts x 2] evaluation points (for K: the cloud's sorted_nodes) */
  const int32_t *p1;        /* [R] index into pts of the first evaluation point */
  const int32_t *p2;        /* [R] second point or -1 (periodic rows, assembly.py:215-267) */
  const double *cphi1;      /* [R x 5] coefficients on RBF columns at p1 */
  const double *cphi2;      /* [R x 5] ... at p2 (read only where p2 >= 0) */
  const double *cpol1;      /* [R x 5] coefficients on monomial columns at p1 (differs from
                               cphi1 only for the Robin normal quirk, assembly.py:206 vs :303) */
  const double *cpol2;      /* [R x 5] ... at p2 */
  const int32_t *skip;      /* [R] centre index whose column is left 0 (the reference drops a
                               node from its own support, cloud.py:110-112), or -1 */
} UpdesRows;

/* ---- assembly (replaces assembly.py:93-362 + the P^T block of :62-85) --------------------
 * Fills rows [row0, row0+nrows) x columns [0, ld) of the (N+M) x (N+M) collocation matrix
 *   K = [[op(Phi) op(P)], [bd(Phi) bd(P)], [P^T 0]]            
 * into `out` (row-major, leading dimension ld, out[0] = K[row0][0]).  Padding columns
 * [N+M, ld) are zeroed.  centres = sorted_nodes [N x 2].  rows describes collocation rows
 * 0..N-1 (rows >= N are the P^T rows).  Whole matrix: row0 = 0, nrows = N+M.
 * jet_mask: which parts of the jet the coefficient rows in the range actually use
 * (UPDES_JET_VAL | UPDES_JET_GRAD | UPDES_JET_HESS); the kernel is specialised on it so that a
 * Laplace operator does not pay for phi and grad phi.  0 or 7 = everything.
 */
#define UPDES_JET_VAL 1
#define UPDES_JET_GRAD 2
#define UPDES_JET_HESS 4
/* every row in the range is c0*phi + c3*(phi_xx + phi_yy) (no gradient term, c3 == c4): the kernel
 * then evaluates the radial Laplacian in closed form (combine with UPDES_JET_VAL if c0 != 0) */
#define UPDES_JET_ISO_LAPLACIAN 8
int updes_assemble_rows(int rbf_kind, double rbf_param, int N, int M, const double *centres,
                        const UpdesRows *rows, int64_t row0, int64_t nrows, int jet_mask,
                        double *out, int64_t ld, void *stream);

/* Same entries for an arbitrary rectangular block [row0,row0+nrows) x [col0,col0+ncols)
 * written to out[(r-row0)*ld + (c-col0)] -- the tile form used by block-cyclic sharding. */
int updes_assemble_block(int rbf_kind, double rbf_param, int N, int M, const double *centres,
                         const UpdesRows *rows, int64_t row0, int64_t nrows, int64_t col0,
                         int64_t ncols, int jet_mask, double *out, int64_t ld, void *stream);

/* ---- matrix-free jets (replaces operators.py:118-351 value/gradient/laplacian and gives
 * K.c, [Phi P].c without storing a matrix) ---------------------------------------------------
 * For every evaluation point i < npts and field f < nf:
 *   jphi[(f*npts + i)*5 + k] = sum_{j<N, j != skip[i]} coeffs[f*ldc + j] * jet_k(pts_i, centre_j)
 *   jpol[(f*npts + i)*5 + k] = sum_{m<M} coeffs[f*ldc + N + m] * jet_k(monomial_m)(pts_i)
 * skip may be NULL (no column skipped).  workspace: updes_eval_jets_workspace_bytes().
 */
size_t updes_eval_jets_workspace_bytes(int N, int npts, int nf);
int updes_eval_jets(int rbf_kind, double rbf_param, int N, int M, const double *centres,
                    const double *coeffs, int64_t ldc, int nf, const double *pts, int npts,
                    const int32_t *skip, double *jphi, double *jpol, void *workspace, void *stream);

/* ---- dense LU (replaces jnp.linalg.inv + GEMM + lineax QR, assembly.py:87-90,:398-401,
 * operators.py:612-616, by one factorisation P K = L U) ------------------------
 * Opaque handle: owns tensor maps and a small device workspace; release with updes_lu_destroy.
 */
typedef struct UpdesLU UpdesLU;

int updes_lu_create(UpdesLU **handle, int64_t n, int64_t ld);
int updes_lu_destroy(UpdesLU *handle);
/* Factor K in place (row-major, partial pivoting by rows).  ipiv[n] (device, int32, 0-based:
 * row k was exchanged with row ipiv[k]); info (device int32): 0 or 1-based first zero pivot. */
int updes_lu_factor(UpdesLU *handle, double *K, int32_t *ipiv, int32_t *info, void *stream);
/* Same with row equilibration first (`scale` vector): every row is multiplied by the power of two
 * that brings its largest magnitude into [1, 2) -- exact, so the factors are those of diag(scale) K -- and the
 * factors are written to scale[n] (device, caller-owned; it must stay alive as long as the handle solves:
 * updes_lu_solve multiplies right-hand sides by it).  Collocation rows differ in scale by orders of magnitude
 * (operator rows ~ 1/DT or 9 r, boundary rows ~ r^3, polynomial rows ~ 1), which is what partial pivoting
 * compares across. */
int updes_lu_factor_scaled(UpdesLU *handle, double *K, int32_t *ipiv, double *scale, int32_t *info, void *stream);
/* Building blocks of the same for column-sharded matrices (row maxima are combined across ranks by the caller):
 * out[r] = max_c |A[r][c]| over c < cols;  scale[i] = 2^-floor(log2 absmax[i]) (1 for zero rows);
 * A[r][:] *= scale[r];  updes_lu_set_row_scale tells the handle which factors its solves apply (NULL = none). */
int updes_row_absmax(const double *A, int64_t rows, int64_t cols, int64_t ld, double *out, void *stream);
int updes_scale_from_absmax(const double *absmax, int64_t n, double *scale, void *stream);
int updes_row_scale(double *A, int64_t rows, int64_t cols, int64_t ld, const double *scale, void *stream);
int updes_lu_set_row_scale(UpdesLU *handle, const double *scale);
/* Internal-failure flags of the handle's kernels (synchronises `stream`): bit 0 = a wait inside a triangular
 * sweep timed out.  A timed-out grid barrier of the panel kernel is reported through info = -1. */
int updes_lu_status(UpdesLU *handle, int32_t *host_flags, void *stream);
/* Solve K X = B for nrhs right-hand sides using the factors.  B is [nrhs][ldb] (each
 * right-hand side contiguous, ldb >= n), overwritten by X.  transpose != 0 solves K^T X = B (adjoint
 * solves; same factors: U^T w = b, L^T z = w, x = P^T z).  Use the handle that factored. */
int updes_lu_solve(UpdesLU *handle, const double *LU, const int32_t *ipiv, double *B,
                   int64_t ldb, int nrhs, int transpose, void *stream);

/* Building blocks, exported so tests can pin each kernel against the oracle. */
int updes_dgemm_sub(UpdesLU *handle, double *K, int64_t rc, int64_t cc, int64_t ra, int64_t ca,
                    int64_t rb, int64_t cb, int64_t m, int64_t n, int64_t k, void *stream);
int updes_lu_panel(UpdesLU *handle, double *K, int64_t r0, int64_t nc, int32_t *ipiv,
                   int32_t *info, void *stream);

/* ---- multi-GPU building blocks (column-block-cyclic LU, one process per GPU) ----------------------
 * The local matrix (all n rows, this rank's column blocks) is slot 0 of the handle
 * (updes_lu_create(n, ld_local) + updes_lu_bind(h, 0, ...)); panels received from other ranks live in
 * auxiliary buffers bound to slots 1..3.  Pivot indices are global row numbers.  The host driver
 * (updes_b200/distributed.py) strings these together with NCCL broadcasts of the panels.
 */
int updes_lu_bind(UpdesLU *handle, int slot, double *ptr, int64_t rows, int64_t ld);
/* cap the persistent GEMM grid (0 = one CTA per SM) so NCCL kernels can run beside the update */
int updes_lu_set_gemm_ctas(UpdesLU *handle, int ctas);
/* trailing-update GEMM schedule, bit 0: 1 = ping-pong (two 128x64 CTAs per SM), 0 = one 128x128 CTA per SM;
 * bit 1: 32-deep pipeline stages (two 16-k sub-tiles per barrier round); bit 3: 1 = epilogue as a read-modify-write
 * of C instead of the default fire-and-forget red.global.add.f64 (every C element belongs to exactly one thread of one
 * launch, so both give the same bits); bit 2 (with bit 3): 1 = do not prefetch the C tile into L2 before that epilogue.
 * Bits 2-3 are comparison hooks. */
int updes_lu_set_gemm_variant(UpdesLU *handle, int variant);
/* base panels: 2 (default) = implicit-pivoting kernels -- panels of <= 8 192 rows run in one thread-block cluster
 * whose CTAs push their candidates into each other's shared memory (one split cluster barrier per column), taller
 * ones in the grid-wide cooperative kernel; 1 = first-generation cluster + grid kernels; 0 = first-generation grid
 * kernel only */
int updes_lu_set_panel_variant(UpdesLU *handle, int variant);
/* tuning hook: largest block (rows: 32, 64, 96 or 128; default 128) of the unit-lower triangular solve
 * U12 = L11^-1 A12 handled by one substitution kernel instead of recursion + small GEMMs */
int updes_lu_set_trsm_base(UpdesLU *handle, int rows);
/* triangular solves: 2 (default) = row-block streaming sweeps (one launch per direction, solved blocks published
 * through the data), 1 = step-synchronous persistent sweeps, 0 = one launch per 128-row block */
int updes_lu_set_solve_variant(UpdesLU *handle, int variant);
/* test hook: rows the 32-wide register-resident panel holds (0 = default 148*640); smaller values
 * force the 16- / 8-wide base panels used for panels taller than 94 720 / 189 440 rows */
int updes_lu_set_panel_capacity(UpdesLU *handle, int64_t rows);
/* LU of the tall panel rows [r0, rows) x columns [c0, c0+nc) of `slot`; interchanges are applied to
 * the panel columns only; ipiv[r0 .. r0+nc) receives the pivots. */
int updes_lu_panel_factor(UpdesLU *handle, int slot, int64_t r0, int64_t c0, int64_t nc, int32_t *ipiv,
                          int32_t *info, void *stream);
/* apply interchanges k0+t <-> ipiv[k0+t], t < npiv, to columns [c_lo, c_hi) of `slot` */
int updes_lu_apply_swaps(UpdesLU *handle, int slot, int64_t c_lo, int64_t c_hi, int64_t k0, int64_t npiv,
                         const int32_t *ipiv, void *stream);
/* B <- L^-1 B: L unit-lower n1 x n1 at (rl, cl) of slot_l, B n1 x ncols at (rb, cb) of slot_b */
int updes_lu_trsm(UpdesLU *handle, int slot_l, int64_t rl, int64_t cl, int64_t n1, int slot_b, int64_t rb,
                  int64_t cb, int64_t ncols, void *stream);
/* C -= A B with A (m x k) in slot_a, B (k x n) in slot_b, C (m x n) in slot_c */
int updes_lu_gemm(UpdesLU *handle, int slot_a, int64_t ra, int64_t ca, int slot_b, int64_t rb, int64_t cb,
                  int slot_c, int64_t rc, int64_t cc, int64_t m, int64_t n, int64_t k, void *stream);
/* solves: row permutation from a complete pivot list; gather X = B[perm]; substitution restricted
 * to one column block whose diagonal sits at rows [r0, r0+width), columns [c0, c0+width) of `slot` */
int updes_lu_set_pivots(UpdesLU *handle, const int32_t *ipiv, void *stream);
int updes_lu_permute_rhs(UpdesLU *handle, const double *B, int64_t ldb, int nrhs, double *X, void *stream);
int updes_tri_block_sweep(UpdesLU *handle, int slot, int upper, int64_t r0, int64_t c0, int64_t width,
                          double *X, int nrhs, void *stream);
/* left-looking distributed substitution (updes_b200/distributed.py): every rank forms the partial products of one
 * row block against the solution entries it owns, out[i] = sum_{c_lo <= c < c_hi} A[r0+i][c] x[c] (x indexed by LOCAL
 * column; an empty range writes zeros), the partials are summed on the block's owner, which then solves ONLY the
 * width x width diagonal block in place on X[r0 .. r0+width) (upper = 0: unit lower, 1: upper with diagonal) */
int updes_block_gemv(UpdesLU *handle, int slot, int64_t r0, int64_t nrows, int64_t c_lo, int64_t c_hi,
                     const double *x, double *out, void *stream);
int updes_tri_diag_solve(UpdesLU *handle, int slot, int upper, int64_t r0, int64_t c0, int64_t width, double *X,
                         void *stream);

/* ---- misc ------------------------------------------------------------------------------------ */
const char *updes_b200_version(void);
/* number of kernels launched by this library since load (bench.py's gpu_launches) */
int64_t updes_launch_count(void);
/* Optional CUDA-event timing per kernel class (bench.py's roofline legs).  enable(1) clears the
 * records and starts recording on the launching stream; read() synchronises and returns the summed
 * milliseconds, work (flops for 0-3, bytes for 4-5) and launch count of one class:
 * 0 trailing-update GEMM, 1 panel, 2 row interchanges, 3 triangular base solve, 4 assembly, 5 solve. */
int updes_profile_enable(int on);
int updes_profile_read(int cat, double *ms, double *work, int64_t *count);
/* per-launch records of one class in launch order (ms[i], work[i], i < min(return value, max)); returns the number
 * of records of the class, or a negative CUDA error code */
int64_t updes_profile_records(int cat, double *ms, double *work, int64_t max);
/* tuning hook: columns per thread of the assembly kernel (bit 0: closed-form Laplacian rows 4 instead of 2;
 * bit 1: general jets 2 instead of 4) */
int updes_assemble_set_variant(int variant);

#ifdef __cplusplus
}
#endif
#endif /* UPDES_B200_H */

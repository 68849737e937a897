// Triangular solves with the LU factors (factor once / solve many; replaces the reference's
// repeated inv(A) matvecs and QR solves, updes/assembly.py:404-410, updes/operators.py:612-616).
//
// HBM-bound: L and U are each streamed exactly once per solve (8 n^2 bytes in total).
// Column-sweep blocked substitution on the row-major factors: for each diagonal block of SB rows,
// one CTA solves the SB x SB triangle out of shared memory, then a grid-wide kernel subtracts the
// block's contribution from all remaining rows (one warp per row: a 1 KB contiguous row segment
// dotted with the block solution held in shared memory).
#include "lu.cuh"

namespace updes {

constexpr int SB = 128;
constexpr int SOLVE_MAX_RHS = 4;

__global__ void gather_rows_kernel(const double *B, long long ldb, int nrhs, const int32_t *perm, long long n,
                                   double *X) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int p = perm[i];
  for (int f = 0; f < nrhs; f++) X[f * n + i] = B[f * ldb + p];
}

__global__ void copy_back_kernel(double *B, long long ldb, int nrhs, long long n, const double *X) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int f = 0; f < nrhs; f++) B[f * ldb + i] = X[f * n + i];
}

// One block step of the substitution, fused in one launch.  Every CTA
//   A. stages the nb x nb diagonal block T = LU[k0.., c0..] in shared memory (the block is read from
//      L2 by all CTAs; 128 KB each) and the current right-hand-side block,
//   B. solves it redundantly: four 32 x 32 triangular solves by one warp (the 32 unknowns live in the
//      lanes, one shuffle broadcast per column) interleaved with block updates by all threads --
//      plain substitution, so the solve keeps LAPACK's backward stability (no explicit inverses),
//   C. subtracts the block's contribution from its share of the remaining rows, one warp per row
//      (a 1 KB contiguous row segment dotted with the block solution held in shared memory).
// CTA 0 stores the solved block into `Yout`; the remaining rows are updated in place in `X`.
// UPPER = false: unit lower (forward sweep); UPPER = true: upper with diagonal (backward sweep).
constexpr int STEP_THREADS = 1024;
constexpr size_t STEP_SMEM = sizeof(double) * (SB * (SB + 1) + SOLVE_MAX_RHS * SB);

template <bool UPPER>
__global__ void __launch_bounds__(STEP_THREADS, 1)
tri_step_kernel(const double *LU, long long ld, long long n, long long k0, long long c0, int nb, long long i0,
                long long i1, double *X, double *Yout, int nrhs) {
  extern __shared__ double sm[];
  double *T = sm;                       // [SB][SB+1]
  double *xs = sm + SB * (SB + 1);      // [nrhs][SB]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NWARPS = STEP_THREADS / 32;
  // ---- A. stage T and the rhs block ------------------------------------------------------------
  if (nb == SB) {
    // 32 warps x 4 rows: issue all 16 loads of a warp before the first store (latency-bound otherwise)
    double v[4][4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const double *src = LU + (k0 + warp + u * NWARPS) * ld + c0;
#pragma unroll
      for (int q = 0; q < 4; q++) v[u][q] = src[lane + 32 * q];
    }
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
      for (int q = 0; q < 4; q++) T[(warp + u * NWARPS) * (SB + 1) + lane + 32 * q] = v[u][q];
  } else {
    for (int r = warp; r < nb; r += NWARPS) {
      const double *src = LU + (k0 + r) * ld + c0;
      for (int c = lane; c < nb; c += 32) T[r * (SB + 1) + c] = src[c];
    }
  }
  for (int t = tid; t < nrhs * SB; t += STEP_THREADS) {
    const int f = t / SB, c = t % SB;
    xs[f * SB + c] = c < nb ? X[f * n + k0 + c] : 0.0;
  }
  __syncthreads();
  // ---- B. triangular solve of the block ------------------------------------------------------------
  const int nsub = (nb + 31) / 32;
  for (int q = 0; q < nsub; q++) {
    const int sb = UPPER ? (nsub - 1 - q) : q;
    const int base = sb * 32;
    const int cnt = min(32, nb - base);
    if (warp == 0) {
      for (int f = 0; f < nrhs; f++) {
        double x = lane < cnt ? xs[f * SB + base + lane] : 0.0;
        if (!UPPER) {
          for (int c = 0; c < cnt; c++) {
            const double xc = __shfl_sync(0xffffffffu, x, c);
            if (lane > c && lane < cnt) x = fma(-T[(base + lane) * (SB + 1) + base + c], xc, x);
          }
        } else {
          for (int c = cnt - 1; c >= 0; c--) {
            if (lane == c) x = x / T[(base + c) * (SB + 1) + base + c];
            const double xc = __shfl_sync(0xffffffffu, x, c);
            if (lane < c) x = fma(-T[(base + lane) * (SB + 1) + base + c], xc, x);
          }
        }
        if (lane < cnt) xs[f * SB + base + lane] = x;
      }
    }
    __syncthreads();
    // rows of the block still to be solved lose the contribution of sub-block sb
    const int rbeg = UPPER ? 0 : base + 32, rend = UPPER ? base : nb;
    for (int t = tid; t < (rend - rbeg) * nrhs; t += STEP_THREADS) {
      const int f = t / (rend - rbeg), r = rbeg + t % (rend - rbeg);
      double v = xs[f * SB + r];
      const double *trow = T + r * (SB + 1) + base;
      for (int c = 0; c < cnt; c++) v = fma(-trow[c], xs[f * SB + base + c], v);
      xs[f * SB + r] = v;
    }
    __syncthreads();
  }
  if (blockIdx.x == 0)
    for (int t = tid; t < nrhs * nb; t += STEP_THREADS) {
      const int f = t / nb, c = t % nb;
      Yout[f * n + k0 + c] = xs[f * SB + c];
    }
  // ---- C. update the remaining rows -------------------------------------------------------------------
  // four rows in flight per warp (4 KB of loads outstanding) and fire-and-forget reductions into X:
  // the phase is latency-bound otherwise (one 1 KB row per warp and a dependent read-modify-write).
  constexpr int RU = 4;
  const long long wstride = (long long)gridDim.x * NWARPS;
  for (long long i = i0 + (long long)blockIdx.x * NWARPS + warp; i < i1; i += RU * wstride) {
    if (nb == SB) {
      double2 v0[RU], v1[RU];
#pragma unroll
      for (int u = 0; u < RU; u++) {
        const long long r = i + u * wstride;
        if (r < i1) {
          const double *row = LU + r * ld + c0;
          v0[u] = *reinterpret_cast<const double2 *>(row + 2 * lane);
          v1[u] = *reinterpret_cast<const double2 *>(row + 64 + 2 * lane);
        } else {
          v0[u] = make_double2(0.0, 0.0); v1[u] = v0[u];
        }
      }
      for (int f = 0; f < nrhs; f++) {
        const double *xf = xs + f * SB;
        const double x0 = xf[2 * lane], x1 = xf[2 * lane + 1], x2 = xf[64 + 2 * lane], x3 = xf[64 + 2 * lane + 1];
        double a[RU];
#pragma unroll
        for (int u = 0; u < RU; u++) a[u] = fma(v1[u].y, x3, fma(v1[u].x, x2, fma(v0[u].y, x1, v0[u].x * x0)));
#pragma unroll
        for (int off = 16; off > 0; off >>= 1)
#pragma unroll
          for (int u = 0; u < RU; u++) a[u] += __shfl_xor_sync(0xffffffffu, a[u], off);
        if (lane < RU) {
          const long long r = i + lane * wstride;
          double mine = a[0];
#pragma unroll
          for (int u = 1; u < RU; u++) mine = lane == u ? a[u] : mine;
          if (r < i1) atomicAdd(X + f * n + r, -mine);
        }
      }
    } else {
#pragma unroll 1
      for (int u = 0; u < RU; u++) {
        const long long r = i + u * wstride;
        if (r >= i1) break;
        const double *row = LU + r * ld + c0;
        for (int f = 0; f < nrhs; f++) {
          double a = 0.0;
          for (int c = lane; c < nb; c += 32) a = fma(row[c], xs[f * SB + c], a);
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) a += __shfl_xor_sync(0xffffffffu, a, off);
          if (lane == 0) atomicAdd(X + f * n + r, -a);
        }
      }
    }
  }
}

static int ensure_solve_attrs() {
  static bool attr = false;
  if (!attr) {
    UPDES_CUDA_TRY(cudaFuncSetAttribute(tri_step_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)STEP_SMEM));
    UPDES_CUDA_TRY(cudaFuncSetAttribute(tri_step_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)STEP_SMEM));
    attr = true;
  }
  return 0;
}

static int step_grid(int num_sms, long long rows) {
  const long long want = (rows + 127) / 128;        // 32 warps x 4 rows per CTA pass
  return (int)(want < 1 ? 1 : (want > num_sms ? num_sms : want));
}

// Substitution restricted to one column block [c0, c0+w) whose diagonal sits at rows [r0, r0+w).
// Forward (unit lower): rows below r0 of X are the running right-hand side; the solved block goes to Y.
// Backward (upper): same upwards.  X and Y are full-length [nrhs][n] vectors (X != Y).
int tri_block_sweep(int num_sms, const double *LU, long long ld, long long n, bool upper, long long r0, long long c0,
                    long long w, double *X, double *Y, int nrhs, cudaStream_t st) {
  int rc = ensure_solve_attrs();
  if (rc) return rc;
  const long long nsub = (w + SB - 1) / SB;
  for (long long t = 0; t < nsub; t++) {
    const long long s = upper ? (nsub - 1 - t) : t;
    const long long k0 = r0 + s * SB, cc = c0 + s * SB;
    const int nb = (int)((w - s * SB) < SB ? (w - s * SB) : SB);
    if (!upper) {
      const long long i0 = k0 + nb;
      tri_step_kernel<false><<<step_grid(num_sms, n - i0), STEP_THREADS, STEP_SMEM, st>>>(LU, ld, n, k0, cc, nb, i0, n, X,
                                                                                         Y, nrhs);
    } else {
      tri_step_kernel<true><<<step_grid(num_sms, k0), STEP_THREADS, STEP_SMEM, st>>>(LU, ld, n, k0, cc, nb, 0, k0, X, Y,
                                                                                     nrhs);
    }
    UPDES_LAUNCH_CHECK();
  }
  return 0;
}

// whole solve on one GPU: X = P b in xbuf[0]; forward sweep X -> Y; backward sweep Y -> X
static int solve_chunk(UpdesLU *h, const double *LU, double *X, double *Y, int nrhs, cudaStream_t st) {
  const long long n = h->n;
  int rc = tri_block_sweep(h->num_sms, LU, h->ld, n, false, 0, 0, n, X, Y, nrhs, st);
  if (rc) return rc;
  return tri_block_sweep(h->num_sms, LU, h->ld, n, true, 0, 0, n, Y, X, nrhs, st);
}

}  // namespace updes

extern "C" int updes_lu_solve(UpdesLU *h, const double *LU, const int32_t *ipiv, double *B, int64_t ldb, int nrhs,
                              int transpose, void *stream) {
  using namespace updes;
  if (!h) return -1;
  if (!LU) return -2;
  if (!ipiv) return -3;
  if (!B) return -4;
  if (ldb < h->n) return -5;
  if (nrhs <= 0) return 0;
  if (transpose) return -7;   // K^T solves: not built yet
  cudaStream_t st = (cudaStream_t)stream;
  const long long n = h->n;
  for (int f0 = 0; f0 < nrhs; f0 += SOLVE_MAX_RHS) {
    const int nf = (nrhs - f0) < SOLVE_MAX_RHS ? (nrhs - f0) : SOLVE_MAX_RHS;
    double *Bf = B + (long long)f0 * ldb;
    gather_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(Bf, ldb, nf, h->perm, n, h->xbuf);
    UPDES_LAUNCH_CHECK();
    prof_begin(PROF_SOLVE, 8.0 * (double)n * (double)n, st);
    int rc = solve_chunk(h, LU, h->xbuf, h->xbuf + (size_t)SOLVE_MAX_RHS * n, nf, st);
    prof_end(st);
    if (rc) return rc;
    copy_back_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(Bf, ldb, nf, n, h->xbuf);
    UPDES_LAUNCH_CHECK();
  }
  return 0;
}

// ---- building blocks of the distributed solve -----------------------------------------------------
extern "C" int updes_lu_set_pivots(UpdesLU *h, const int32_t *ipiv, void *stream) {
  if (!h) return -1;
  if (!ipiv) return -2;
  return updes::build_permutation(h, ipiv, (cudaStream_t)stream);
}

extern "C" int updes_lu_permute_rhs(UpdesLU *h, const double *B, int64_t ldb, int nrhs, double *X, void *stream) {
  using namespace updes;
  if (!h) return -1;
  if (!B) return -2;
  if (ldb < h->n) return -3;
  if (!X) return -5;
  if (nrhs <= 0) return 0;
  gather_rows_kernel<<<(unsigned)((h->n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(B, ldb, nrhs, h->perm, h->n, X);
  UPDES_LAUNCH_CHECK();
  return 0;
}

__global__ void copy_block_kernel(double *X, const double *Y, long long n, long long r0, long long w, int nrhs) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= w * nrhs) return;
  const long long f = t / w, c = t % w;
  X[f * n + r0 + c] = Y[f * n + r0 + c];
}

extern "C" int updes_tri_block_sweep(UpdesLU *h, int slot, int upper, int64_t r0, int64_t c0, int64_t width,
                                     double *X, int nrhs, void *stream) {
  using namespace updes;
  if (!h) return -1;
  if (slot < 0 || slot >= UPDES_MAX_VIEWS || !h->view[slot].ptr) return -2;
  if (r0 < 0 || r0 + width > h->view[slot].rows) return -4;
  if (c0 < 0 || (c0 & 1) || c0 + width > h->view[slot].ld) return -5;
  if (!X) return -7;
  if (nrhs <= 0 || nrhs > SOLVE_MAX_RHS) return -8;
  cudaStream_t st = (cudaStream_t)stream;
  const long long n = h->view[slot].rows;
  double *Y = h->xbuf + (size_t)SOLVE_MAX_RHS * h->n;     // scratch for the solved block
  int rc = tri_block_sweep(h->num_sms, h->view[slot].ptr, h->view[slot].ld, n, upper != 0, r0, c0, width, X, Y, nrhs, st);
  if (rc) return rc;
  copy_block_kernel<<<(unsigned)((width * nrhs + 255) / 256), 256, 0, st>>>(X, Y, n, r0, width, nrhs);
  UPDES_LAUNCH_CHECK();
  return 0;
}

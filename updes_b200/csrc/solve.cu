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

// Solve the diagonal block [k0, k0+nb) in place.  UPPER = false: unit lower (forward);
// UPPER = true: non-unit upper (backward).
template <bool UPPER>
__global__ void __launch_bounds__(SB) diag_solve_kernel(const double *LU, long long ld, long long n, long long k0,
                                                        long long c0, int nb, double *X, int nrhs) {
  extern __shared__ double sm[];
  double *T = sm;                       // [SB][SB+1]
  double *xs = sm + SB * (SB + 1);      // [nrhs][SB]
  const int tid = threadIdx.x;
  for (int i = 0; i < nb; i++)
    if (tid < nb) T[i * (SB + 1) + tid] = LU[(k0 + i) * ld + c0 + tid];
  for (int f = 0; f < nrhs; f++)
    if (tid < nb) xs[f * SB + tid] = X[f * n + k0 + tid];
  __syncthreads();
  if (!UPPER) {
    for (int c = 0; c < nb; c++) {
      if (tid > c && tid < nb) {
        const double l = T[tid * (SB + 1) + c];
        for (int f = 0; f < nrhs; f++) xs[f * SB + tid] = fma(-l, xs[f * SB + c], xs[f * SB + tid]);
      }
      __syncthreads();
    }
  } else {
    for (int c = nb - 1; c >= 0; c--) {
      if (tid == c)
        for (int f = 0; f < nrhs; f++) xs[f * SB + c] = xs[f * SB + c] / T[c * (SB + 1) + c];
      __syncthreads();
      if (tid < c) {
        const double u = T[tid * (SB + 1) + c];
        for (int f = 0; f < nrhs; f++) xs[f * SB + tid] = fma(-u, xs[f * SB + c], xs[f * SB + tid]);
      }
      __syncthreads();
    }
  }
  for (int f = 0; f < nrhs; f++)
    if (tid < nb) X[f * n + k0 + tid] = xs[f * SB + tid];
}

// X[i] -= sum_c LU[i][k0 + c] * X[k0 + c] for rows i in [i0, i1); one warp per row.
__global__ void __launch_bounds__(256) block_update_kernel(const double *LU, long long ld, long long n, long long k0,
                                                          long long c0, int nb, long long i0, long long i1, double *X,
                                                          int nrhs) {
  __shared__ double xs[SOLVE_MAX_RHS][SB];
  for (int t = threadIdx.x; t < nrhs * SB; t += blockDim.x) {
    const int f = t / SB, c = t % SB;
    xs[f][c] = c < nb ? X[f * n + k0 + c] : 0.0;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long i = i0 + (long long)blockIdx.x * 8 + warp;
  if (i >= i1) return;
  const double *row = LU + i * ld + c0;
  double acc[SOLVE_MAX_RHS] = {0, 0, 0, 0};
  if (nb == SB) {
    // 128 contiguous doubles: each lane takes two 16-byte pieces (c0 and ld are even)
    const double2 v0 = *reinterpret_cast<const double2 *>(row + 2 * lane);
    const double2 v1 = *reinterpret_cast<const double2 *>(row + 64 + 2 * lane);
    for (int f = 0; f < nrhs; f++) {
      acc[f] = v0.x * xs[f][2 * lane];
      acc[f] = fma(v0.y, xs[f][2 * lane + 1], acc[f]);
      acc[f] = fma(v1.x, xs[f][64 + 2 * lane], acc[f]);
      acc[f] = fma(v1.y, xs[f][64 + 2 * lane + 1], acc[f]);
    }
  } else {
    for (int c = lane; c < nb; c += 32) {
      const double v = row[c];
      for (int f = 0; f < nrhs; f++) acc[f] = fma(v, xs[f][c], acc[f]);
    }
  }
  for (int f = 0; f < nrhs; f++) {
    double a = acc[f];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) a += __shfl_xor_sync(0xffffffffu, a, off);
    if (lane == 0) X[f * n + i] -= a;
  }
}

constexpr size_t DIAG_SMEM = sizeof(double) * (SB * (SB + 1) + SOLVE_MAX_RHS * SB);

static int ensure_solve_attrs() {
  static bool attr = false;
  if (!attr) {
    UPDES_CUDA_TRY(cudaFuncSetAttribute(diag_solve_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DIAG_SMEM));
    UPDES_CUDA_TRY(cudaFuncSetAttribute(diag_solve_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DIAG_SMEM));
    attr = true;
  }
  return 0;
}

// One column block [c0, c0+w) of a (possibly column-distributed) factor whose diagonal sits at
// rows [r0, r0+w): forward (unit lower) or backward (upper) substitution restricted to this block,
// updating X for all rows below / above.  X holds the full-length right-hand sides, [nrhs][n].
int tri_block_sweep(const double *LU, long long ld, long long n, bool upper, long long r0, long long c0, long long w,
                    double *X, int nrhs, cudaStream_t st) {
  int rc = ensure_solve_attrs();
  if (rc) return rc;
  const long long nsub = (w + SB - 1) / SB;
  for (long long t = 0; t < nsub; t++) {
    const long long s = upper ? (nsub - 1 - t) : t;
    const long long k0 = r0 + s * SB, cc = c0 + s * SB;
    const int nb = (int)((w - s * SB) < SB ? (w - s * SB) : SB);
    if (!upper) {
      diag_solve_kernel<false><<<1, SB, DIAG_SMEM, st>>>(LU, ld, n, k0, cc, nb, X, nrhs);
      UPDES_LAUNCH_CHECK();
      const long long i0 = k0 + nb;
      if (i0 < n) {
        block_update_kernel<<<(unsigned)((n - i0 + 7) / 8), 256, 0, st>>>(LU, ld, n, k0, cc, nb, i0, n, X, nrhs);
        UPDES_LAUNCH_CHECK();
      }
    } else {
      diag_solve_kernel<true><<<1, SB, DIAG_SMEM, st>>>(LU, ld, n, k0, cc, nb, X, nrhs);
      UPDES_LAUNCH_CHECK();
      if (k0 > 0) {
        block_update_kernel<<<(unsigned)((k0 + 7) / 8), 256, 0, st>>>(LU, ld, n, k0, cc, nb, 0, k0, X, nrhs);
        UPDES_LAUNCH_CHECK();
      }
    }
  }
  return 0;
}

static int solve_chunk(UpdesLU *h, const double *LU, double *X, int nrhs, cudaStream_t st) {
  const long long n = h->n, ld = h->ld;
  const size_t smem = DIAG_SMEM;
  int rc0 = ensure_solve_attrs();
  if (rc0) return rc0;
  const long long nblk = (n + SB - 1) / SB;
  // forward: L y = P b
  for (long long kb = 0; kb < nblk; kb++) {
    const long long k0 = kb * SB;
    const int nb = (int)((n - k0) < SB ? (n - k0) : SB);
    diag_solve_kernel<false><<<1, SB, smem, st>>>(LU, ld, n, k0, k0, nb, X, nrhs);
    UPDES_LAUNCH_CHECK();
    const long long i0 = k0 + nb;
    if (i0 < n) {
      block_update_kernel<<<(unsigned)((n - i0 + 7) / 8), 256, 0, st>>>(LU, ld, n, k0, k0, nb, i0, n, X, nrhs);
      UPDES_LAUNCH_CHECK();
    }
  }
  // backward: U x = y
  for (long long kb = nblk - 1; kb >= 0; kb--) {
    const long long k0 = kb * SB;
    const int nb = (int)((n - k0) < SB ? (n - k0) : SB);
    diag_solve_kernel<true><<<1, SB, smem, st>>>(LU, ld, n, k0, k0, nb, X, nrhs);
    UPDES_LAUNCH_CHECK();
    if (k0 > 0) {
      block_update_kernel<<<(unsigned)((k0 + 7) / 8), 256, 0, st>>>(LU, ld, n, k0, k0, nb, 0, k0, X, nrhs);
      UPDES_LAUNCH_CHECK();
    }
  }
  return 0;
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
    int rc = solve_chunk(h, LU, h->xbuf, nf, st);
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

extern "C" int updes_tri_block_sweep(UpdesLU *h, int slot, int upper, int64_t r0, int64_t c0, int64_t width,
                                     double *X, int nrhs, void *stream) {
  using namespace updes;
  if (!h) return -1;
  if (slot < 0 || slot >= UPDES_MAX_VIEWS || !h->view[slot].ptr) return -2;
  if (r0 < 0 || r0 + width > h->view[slot].rows) return -4;
  if (c0 < 0 || (c0 & 1) || c0 + width > h->view[slot].ld) return -5;
  if (!X) return -7;
  if (nrhs <= 0 || nrhs > SOLVE_MAX_RHS) return -8;
  return tri_block_sweep(h->view[slot].ptr, h->view[slot].ld, h->view[slot].rows, upper != 0, r0, c0, width, X, nrhs,
                         (cudaStream_t)stream);
}

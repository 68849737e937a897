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
__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int *p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
}  // namespace updes

namespace updes {

constexpr int SB = 128;
constexpr int SOLVE_MAX_RHS = 4;

// X = P (diag(scale) B): row equilibration factors (if any) are applied while gathering
__global__ void gather_rows_kernel(const double *B, long long ldb, int nrhs, const int32_t *perm, long long n,
                                   double *X, long long ldx, const double *scale) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int p = perm[i];
  const double sc = scale ? scale[p] : 1.0;
  for (int f = 0; f < nrhs; f++) X[f * ldx + i] = B[f * ldb + p] * sc;
}

__global__ void copy_back_kernel(double *B, long long ldb, int nrhs, long long n, const double *X, long long ldx) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int f = 0; f < nrhs; f++) B[f * ldb + i] = X[f * ldx + i];
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

// ------------------------------------------------------------------------------------------------
// Transposed solves  K^T x = b  with the same factors (adjoint / VJP of the solve).
//   P K = L U  =>  K^T = U^T L^T P :  U^T w = b (forward, lower with diagonal),
//                                     L^T z = w (backward, unit upper),  x[perm[i]] = z[i].
// One launch per 128-row block: every CTA stages the diagonal block, solves its transpose redundantly,
// then updates its share of the remaining unknowns.  With row-major factors the update
// X[i] -= sum_c F[k0+c][i] w_c runs one thread per column i (128 coalesced row reads).
// FWD = true: U^T (forward);  FWD = false: L^T (backward, unit diagonal).
// ------------------------------------------------------------------------------------------------
template <bool FWD>
__global__ void __launch_bounds__(STEP_THREADS, 1)
tri_step_t_kernel(const double *LU, long long ld, long long n, long long k0, int nb, long long i0, long long i1,
                  double *X, double *Yout, int nrhs) {
  extern __shared__ double sm[];
  double *T = sm;
  double *xs = sm + SB * (SB + 1);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NWARPS = STEP_THREADS / 32;
  for (int r = warp; r < SB; r += NWARPS) {
    const double *src = LU + (k0 + r) * ld + k0;
    for (int c = lane; c < SB; c += 32) T[r * (SB + 1) + c] = (r < nb && c < nb) ? src[c] : (r == c ? 1.0 : 0.0);
  }
  for (int t = tid; t < nrhs * SB; t += STEP_THREADS) {
    const int f = t / SB, c = t % SB;
    xs[f * SB + c] = c < nb ? X[f * n + k0 + c] : 0.0;
  }
  __syncthreads();
  // solve T^T y = xs.  FWD: T upper (U block), T^T lower with diagonal -> ascending;
  //                   !FWD: T unit lower (L block), T^T unit upper -> descending.
  for (int q = 0; q < 4; q++) {
    const int sb = FWD ? q : 3 - q;
    const int base = sb * 32;
    if (warp == 0) {
      for (int f = 0; f < nrhs; f++) {
        double x = xs[f * SB + base + lane];
        if (FWD) {
          for (int c = 0; c < 32; c++) {
            if (lane == c) x = x / T[(base + c) * (SB + 1) + base + c];
            const double xc = __shfl_sync(0xffffffffu, x, c);
            if (lane > c) x = fma(-T[(base + c) * (SB + 1) + base + lane], xc, x);     // T^T[lane][c] = T[c][lane]
          }
        } else {
          for (int c = 31; c >= 0; c--) {
            const double xc = __shfl_sync(0xffffffffu, x, c);
            if (lane < c) x = fma(-T[(base + c) * (SB + 1) + base + lane], xc, x);
          }
        }
        xs[f * SB + base + lane] = x;
      }
    }
    __syncthreads();
    const int rbeg = FWD ? base + 32 : 0, rend = FWD ? SB : base;
    for (int t = tid; t < (rend - rbeg) * nrhs; t += STEP_THREADS) {
      const int f = t / (rend - rbeg), r = rbeg + t % (rend - rbeg);
      double v = xs[f * SB + r];
      for (int c = 0; c < 32; c++) v = fma(-T[(base + c) * (SB + 1) + r], xs[f * SB + base + c], v);
      xs[f * SB + r] = v;
    }
    __syncthreads();
  }
  if (blockIdx.x == 0)
    for (int t = tid; t < nrhs * nb; t += STEP_THREADS) {
      const int f = t / nb, c = t % nb;
      Yout[f * n + k0 + c] = xs[f * SB + c];
    }
  // remaining unknowns i in [i0, i1): X[i] -= sum_c F[k0+c][i] * y_c
  for (long long i = i0 + (long long)blockIdx.x * STEP_THREADS + tid; i < i1; i += (long long)gridDim.x * STEP_THREADS) {
    const double *col = LU + k0 * ld + i;
    double acc[SOLVE_MAX_RHS] = {0, 0, 0, 0};
    for (int c0 = 0; c0 < nb; c0 += 8) {
      double v[8];
#pragma unroll
      for (int u = 0; u < 8; u++) v[u] = (c0 + u < nb) ? col[(long long)(c0 + u) * ld] : 0.0;
      for (int f = 0; f < nrhs; f++)
#pragma unroll
        for (int u = 0; u < 8; u++) acc[f] = fma(v[u], xs[f * SB + c0 + u], acc[f]);
    }
    for (int f = 0; f < nrhs; f++) X[f * n + i] -= acc[f];
  }
}

// with row equilibration  P diag(s) K = L U,  K^T = U^T L^T P diag(s)^-1,  so  x = diag(s) P^T z
__global__ void scatter_rows_kernel(const double *Z, long long n, int nrhs, const int32_t *perm, double *B, long long ldb,
                                    const double *scale) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int p = perm[i];
  const double sc = scale ? scale[p] : 1.0;
  for (int f = 0; f < nrhs; f++) B[f * ldb + p] = Z[f * n + i] * sc;
}

__global__ void copy_in_kernel(const double *B, long long ldb, int nrhs, long long n, double *X) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int f = 0; f < nrhs; f++) X[f * n + i] = B[f * ldb + i];
}

static int solve_chunk_transposed(UpdesLU *h, const double *LU, double *X, double *Y, int nrhs, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    UPDES_CUDA_TRY(cudaFuncSetAttribute(tri_step_t_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)STEP_SMEM));
    UPDES_CUDA_TRY(cudaFuncSetAttribute(tri_step_t_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)STEP_SMEM));
    attr = true;
  }
  const long long n = h->n, ld = h->ld;
  const long long nblk = (n + SB - 1) / SB;
  auto grid_for = [&](long long cols) {
    const long long want = (cols + STEP_THREADS - 1) / STEP_THREADS;
    return (int)(want < 1 ? 1 : (want > h->num_sms ? h->num_sms : want));
  };
  for (long long kb = 0; kb < nblk; kb++) {                 // U^T w = b : X -> Y
    const long long k0 = kb * SB;
    const int nb = (int)((n - k0) < SB ? (n - k0) : SB);
    tri_step_t_kernel<true><<<grid_for(n - k0 - nb), STEP_THREADS, STEP_SMEM, st>>>(LU, ld, n, k0, nb, k0 + nb, n, X, Y, nrhs);
    UPDES_LAUNCH_CHECK();
  }
  for (long long kb = nblk - 1; kb >= 0; kb--) {            // L^T z = w : Y -> X
    const long long k0 = kb * SB;
    const int nb = (int)((n - k0) < SB ? (n - k0) : SB);
    tri_step_t_kernel<false><<<grid_for(k0), STEP_THREADS, STEP_SMEM, st>>>(LU, ld, n, k0, nb, 0, k0, Y, X, nrhs);
    UPDES_LAUNCH_CHECK();
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Persistent pipelined sweep (default): ONE cooperative launch per sweep direction.
//
// The rows are cut into 128-row blocks owned cyclically by the resident CTAs (block j -> CTA j % G).
// A CTA applies every block step's update to the rows it owns; the owner of the NEXT diagonal block
// updates that block first, solves it at once (its diagonal 128x128 tile was staged in shared memory
// while it was still waiting) and publishes the solution behind a flag, and only then catches up with
// the rest of its rows.  The serial chain per step is therefore: flag -> 128 rows of update -> warp
// substitution -> publish (a few microseconds), while the other ~147 CTAs stream L / U at HBM rate in
// the background.  Diagonal 32x32 sub-blocks are held in registers (one row per lane) during the
// substitution so the chain is one shuffle + one FMA per column.
// ------------------------------------------------------------------------------------------------
constexpr int SWEEP_THREADS = 512;
constexpr size_t SWEEP_SMEM = sizeof(double) * (SB * (SB + 1) + 2 * SOLVE_MAX_RHS * SB);

struct SweepParams {
  const double *LU;
  long long ld, n;
  long long cbase;          // column of diagonal block kb = cbase + 128*kb
  int kb_begin, kb_end;     // block steps [kb_begin, kb_end) of the 128-row partition of [0, n)
  double *X, *Y;            // running right-hand side (updated in place) / solved blocks
  int nrhs;
  unsigned int *flags;      // [ceil(n/128)]: == epoch once block kb's solution is in Y
  unsigned int epoch;
  int *err;                 // set to 1 if a wait timed out (never expected)
};

__device__ __forceinline__ void st_release_u32(unsigned int *p, unsigned int v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <bool UPPER>
__device__ __forceinline__ void sweep_stage_T(const SweepParams &P, int kb, double *T) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int NW = SWEEP_THREADS / 32;
  const long long k0 = 128LL * kb;
  const int nb = (int)min(128LL, P.n - k0);
  const double *base = P.LU + k0 * P.ld + P.cbase + k0;
  if (nb == SB) {
#pragma unroll
    for (int g = 0; g < SB / NW; g += 4) {
      double v[4][4];
#pragma unroll
      for (int u = 0; u < 4; u++)
#pragma unroll
        for (int q = 0; q < 4; q++) v[u][q] = base[(long long)(warp + (g + u) * NW) * P.ld + lane + 32 * q];
#pragma unroll
      for (int u = 0; u < 4; u++)
#pragma unroll
        for (int q = 0; q < 4; q++) T[(warp + (g + u) * NW) * (SB + 1) + lane + 32 * q] = v[u][q];
    }
  } else {
    for (int r = warp; r < SB; r += NW)
      for (int c = lane; c < SB; c += 32)
        T[r * (SB + 1) + c] = (r < nb && c < nb) ? base[(long long)r * P.ld + c] : (r == c ? 1.0 : 0.0);
  }
}

// Solve the staged diagonal block against X[block kb]; result in xs (shared) and Y; publish the flag.
template <bool UPPER>
__device__ __forceinline__ void sweep_solve_block(const SweepParams &P, int kb, const double *T, double *xs) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long k0 = 128LL * kb;
  const int nb = (int)min(128LL, P.n - k0);
  for (int t = tid; t < P.nrhs * SB; t += SWEEP_THREADS) {
    const int f = t / SB, c = t % SB;
    xs[f * SB + c] = c < nb ? P.X[f * P.n + k0 + c] : 0.0;
  }
  __syncthreads();
  for (int q = 0; q < 4; q++) {
    const int sb = UPPER ? 3 - q : q;
    const int base = sb * 32;
    if (warp == 0) {
      double trow[32];
#pragma unroll
      for (int c = 0; c < 32; c++) trow[c] = T[(base + lane) * (SB + 1) + base + c];
      double rd = 1.0;
      if (UPPER) {
        double d = trow[0];
#pragma unroll
        for (int c = 1; c < 32; c++) d = lane == c ? trow[c] : d;
        rd = 1.0 / d;
      }
      for (int f = 0; f < P.nrhs; f++) {
        double x = xs[f * SB + base + lane];
        if (!UPPER) {
#pragma unroll
          for (int c = 0; c < 31; c++) {
            const double xc = __shfl_sync(0xffffffffu, x, c);
            if (lane > c) x = fma(-trow[c], xc, x);
          }
        } else {
#pragma unroll
          for (int c = 31; c >= 0; c--) {
            if (lane == c) x *= rd;
            const double xc = __shfl_sync(0xffffffffu, x, c);
            if (lane < c) x = fma(-trow[c], xc, x);
          }
        }
        xs[f * SB + base + lane] = x;
      }
    }
    __syncthreads();
    // remaining sub-blocks of this diagonal block: 4 threads per row, 8 columns each
    const int rbeg = UPPER ? 0 : base + 32, rend = UPPER ? base : SB;
    const int nrow = rend - rbeg;
    for (int t = tid; t < nrow * 4 * P.nrhs; t += SWEEP_THREADS) {
      const int part = t & 3, rr = (t >> 2) % nrow, f = (t >> 2) / nrow;
      const int r = rbeg + rr;
      const double *trow = T + r * (SB + 1) + base + part * 8;
      const double *xv = xs + f * SB + base + part * 8;
      double v = 0.0;
#pragma unroll
      for (int c = 0; c < 8; c++) v = fma(trow[c], xv[c], v);
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      if (part == 0) xs[f * SB + r] -= v;
    }
    __syncthreads();
  }
  for (int t = tid; t < P.nrhs * nb; t += SWEEP_THREADS) {
    const int f = t / nb, c = t % nb;
    P.Y[f * P.n + k0 + c] = xs[f * SB + c];
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) st_release_u32(P.flags + kb, P.epoch);
}

// X[rows of block j] -= LU[rows of block j][columns of block kb] * xs   (this CTA owns block j)
__device__ __forceinline__ void sweep_update_block(const SweepParams &P, int j, int kb, const double *xs) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int NW = SWEEP_THREADS / 32;
  constexpr int RW = SB / NW;                 // rows per warp (8)
  const long long r0 = 128LL * j, k0 = 128LL * kb;
  const int nbj = (int)min(128LL, P.n - r0), nb = (int)min(128LL, P.n - k0);
  const double *base = P.LU + r0 * P.ld + P.cbase + k0;
  if (nb == SB) {
    double2 v0[RW], v1[RW];
#pragma unroll
    for (int u = 0; u < RW; u++) {
      const int r = warp + u * NW;
      if (r < nbj) {
        const double *row = base + (long long)r * P.ld;
        v0[u] = *reinterpret_cast<const double2 *>(row + 2 * lane);
        v1[u] = *reinterpret_cast<const double2 *>(row + 64 + 2 * lane);
      } else {
        v0[u] = make_double2(0.0, 0.0); v1[u] = v0[u];
      }
    }
    for (int f = 0; f < P.nrhs; f++) {
      const double *xf = xs + f * SB;
      const double x0 = xf[2 * lane], x1 = xf[2 * lane + 1], x2 = xf[64 + 2 * lane], x3 = xf[64 + 2 * lane + 1];
      double a[RW];
#pragma unroll
      for (int u = 0; u < RW; u++) a[u] = fma(v1[u].y, x3, fma(v1[u].x, x2, fma(v0[u].y, x1, v0[u].x * x0)));
#pragma unroll
      for (int off = 16; off > 0; off >>= 1)
#pragma unroll
        for (int u = 0; u < RW; u++) a[u] += __shfl_xor_sync(0xffffffffu, a[u], off);
      if (lane < RW) {
        double mine = a[0];
#pragma unroll
        for (int u = 1; u < RW; u++) mine = lane == u ? a[u] : mine;
        const int r = warp + lane * NW;
        if (r < nbj) P.X[f * P.n + r0 + r] -= mine;
      }
    }
  } else {
    for (int r = warp; r < nbj; r += NW) {
      const double *row = base + (long long)r * P.ld;
      for (int f = 0; f < P.nrhs; f++) {
        double a = 0.0;
        for (int c = lane; c < nb; c += 32) a = fma(row[c], xs[f * SB + c], a);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) a += __shfl_xor_sync(0xffffffffu, a, off);
        if (lane == 0) P.X[f * P.n + r0 + r] -= a;
      }
    }
  }
}

template <bool UPPER>
__global__ void __launch_bounds__(SWEEP_THREADS, 1) tri_sweep_kernel(SweepParams P) {
  extern __shared__ double sm[];
  double *T = sm;                                   // [SB][SB+1]
  double *xsbuf = sm + SB * (SB + 1);               // [2][SOLVE_MAX_RHS][SB]
  const int G = gridDim.x, me = blockIdx.x, tid = threadIdx.x;
  const int nblk = (int)((P.n + SB - 1) / SB);
  const int nsteps = P.kb_end - P.kb_begin;
  int cur = 0;
  bool have_local = false;
  {
    const int kb0 = UPPER ? P.kb_end - 1 : P.kb_begin;
    if (kb0 % G == me) {                            // the first diagonal block has no predecessor
      sweep_stage_T<UPPER>(P, kb0, T);
      __syncthreads();
      sweep_solve_block<UPPER>(P, kb0, T, xsbuf);
      have_local = true;
    }
  }
  for (int t = 0; t < nsteps; t++) {
    const int kb = UPPER ? P.kb_end - 1 - t : P.kb_begin + t;
    const int kn = UPPER ? kb - 1 : kb + 1;
    const bool next_mine = (t + 1 < nsteps) && (kn % G == me);
    double *xk = xsbuf + cur * SOLVE_MAX_RHS * SB;
    if (next_mine) sweep_stage_T<UPPER>(P, kn, T);  // before waiting: hides the staging latency
    if (!have_local) {
      if (tid == 0) {
        unsigned int polls = 0;
        while (ld_acquire_u32(P.flags + kb) != P.epoch) {
          if (++polls > (1u << 24)) { atomicExch(P.err, 1); break; }
        }
      }
      __syncthreads();
      const int nb = (int)min(128LL, P.n - 128LL * kb);
      for (int i = tid; i < P.nrhs * SB; i += SWEEP_THREADS) {
        const int f = i / SB, c = i % SB;
        xk[f * SB + c] = c < nb ? __ldcg(P.Y + f * P.n + 128LL * kb + c) : 0.0;
      }
    }
    __syncthreads();
    have_local = false;
    if (next_mine) {
      sweep_update_block(P, kn, kb, xk);
      __threadfence_block();
      __syncthreads();
      sweep_solve_block<UPPER>(P, kn, T, xsbuf + (cur ^ 1) * SOLVE_MAX_RHS * SB);
      have_local = true;
    }
    // the rest of my rows (forward: blocks after kb; backward: blocks before kb)
    if (!UPPER) {
      int j = kb + 1 + (((me - (kb + 1)) % G) + G) % G;
      for (; j < nblk; j += G)
        if (!(next_mine && j == kn)) sweep_update_block(P, j, kb, xk);
    } else {
      for (int j = me; j < kb; j += G)
        if (!(next_mine && j == kn)) sweep_update_block(P, j, kb, xk);
    }
    __syncthreads();
    cur ^= 1;
  }
}

// ------------------------------------------------------------------------------------------------
// Row-block streaming sweep (default, solve_variant 2): ONE launch per sweep direction.
//
// The step-synchronous kernel above makes every CTA wait for block step kb before touching any of its
// tiles of that step; its serial chain (flag -> update -> substitution -> publish, ~6.5 us per 128-row
// block) therefore gates ALL traffic and the sweep ran at 52 % of HBM.  Here the work is cut the other
// way: a CTA takes one 128-row block j at a time (dynamic ticket, handed out in dependency order) and
// streams that block's row panel -- tiles (j, kb) for every block step kb before it -- accumulating
// X[j] - sum_kb L[j,kb] x_kb in registers.  Only the LAST tile of a block is on the critical path; all
// earlier ones have slack, so while one CTA at a time finishes "its" diagonal solve, the other ~147
// keep streaming L / U at HBM rate.  No block step ever waits for stragglers.
//   * solved blocks are published through the data itself: Y is pre-filled with a signalling-NaN
//     sentinel, the solver overwrites it, consumers (one warp per CTA) poll the 1 KB they need and
//     re-broadcast it through shared memory -- one L2 round trip, no flag, no fence on the chain;
//   * deadlock-free without co-residency: tickets are taken in dependency order by running CTAs, so
//     every block a CTA waits for is held by a CTA that is already executing;
//   * the diagonal tile is staged in shared memory when the block starts (long before its turn), the
//     32x32 triangular sub-solves keep one row per lane in registers: plain substitution, no inverses.
// The same kernel serves the multi-GPU path: block steps restricted to [kb_begin, kb_end) (the column
// block this rank owns); row blocks past kb_end only get their X updated.
// ------------------------------------------------------------------------------------------------
constexpr int SW2_THREADS = 512;
constexpr int SW2_NW = SW2_THREADS / 32;
constexpr int SW2_RW = SB / SW2_NW;            // rows per warp (8)
constexpr unsigned long long SW2_SENTINEL = 0x7FF4DEADBEEF5EEDull;   // signalling NaN never produced by arithmetic

template <int NR>
struct Sw2Smem {
  double T[SB * (SB + 1)];
  double xs[NR][SB];            // right-hand side / solution of the diagonal block being solved
  double sx[2][NR][SB];         // x_kb of the tile in flight (double-buffered)
  int ticket;
};

struct Sweep2Params {
  const double *LU;
  long long ld, n;
  long long cbase;              // column of diagonal block kb = cbase + 128 kb
  int kb_begin, kb_end;         // block steps available in this factor
  int nblk;                     // ceil(n / 128)
  double *X, *Y;
  long long ldx;                // stride between right-hand sides in X and Y (even when there is more than one)
  unsigned int *ticket;         // dynamic row-block counter (monotonic across launches)
  unsigned int ticket_base;
  int *err;
  int diag_only;                // 1: only the row blocks of the column block itself (no update of the rows beyond it)
};

__device__ __forceinline__ double2 ld_relaxed_f64x2(const double *p) {
  double2 v;
  asm volatile("ld.relaxed.gpu.global.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ bool is_sentinel(double v) { return (unsigned long long)__double_as_longlong(v) == SW2_SENTINEL; }

__global__ void fill_sentinel_kernel(double *Y, long long ldx, long long r0, long long cnt, int nrhs) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= cnt * nrhs) return;
  const long long f = t / cnt, c = t % cnt;
  Y[f * ldx + r0 + c] = __longlong_as_double((long long)SW2_SENTINEL);
}

template <bool UPPER, int NR>
__global__ void __launch_bounds__(SW2_THREADS, 1) tri_sweep2_kernel(Sweep2Params P) {
  extern __shared__ __align__(16) unsigned char sw2_raw[];
  Sw2Smem<NR> &S = *reinterpret_cast<Sw2Smem<NR> *>(sw2_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nsteps = P.kb_end - P.kb_begin;
  const int nrowblocks = P.diag_only ? nsteps : (UPPER ? P.kb_end : P.nblk - P.kb_begin);
  const long long n = P.n, ldx = P.ldx;

  while (true) {
    __syncthreads();                                       // previous block's shared state is dead
    if (tid == 0) S.ticket = (int)(atomicAdd(P.ticket, 1u) - P.ticket_base);
    __syncthreads();
    const int i = S.ticket;                                // order index of my row block
    if (i >= nrowblocks) break;
    const int j = UPPER ? P.kb_end - 1 - i : P.kb_begin + i;
    const bool diag = i < nsteps;
    const int ntiles = diag ? i : nsteps;
    const long long r0 = 128LL * j;
    const int nbj = (int)min(128LL, n - r0);

    if (diag) {                                            // stage the diagonal tile: off the critical path
      const double *base = P.LU + r0 * P.ld + P.cbase + r0;
      if (nbj == SB) {
#pragma unroll
        for (int g = 0; g < SB / SW2_NW; g += 4) {
          double v[4][4];
#pragma unroll
          for (int u = 0; u < 4; u++)
#pragma unroll
            for (int q = 0; q < 4; q++) v[u][q] = base[(long long)(warp + (g + u) * SW2_NW) * P.ld + lane + 32 * q];
#pragma unroll
          for (int u = 0; u < 4; u++)
#pragma unroll
            for (int q = 0; q < 4; q++) S.T[(warp + (g + u) * SW2_NW) * (SB + 1) + lane + 32 * q] = v[u][q];
        }
      } else {
        for (int r = warp; r < SB; r += SW2_NW)
          for (int c = lane; c < SB; c += 32)
            S.T[r * (SB + 1) + c] = (r < nbj && c < nbj) ? base[(long long)r * P.ld + c] : (r == c ? 1.0 : 0.0);
      }
    }

    // right-hand side of my rows, fetched now so that it is not on the critical path after the last tile
    double xpre[NR];
#pragma unroll
    for (int f = 0; f < NR; f++) {
      const int r = warp + lane * SW2_NW;
      xpre[f] = (diag && lane < SW2_RW && r < nbj) ? P.X[(long long)f * ldx + r0 + r] : 0.0;
    }

    // ---- stream the row panel: acc[u][f] = partial sums of row (warp + u*NW) over my 4 columns per tile ----
    double acc[SW2_RW][NR];
#pragma unroll
    for (int u = 0; u < SW2_RW; u++)
#pragma unroll
      for (int f = 0; f < NR; f++) acc[u][f] = 0.0;

    for (int t = 0; t < ntiles; t++) {
      const int kb = UPPER ? P.kb_end - 1 - t : P.kb_begin + t;
      const long long k0 = 128LL * kb;
      const int nb = (int)min(128LL, n - k0);
      const double *tile = P.LU + r0 * P.ld + P.cbase + k0;
      // NR == 1: the whole tile (8 rows per warp) is issued before waiting for x_kb; with 4 right-hand sides
      // the accumulators take the registers, so the tile goes through in two halves of 4 rows per warp
      constexpr int HR = NR == 1 ? SW2_RW : SW2_RW / 2;
      double2 v0[HR], v1[HR];
      auto load_rows = [&](int ubase) {
        if (nb == SB && nbj == SB) {
#pragma unroll
          for (int u = 0; u < HR; u++) {
            const double *row = tile + (long long)(warp + (ubase + u) * SW2_NW) * P.ld;
            v0[u] = *reinterpret_cast<const double2 *>(row + 2 * lane);
            v1[u] = *reinterpret_cast<const double2 *>(row + 64 + 2 * lane);
          }
        } else {
#pragma unroll
          for (int u = 0; u < HR; u++) {
            const int r = warp + (ubase + u) * SW2_NW;
            const double *row = tile + (long long)r * P.ld;
            const bool rv = r < nbj;
            v0[u].x = (rv && 2 * lane < nb) ? row[2 * lane] : 0.0;
            v0[u].y = (rv && 2 * lane + 1 < nb) ? row[2 * lane + 1] : 0.0;
            v1[u].x = (rv && 64 + 2 * lane < nb) ? row[64 + 2 * lane] : 0.0;
            v1[u].y = (rv && 65 + 2 * lane < nb) ? row[65 + 2 * lane] : 0.0;
          }
        }
      };
      load_rows(0);
      // x_kb: warp 0 polls the published block (sentinel = not solved yet) and re-broadcasts it
      double (*sx)[SB] = S.sx[t & 1];
      if (warp == 0) {
        unsigned int polls = 0;
#pragma unroll
        for (int f = 0; f < NR; f++) {
          const double *yb = P.Y + (long long)f * ldx + k0;
          double2 a = make_double2(0.0, 0.0), b = make_double2(0.0, 0.0);
          while (true) {
            bool ok = true;
            if (nb == SB) {
              a = ld_relaxed_f64x2(yb + 2 * lane);
              b = ld_relaxed_f64x2(yb + 64 + 2 * lane);
              ok = !(is_sentinel(a.x) || is_sentinel(a.y) || is_sentinel(b.x) || is_sentinel(b.y));
            } else {
              volatile const double *yv = yb;
              a.x = 2 * lane < nb ? yv[2 * lane] : 0.0;
              a.y = 2 * lane + 1 < nb ? yv[2 * lane + 1] : 0.0;
              b.x = 64 + 2 * lane < nb ? yv[64 + 2 * lane] : 0.0;
              b.y = 65 + 2 * lane < nb ? yv[65 + 2 * lane] : 0.0;
              ok = !(is_sentinel(a.x) || is_sentinel(a.y) || is_sentinel(b.x) || is_sentinel(b.y));
            }
            if (__all_sync(0xffffffffu, ok)) break;
            if (++polls > (1u << 22)) { if (lane == 0) atomicExch(P.err, 1); break; }   // never expected
          }
          *reinterpret_cast<double2 *>(&sx[f][2 * lane]) = a;
          *reinterpret_cast<double2 *>(&sx[f][64 + 2 * lane]) = b;
        }
      }
      __syncthreads();
#pragma unroll
      for (int half = 0; half < SW2_RW / HR; half++) {
        if (half > 0) load_rows(half * HR);
#pragma unroll
        for (int f = 0; f < NR; f++) {
          const double2 xa = *reinterpret_cast<const double2 *>(&sx[f][2 * lane]);
          const double2 xb = *reinterpret_cast<const double2 *>(&sx[f][64 + 2 * lane]);
#pragma unroll
          for (int u = 0; u < HR; u++)
            acc[half * HR + u][f] = fma(v1[u].y, xb.y, fma(v1[u].x, xb.x, fma(v0[u].y, xa.y, fma(v0[u].x, xa.x, acc[half * HR + u][f]))));
        }
      }
    }

    // ---- reduce across lanes: lane u (< RW) ends up with the sum of row warp + u*NW ----------------------
#pragma unroll
    for (int f = 0; f < NR; f++) {
#pragma unroll
      for (int off = 16; off > 0; off >>= 1)
#pragma unroll
        for (int u = 0; u < SW2_RW; u++) acc[u][f] += __shfl_xor_sync(0xffffffffu, acc[u][f], off);
      if (lane < SW2_RW) {
        double mine = acc[0][f];
#pragma unroll
        for (int u = 1; u < SW2_RW; u++) mine = lane == u ? acc[u][f] : mine;
        const int r = warp + lane * SW2_NW;
        if (r < nbj) {
          double *xp = P.X + (long long)f * ldx + r0 + r;
          if (diag) S.xs[f][r] = xpre[f] - mine;
          else *xp -= mine;
        } else if (diag) {
          S.xs[f][r] = 0.0;
        }
      }
    }
    if (!diag) continue;
    __syncthreads();

    // ---- triangular solve of the diagonal block (substitution, 32x32 sub-blocks in registers) ------------
    for (int q = 0; q < 4; q++) {
      const int sb = UPPER ? 3 - q : q;
      const int base = sb * 32;
      if (warp == 0) {
        double trow[32];
#pragma unroll
        for (int c = 0; c < 32; c++) trow[c] = S.T[(base + lane) * (SB + 1) + base + c];
        double rd = 1.0;
        if (UPPER) {
          double d = trow[0];
#pragma unroll
          for (int c = 1; c < 32; c++) d = lane == c ? trow[c] : d;
          rd = 1.0 / d;
        }
#pragma unroll
        for (int f = 0; f < NR; f++) {
          double x = S.xs[f][base + lane];
          if (!UPPER) {
#pragma unroll
            for (int c = 0; c < 31; c++) {
              const double xc = __shfl_sync(0xffffffffu, x, c);
              if (lane > c) x = fma(-trow[c], xc, x);
            }
          } else {
#pragma unroll
            for (int c = 31; c >= 0; c--) {
              if (lane == c) x *= rd;
              const double xc = __shfl_sync(0xffffffffu, x, c);
              if (lane < c) x = fma(-trow[c], xc, x);
            }
          }
          S.xs[f][base + lane] = x;
          // publish this sub-block at once: consumers poll the data itself
          if (base + lane < nbj) __stcg(P.Y + (long long)f * ldx + r0 + base + lane, x);
        }
      }
      __syncthreads();
      if (q == 3) break;
      // remaining sub-blocks of this diagonal block: 4 threads per row, 8 columns each
      const int rbeg = UPPER ? 0 : base + 32, rend = UPPER ? base : SB;
      const int nrow = rend - rbeg;
      for (int t = tid; t < nrow * 4 * NR; t += SW2_THREADS) {
        const int part = t & 3, rr = (t >> 2) % nrow, f = (t >> 2) / nrow;
        const int r = rbeg + rr;
        const double *trow = S.T + r * (SB + 1) + base + part * 8;
        const double *xv = &S.xs[f][base + part * 8];
        double v = 0.0;
#pragma unroll
        for (int c = 0; c < 8; c++) v = fma(trow[c], xv[c], v);
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        if (part == 0) S.xs[f][r] -= v;
      }
      __syncthreads();
    }
  }
}

template <bool UPPER, int NR>
static int launch_sweep2(UpdesLU *h, Sweep2Params &P, int grid, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    UPDES_CUDA_TRY(cudaFuncSetAttribute(tri_sweep2_kernel<UPPER, NR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)sizeof(Sw2Smem<NR>)));
    attr = true;
  }
  tri_sweep2_kernel<UPPER, NR><<<grid, SW2_THREADS, sizeof(Sw2Smem<NR>), st>>>(P);
  UPDES_LAUNCH_CHECK();
  (void)h;
  return 0;
}

// Block steps [kb_begin, kb_end) of the factor stored at `LU` (diagonal block kb at column cbase + 128 kb).
// X is the running right-hand side (rows past the solved range are updated in place), Y receives the
// solved blocks.  nrhs <= SOLVE_MAX_RHS.
int tri_sweep_rowblock(UpdesLU *h, const double *LU, long long ld, long long n, bool upper, long long cbase,
                       int kb_begin, int kb_end, double *X, double *Y, long long ldx, int nrhs, cudaStream_t st,
                       bool diag_only = false) {
  if (kb_end <= kb_begin) return 0;
  if (nrhs > 1 && ((ldx & 1) || (((uintptr_t)X | (uintptr_t)Y) & 15))) return -9;   // 16-byte accesses to x blocks
  const int nblk = (int)((n + SB - 1) / SB);
  const long long y0 = 128LL * kb_begin;
  const long long ycnt = (128LL * kb_end < n ? 128LL * kb_end : n) - y0;
  const int nr = nrhs > 1 ? SOLVE_MAX_RHS : 1;
  fill_sentinel_kernel<<<(unsigned)((ycnt * nr + 255) / 256), 256, 0, st>>>(Y, ldx, y0, ycnt, nr);
  UPDES_LAUNCH_CHECK();
  Sweep2Params P;
  P.LU = LU; P.ld = ld; P.n = n; P.cbase = cbase; P.kb_begin = kb_begin; P.kb_end = kb_end; P.nblk = nblk;
  P.X = X; P.Y = Y; P.ldx = ldx; P.ticket = h->sweep_ticket; P.ticket_base = h->sweep_ticket_count; P.err = h->sweep_err;
  P.diag_only = diag_only ? 1 : 0;
  const int nrowblocks = diag_only ? kb_end - kb_begin : (upper ? kb_end : nblk - kb_begin);
  const int grid = nrowblocks < h->num_sms ? nrowblocks : h->num_sms;
  h->sweep_ticket_count += (unsigned int)(nrowblocks + grid);      // every CTA draws one ticket past the end
  if (nr == 1) return upper ? launch_sweep2<true, 1>(h, P, grid, st) : launch_sweep2<false, 1>(h, P, grid, st);
  return upper ? launch_sweep2<true, SOLVE_MAX_RHS>(h, P, grid, st) : launch_sweep2<false, SOLVE_MAX_RHS>(h, P, grid, st);
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
                    long long w, double *X, double *Y, int nrhs, cudaStream_t st, bool diag_only = false) {
  const long long row_lo = diag_only ? r0 : 0, row_hi = diag_only ? r0 + w : n;   // rows whose X may be updated
  int rc = ensure_solve_attrs();
  if (rc) return rc;
  const long long nsub = (w + SB - 1) / SB;
  for (long long t = 0; t < nsub; t++) {
    const long long s = upper ? (nsub - 1 - t) : t;
    const long long k0 = r0 + s * SB, cc = c0 + s * SB;
    const int nb = (int)((w - s * SB) < SB ? (w - s * SB) : SB);
    if (!upper) {
      const long long i0 = k0 + nb;
      tri_step_kernel<false><<<step_grid(num_sms, row_hi - i0), STEP_THREADS, STEP_SMEM, st>>>(LU, ld, n, k0, cc, nb, i0,
                                                                                              row_hi, X, Y, nrhs);
    } else {
      tri_step_kernel<true><<<step_grid(num_sms, k0 - row_lo), STEP_THREADS, STEP_SMEM, st>>>(LU, ld, n, k0, cc, nb, row_lo,
                                                                                             k0, X, Y, nrhs);
    }
    UPDES_LAUNCH_CHECK();
  }
  return 0;
}

// One cooperative launch: block steps [kb_begin, kb_end) of the factor stored at `LU` (diagonal block
// kb at column cbase + 128 kb).  X is the running right-hand side, Y receives the solved blocks.
int tri_sweep_persistent(UpdesLU *h, const double *LU, long long ld, long long n, bool upper, long long cbase,
                         int kb_begin, int kb_end, double *X, double *Y, int nrhs, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    UPDES_CUDA_TRY(cudaFuncSetAttribute(tri_sweep_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SWEEP_SMEM));
    UPDES_CUDA_TRY(cudaFuncSetAttribute(tri_sweep_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SWEEP_SMEM));
    attr = true;
  }
  if (kb_end <= kb_begin) return 0;
  SweepParams P;
  P.LU = LU; P.ld = ld; P.n = n; P.cbase = cbase; P.kb_begin = kb_begin; P.kb_end = kb_end; P.X = X; P.Y = Y;
  P.nrhs = nrhs; P.flags = h->sweep_flags; P.epoch = ++h->sweep_epoch; P.err = h->sweep_err;
  const int nblk = (int)((n + SB - 1) / SB);
  const int grid = nblk < h->num_sms ? nblk : h->num_sms;
  void *args[] = {&P};
  cudaError_t e = upper ? cudaLaunchCooperativeKernel((void *)tri_sweep_kernel<true>, dim3(grid), dim3(SWEEP_THREADS), args, SWEEP_SMEM, st)
                        : cudaLaunchCooperativeKernel((void *)tri_sweep_kernel<false>, dim3(grid), dim3(SWEEP_THREADS), args, SWEEP_SMEM, st);
  if (e != cudaSuccess) return (int)e;
  ++g_launch_count;
  return 0;
}

// whole solve on one GPU: X = P b in xbuf[0]; forward sweep X -> Y; backward sweep Y -> X
static int solve_chunk(UpdesLU *h, const double *LU, double *X, double *Y, int nrhs, cudaStream_t st) {
  const long long n = h->n;
  const int nblk = (int)((n + SB - 1) / SB);
  if (h->solve_variant == 0) {
    int rc = tri_block_sweep(h->num_sms, LU, h->ld, n, false, 0, 0, n, X, Y, nrhs, st);
    if (rc) return rc;
    return tri_block_sweep(h->num_sms, LU, h->ld, n, true, 0, 0, n, Y, X, nrhs, st);
  }
  if (h->solve_variant == 2) {
    const long long ldx = (n + 1) & ~1LL;
    int rc = tri_sweep_rowblock(h, LU, h->ld, n, false, 0, 0, nblk, X, Y, ldx, nrhs, st);
    if (rc) return rc;
    return tri_sweep_rowblock(h, LU, h->ld, n, true, 0, 0, nblk, Y, X, ldx, nrhs, st);
  }
  int rc = tri_sweep_persistent(h, LU, h->ld, n, false, 0, 0, nblk, X, Y, nrhs, st);
  if (rc) return rc;
  return tri_sweep_persistent(h, LU, h->ld, n, true, 0, 0, nblk, Y, X, nrhs, st);
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
  cudaStream_t st = (cudaStream_t)stream;
  const long long n = h->n;
  for (int f0 = 0; f0 < nrhs; f0 += SOLVE_MAX_RHS) {
    const int nf = (nrhs - f0) < SOLVE_MAX_RHS ? (nrhs - f0) : SOLVE_MAX_RHS;
    double *Bf = B + (long long)f0 * ldb;
    if (transpose) {
      double *Xt = h->xbuf, *Yt = h->xbuf + (size_t)SOLVE_MAX_RHS * n;
      copy_in_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(Bf, ldb, nf, n, Xt);
      UPDES_LAUNCH_CHECK();
      int rct = solve_chunk_transposed(h, LU, Xt, Yt, nf, st);
      if (rct) return rct;
      scatter_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(Xt, n, nf, h->perm, Bf, ldb, h->row_scale);
      UPDES_LAUNCH_CHECK();
      continue;
    }
    // the streaming sweeps read x blocks with 16-byte accesses: right-hand sides sit at an even stride there
    const long long ldx = h->solve_variant == 2 ? ((n + 1) & ~1LL) : n;
    gather_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(Bf, ldb, nf, h->perm, n, h->xbuf, ldx, h->row_scale);
    UPDES_LAUNCH_CHECK();
    prof_begin(PROF_SOLVE, 8.0 * (double)n * (double)n, st);
    int rc = solve_chunk(h, LU, h->xbuf, h->xbuf + (size_t)SOLVE_MAX_RHS * ldx, nf, st);
    prof_end(st);
    if (rc) return rc;
    copy_back_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(Bf, ldb, nf, n, h->xbuf, ldx);
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
  gather_rows_kernel<<<(unsigned)((h->n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(B, ldb, nrhs, h->perm, h->n, X, h->n, h->row_scale);
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
  int rc;
  // the persistent kernel works on the 128-row partition of [0, n): the column block must be aligned to it
  const bool aligned = (r0 % SB) == 0 && ((width % SB) == 0 || r0 + width == n);
  if (h->solve_variant == 2 && aligned && nrhs == 1)
    rc = tri_sweep_rowblock(h, h->view[slot].ptr, h->view[slot].ld, n, upper != 0, c0 - r0, (int)(r0 / SB),
                            (int)((r0 + width + SB - 1) / SB), X, Y, n, nrhs, st);
  else if (h->solve_variant != 0 && aligned)
    rc = tri_sweep_persistent(h, h->view[slot].ptr, h->view[slot].ld, n, upper != 0, c0 - r0, (int)(r0 / SB),
                              (int)((r0 + width + SB - 1) / SB), X, Y, nrhs, st);
  else
    rc = tri_block_sweep(h->num_sms, h->view[slot].ptr, h->view[slot].ld, n, upper != 0, r0, c0, width, X, Y, nrhs, st);
  if (rc) return rc;
  copy_block_kernel<<<(unsigned)((width * nrhs + 255) / 256), 256, 0, st>>>(X, Y, n, r0, width, nrhs);
  UPDES_LAUNCH_CHECK();
  return 0;
}

// ---- left-looking distributed substitution: partial products + diagonal-block solves ---------------------------
// out[i] = sum_{c in [c_lo, c_hi)} A[r0 + i][c] * x[c]: one CTA per row, the row segment is contiguous in the row-major
// local matrix (16-byte loads, four per thread in flight); HBM-read bound, (c_hi - c_lo) * 8 bytes per row.
namespace updes {
constexpr int GEMV_THREADS = 128;
__global__ void __launch_bounds__(GEMV_THREADS) block_gemv_kernel(const double *__restrict__ A, long long ld, long long r0,
                                                                  long long c_lo, long long c_hi,
                                                                  const double *__restrict__ x, double *__restrict__ out) {
  const double *row = A + (r0 + blockIdx.x) * ld;
  const int tid = threadIdx.x;
  double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
  const long long c_even = c_lo + ((c_hi - c_lo) & ~1LL);          // c_lo is even (caller checks): 16-byte aligned pairs
  long long c = c_lo + 2 * tid;
  for (; c + 6 * GEMV_THREADS < c_even; c += 8 * GEMV_THREADS) {
    const double2 a0 = *reinterpret_cast<const double2 *>(row + c), a1 = *reinterpret_cast<const double2 *>(row + c + 2 * GEMV_THREADS);
    const double2 a2 = *reinterpret_cast<const double2 *>(row + c + 4 * GEMV_THREADS), a3 = *reinterpret_cast<const double2 *>(row + c + 6 * GEMV_THREADS);
    const double2 x0 = *reinterpret_cast<const double2 *>(x + c), x1 = *reinterpret_cast<const double2 *>(x + c + 2 * GEMV_THREADS);
    const double2 x2 = *reinterpret_cast<const double2 *>(x + c + 4 * GEMV_THREADS), x3 = *reinterpret_cast<const double2 *>(x + c + 6 * GEMV_THREADS);
    acc0 = fma(a0.y, x0.y, fma(a0.x, x0.x, acc0)); acc1 = fma(a1.y, x1.y, fma(a1.x, x1.x, acc1));
    acc2 = fma(a2.y, x2.y, fma(a2.x, x2.x, acc2)); acc3 = fma(a3.y, x3.y, fma(a3.x, x3.x, acc3));
  }
  for (; c < c_even; c += 2 * GEMV_THREADS) {
    const double2 a0 = *reinterpret_cast<const double2 *>(row + c);
    const double2 x0 = *reinterpret_cast<const double2 *>(x + c);
    acc0 = fma(a0.y, x0.y, fma(a0.x, x0.x, acc0));
  }
  if (tid == 0 && c_even < c_hi) acc1 = fma(row[c_even], x[c_even], acc1);
  double v = (acc0 + acc1) + (acc2 + acc3);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  __shared__ double part[GEMV_THREADS / 32];
  if ((tid & 31) == 0) part[tid >> 5] = v;
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < GEMV_THREADS / 32; w++) t += part[w];
    out[blockIdx.x] = t;
  }
}
}  // namespace updes

/* out[i] = sum_{c_lo <= c < c_hi} A[r0+i][c] x[c] for i < nrows, A = the buffer bound to `slot`; x is indexed by the
 * LOCAL column (length >= c_hi); an empty column range writes zeros */
extern "C" int updes_block_gemv(UpdesLU *h, int slot, int64_t r0, int64_t nrows, int64_t c_lo, int64_t c_hi,
                                const double *x, double *out, void *stream) {
  using namespace updes;
  if (!h) return -1;
  if (slot < 0 || slot >= UPDES_MAX_VIEWS || !h->view[slot].ptr) return -2;
  if (r0 < 0 || nrows < 0 || r0 + nrows > h->view[slot].rows) return -3;
  if (c_lo < 0 || (c_lo & 1) || c_hi > h->view[slot].ld) return -5;
  if (!x || (((uintptr_t)x) & 15)) return -7;
  if (!out) return -8;
  if (nrows == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (c_hi <= c_lo) {
    UPDES_CUDA_TRY(cudaMemsetAsync(out, 0, sizeof(double) * (size_t)nrows, st));
    return 0;
  }
  prof_begin(PROF_SOLVE, 8.0 * (double)nrows * (double)(c_hi - c_lo), st);
  block_gemv_kernel<<<(unsigned)nrows, GEMV_THREADS, 0, st>>>(h->view[slot].ptr, h->view[slot].ld, r0, c_lo, c_hi, x, out);
  prof_end(st);
  UPDES_LAUNCH_CHECK();
  return 0;
}

/* Solve ONLY the width x width triangular diagonal block at rows [r0, r0+width), columns [c0, c0+width) of `slot`,
 * in place on X[r0 .. r0+width) (X: full-length vector; rows outside the block are not touched).  upper = 0: unit
 * lower, 1: upper with diagonal. */
extern "C" int updes_tri_diag_solve(UpdesLU *h, int slot, int upper, int64_t r0, int64_t c0, int64_t width, double *X,
                                    void *stream) {
  using namespace updes;
  if (!h) return -1;
  if (slot < 0 || slot >= UPDES_MAX_VIEWS || !h->view[slot].ptr) return -2;
  if (r0 < 0 || width <= 0 || r0 + width > h->view[slot].rows) return -4;
  if (c0 < 0 || (c0 & 1) || c0 + width > h->view[slot].ld) return -5;
  if (!X) return -7;
  cudaStream_t st = (cudaStream_t)stream;
  const long long n = h->view[slot].rows;
  double *Y = h->xbuf + (size_t)SOLVE_MAX_RHS * h->n;     // scratch for the solved block
  const bool aligned = (r0 % SB) == 0 && ((width % SB) == 0 || r0 + width == n);
  int rc;
  prof_begin(PROF_SOLVE, 4.0 * (double)width * (double)width, st);
  if (h->solve_variant == 2 && aligned)
    rc = tri_sweep_rowblock(h, h->view[slot].ptr, h->view[slot].ld, n, upper != 0, c0 - r0, (int)(r0 / SB),
                            (int)((r0 + width + SB - 1) / SB), X, Y, n, 1, st, true);
  else
    rc = tri_block_sweep(h->num_sms, h->view[slot].ptr, h->view[slot].ld, n, upper != 0, r0, c0, width, X, Y, 1, st, true);
  if (!rc) copy_block_kernel<<<(unsigned)((width + 255) / 256), 256, 0, st>>>(X, Y, n, r0, width, 1);
  prof_end(st);
  if (rc) return rc;
  UPDES_LAUNCH_CHECK();
  return 0;
}

extern "C" int updes_lu_set_solve_variant(UpdesLU *handle, int variant) {
  if (!handle) return -1;
  if (variant < 0 || variant > 2) return -2;
  handle->solve_variant = variant;
  return 0;
}

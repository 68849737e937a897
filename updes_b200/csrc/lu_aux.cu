// LU helpers on the row-major matrix: row interchanges, unit-lower triangular solve of a block
// row (U12 = L11^-1 A12), pivot list -> permutation.
#include "lu.cuh"

namespace updes {

// ---- row interchanges -------------------------------------------------------------------------
// Applies interchanges (k0+t <-> ipiv[k0+t]), t = 0..npiv-1, in order, to columns [c0, c0+ncols).
// A sequence of <= 32 interchanges touches <= 64 rows; executing it literally is a chain of 32 dependent
// load/store round trips per thread (24 us per launch at 8k columns, ~20x the traffic time).  Instead warp 0
// first composes the sequence into a permutation of the involved rows (32 steps of register/shuffle work),
// then every thread GATHERS its column of all moved rows (all loads independent and in flight together) and
// scatters them back: one memory round trip, fully coalesced in the row-major layout.
constexpr int SWAP_THREADS = 128;
constexpr int SWAP_MAX_PIV = 32;

// Columns: the range [c0, c0+ncols) with the hole [hole0, hole0+holew) left out (the panel's own columns, which the
// panel kernel already wrote in final row order): one launch covers the columns left AND right of a base panel.
// The kernel walks through ALL npiv_total interchanges in batches of 32: a thread owns one column for the whole launch,
// so consecutive batches need no synchronisation beyond the two barriers around the shared slot tables (the
// multi-GPU path applies 512 - 2048 interchanges per panel to every local column: one launch instead of 16 - 64).
__global__ void __launch_bounds__(SWAP_THREADS) swap_rows_kernel(double *K, long long ld, long long c0, long long ncols,
                                                                long long hole0, long long holew,
                                                                long long k_first, int npiv_total, const int32_t *ipiv) {
  // slots 0..31: the diagonal rows k0+s; slots 32..63: pivot rows outside [k0, k0+npiv), in order of first use
  __shared__ int s_row[64];      // row held by a slot (-1: unused)
  __shared__ int s_src[64];      // slot whose ORIGINAL content ends up in this slot
  const long long tcol = (long long)blockIdx.x * SWAP_THREADS + threadIdx.x;
  long long c = c0 + tcol;
  if (c >= hole0) c += holew;
  const bool active = tcol < ncols - holew;
  for (int b0 = 0; b0 < npiv_total; b0 += SWAP_MAX_PIV) {
  const long long k0 = k_first + b0;
  const int npiv = min(SWAP_MAX_PIV, npiv_total - b0);
  if (b0 > 0) __syncthreads();   // every thread has read the previous batch's tables
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    // all pivots in ONE load (a load per loop step is 32 dependent global round trips: ~10 us of a ~15 us kernel)
    const int mypiv = lane < npiv ? ipiv[k0 + lane] : -1;
    int row_lo = lane < npiv ? (int)(k0 + lane) : -1, row_hi = -1;
    int src_lo = lane, src_hi = lane + 32;
    int nextra = 0;
    for (int t = 0; t < npiv; t++) {
      const int p = __shfl_sync(0xffffffffu, mypiv, t);
      if (p == (int)(k0 + t)) continue;                      // uniform
      int idx;
      if (p < (int)(k0 + npiv)) {
        idx = p - (int)k0;
      } else {
        const unsigned int hit = __ballot_sync(0xffffffffu, row_hi == p);
        if (hit) {
          idx = 32 + __ffs(hit) - 1;
        } else {
          idx = 32 + nextra;
          if (lane == nextra) row_hi = p;
          nextra++;
        }
      }
      const int st = __shfl_sync(0xffffffffu, src_lo, t);
      const int si = idx < 32 ? __shfl_sync(0xffffffffu, src_lo, idx) : __shfl_sync(0xffffffffu, src_hi, idx - 32);
      if (lane == t) src_lo = si;
      if (idx < 32) { if (lane == idx) src_lo = st; }
      else if (lane == idx - 32) src_hi = st;
    }
    s_row[lane] = row_lo; s_row[32 + lane] = row_hi;
    s_src[lane] = row_lo >= 0 ? src_lo : lane;               // unused slots map to themselves: not moved
    s_src[32 + lane] = row_hi >= 0 ? src_hi : 32 + lane;
  }
  __syncthreads();
  if (active) {
    double v[64];
#pragma unroll
    for (int s = 0; s < 64; s++) {
      const int src = s_src[s];
      if (src != s) v[s] = K[(long long)s_row[src] * ld + c];
    }
#pragma unroll
    for (int s = 0; s < 64; s++) {
      if (s_src[s] != s) K[(long long)s_row[s] * ld + c] = v[s];
    }
  }
  }
}

// interchanges on columns [c0, c0+ncols) minus the hole [hole0, hole0+holew) (holew = 0: no hole)
int swap_rows_hole(UpdesLU *h, int v, int64_t c0, int64_t ncols, int64_t hole0, int64_t holew, int64_t k0, int64_t npiv,
                   const int32_t *ipiv, cudaStream_t st) {
  if (holew <= 0) { hole0 = c0 + ncols; holew = 0; }
  if (hole0 < c0 || hole0 + holew > c0 + ncols) return -3;
  const int64_t work = ncols - holew;
  if (work <= 0 || npiv <= 0) return 0;
  double *K = h->view[v].ptr;
  const long long ld = h->view[v].ld;
  if (!K) return -2;
  if (npiv > 0x7fffffff) return -6;
  prof_begin(PROF_SWAP, 32.0 * (double)work * (double)npiv, st);
  swap_rows_kernel<<<(unsigned)((work + SWAP_THREADS - 1) / SWAP_THREADS), SWAP_THREADS, 0, st>>>(
      K, ld, c0, ncols, hole0, holew, k0, (int)npiv, ipiv);
  prof_end(st);
  UPDES_LAUNCH_CHECK();
  return 0;
}

int swap_rows(UpdesLU *h, int v, int64_t c0, int64_t ncols, int64_t k0, int64_t npiv, const int32_t *ipiv,
              cudaStream_t st) {
  return swap_rows_hole(h, v, c0, ncols, 0, 0, k0, npiv, ipiv, st);
}

// ---- unit-lower triangular solve, base case ------------------------------------------------------
// X = L^-1 B with L = K[r0:r0+NB, r0:r0+NB] (unit lower) and B = K[r0:r0+NB, c0:c0+ncols], in place.
// One thread per column of B: the NB values of the column live in registers, L is broadcast from
// shared memory; loads/stores are coalesced across the threads of a warp (adjacent columns).
constexpr int TRSM_THREADS = 64;
template <int NB>
__global__ void __launch_bounds__(TRSM_THREADS) trsm_base_kernel(const double *Lm, long long ldl, long long rl, long long cl,
                                                        double *B, long long ldb, long long rb, long long cb,
                                                        long long ncols) {
  __shared__ double L[NB][NB + 1];
  for (int t = threadIdx.x; t < NB * NB; t += blockDim.x) {
    const int i = t / NB, j = t % NB;
    L[i][j] = Lm[(rl + i) * ldl + cl + j];
  }
  __syncthreads();
  const long long c = cb + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cb + ncols) return;
  double x[NB];
#pragma unroll
  for (int i = 0; i < NB; i++) x[i] = B[(rb + i) * ldb + c];
  // column-oriented elimination: after x[j] is final, the updates of x[j+1..] are independent of each other
  // (the row-oriented form is one dependent FMA chain of length NB(NB-1)/2 per thread)
#pragma unroll
  for (int j = 0; j < NB - 1; j++) {
#pragma unroll
    for (int i = j + 1; i < NB; i++) x[i] = fma(-L[i][j], x[j], x[i]);
  }
#pragma unroll
  for (int i = 1; i < NB; i++) B[(rb + i) * ldb + c] = x[i];
}

// Base case for 32 < NB <= 128 (a multiple of 32): the same thread-per-column substitution, carried out in chunks of 32
// rows.  A chunk first takes the contributions of all earlier (already final) rows, then is solved in registers as
// above.  Replaces, per 128 rows, 4 launches of the 32-row kernel plus 3 tiny GEMM launches (k = 32, 64) of the
// recursive solve, which were launch latency, not work.
//   * L is staged TRANSPOSED in shared memory (Lt[j][i] = L[i][j]): for a fixed solved row j the 32 multipliers of a
//     chunk are contiguous, so one broadcast LDS.128 feeds two FMAs (one LDS.64 per FMA made the first version
//     LSU-issue bound);
//   * the solved rows of a column are re-read from B itself (written by the same thread, coalesced across the warp,
//     L1/L2 hits) instead of a 64 KB shared-memory copy: the CTA needs 130 KB, not 197 KB, and runs up to 256 threads
//     (the first version ran 64 threads = 2 warps per SM: 440 ms of a 90k-node LU);
//   * narrow right-hand sides use smaller CTAs so that the columns spread over all SMs.
constexpr int TRSM_BIG = 128;
constexpr int TRSM_LPT = TRSM_BIG + 2;                   // even (16-byte pairs), and rows 4 banks apart for the staging writes
constexpr int TRSM_BIG_SMEM = TRSM_BIG * TRSM_LPT * (int)sizeof(double);
__global__ void __launch_bounds__(256) trsm_base_big_kernel(const double *__restrict__ Lm, long long ldl, long long rl, long long cl,
                                                            double *B, long long ldb, long long rb, long long cb,
                                                            long long ncols, int nb) {
  extern __shared__ __align__(16) double trsm_sm[];
  double *Lt = trsm_sm;                                  // Lt[j * TRSM_LPT + i] = L[i][j] for j < i < nb
  // staging: 16 independent loads per thread in flight (with 64 threads a plain strided loop is 256 dependent-latency
  // rounds, ~20 us: longer than the substitution itself)
  {
    const int total = nb * nb;
    for (int t0 = threadIdx.x; t0 < total; t0 += 16 * blockDim.x) {
      double v[16];
#pragma unroll
      for (int u = 0; u < 16; u++) {
        const int t = t0 + u * blockDim.x;
        const int i = t / nb, j = t - i * nb;
        v[u] = (t < total && j < i) ? Lm[(rl + i) * ldl + cl + j] : 0.0;
      }
#pragma unroll
      for (int u = 0; u < 16; u++) {
        const int t = t0 + u * blockDim.x;
        const int i = t / nb, j = t - i * nb;
        if (t < total && j < i) Lt[j * TRSM_LPT + i] = v[u];
      }
    }
  }
  __syncthreads();
  const long long c = cb + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cb + ncols) return;
  double *Bc = B + rb * ldb + c;
  for (int q0 = 0; q0 < nb; q0 += 32) {
    double x[32];
#pragma unroll
    for (int i = 0; i < 32; i++) x[i] = Bc[(long long)(q0 + i) * ldb];
    for (int j0 = 0; j0 < q0; j0 += 8) {                 // earlier, final rows: eight re-reads (L2 hits) in flight at a time
      double xs[8];
#pragma unroll
      for (int u = 0; u < 8; u++) xs[u] = -Bc[(long long)(j0 + u) * ldb];
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const double2 *lp = reinterpret_cast<const double2 *>(Lt + (j0 + u) * TRSM_LPT + q0);
#pragma unroll
        for (int i = 0; i < 16; i++) {
          const double2 l = lp[i];
          x[2 * i] = fma(l.x, xs[u], x[2 * i]);
          x[2 * i + 1] = fma(l.y, xs[u], x[2 * i + 1]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 31; j++) {                       // this chunk, column-oriented: the updates of x[j+1..] are independent
      const double *lc = Lt + (q0 + j) * TRSM_LPT + q0;
      const double xj = -x[j];
      if ((j & 1) == 0) x[j + 1] = fma(lc[j + 1], xj, x[j + 1]);
#pragma unroll
      for (int i = (j + 2) & ~1; i < 32; i += 2) {
        const double2 l = *reinterpret_cast<const double2 *>(lc + i);
        x[i] = fma(l.x, xj, x[i]);
        x[i + 1] = fma(l.y, xj, x[i + 1]);
      }
    }
#pragma unroll
    for (int i = 0; i < 32; i++)
      if (q0 + i > 0) Bc[(long long)(q0 + i) * ldb] = x[i];
  }
}

static int trsm_base(UpdesLU *h, int vl, int64_t rl, int64_t cl, int nb, int vb, int64_t rb, int64_t cb,
                     int64_t ncols, cudaStream_t st) {
  const MatView &VL = h->view[vl], &VB = h->view[vb];
  const unsigned grid = (unsigned)((ncols + TRSM_THREADS - 1) / TRSM_THREADS);
  if (nb > 32) {
    if (nb > TRSM_BIG || (nb % 32)) return -4;
    static bool attr = false;
    if (!attr) {
      UPDES_CUDA_TRY(cudaFuncSetAttribute(trsm_base_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TRSM_BIG_SMEM));
      attr = true;
    }
    // one CTA per SM (130 KB of shared memory): the smallest CTA that still covers the columns in one wave
    const int threads = ncols <= 64LL * h->num_sms ? 64 : (ncols <= 128LL * h->num_sms ? 128 : 256);
    const unsigned gridb = (unsigned)((ncols + threads - 1) / threads);
    prof_begin(PROF_TRSM, (double)nb * nb * (double)ncols, st);
    trsm_base_big_kernel<<<gridb, threads, TRSM_BIG_SMEM, st>>>(VL.ptr, VL.ld, rl, cl, VB.ptr, VB.ld, rb, cb, ncols, nb);
    prof_end(st);
    UPDES_LAUNCH_CHECK();
    return 0;
  }
  if (nb != 32 && nb != 16 && nb != 8) return -4;
  prof_begin(PROF_TRSM, (double)nb * nb * (double)ncols, st);
  if (nb == 32) trsm_base_kernel<32><<<grid, TRSM_THREADS, 0, st>>>(VL.ptr, VL.ld, rl, cl, VB.ptr, VB.ld, rb, cb, ncols);
  else if (nb == 16) trsm_base_kernel<16><<<grid, TRSM_THREADS, 0, st>>>(VL.ptr, VL.ld, rl, cl, VB.ptr, VB.ld, rb, cb, ncols);
  else trsm_base_kernel<8><<<grid, TRSM_THREADS, 0, st>>>(VL.ptr, VL.ld, rl, cl, VB.ptr, VB.ld, rb, cb, ncols);
  prof_end(st);
  UPDES_LAUNCH_CHECK();
  return 0;
}

// Recursive blocked solve: split L11, solve the top half, rank-update the bottom half with the
// DMMA GEMM, solve the bottom half.  n1 is a multiple of the base width.
int trsm_unit_lower(UpdesLU *h, int vl, int64_t rl, int64_t cl, int64_t n1, int vb, int64_t rb, int64_t cb,
                    int64_t ncols, cudaStream_t st) {
  if (n1 <= 0 || ncols <= 0) return 0;
  if (!h->view[vl].ptr || !h->view[vb].ptr) return -2;
  if (n1 <= 32 || (h->trsm_base_rows >= n1 && n1 <= TRSM_BIG && (n1 % 32) == 0))
    return trsm_base(h, vl, rl, cl, (int)n1, vb, rb, cb, ncols, st);
  const int64_t hlf = (n1 / 2 + 31) / 32 * 32;
  int rc = trsm_unit_lower(h, vl, rl, cl, hlf, vb, rb, cb, ncols, st);
  if (rc) return rc;
  // B[hlf:, :] -= L[hlf:, :hlf] * B[:hlf, :]
  rc = dgemm_sub(h, vl, rl + hlf, cl, vb, rb, cb, vb, rb + hlf, cb, n1 - hlf, ncols, hlf, st);
  if (rc) return rc;
  return trsm_unit_lower(h, vl, rl + hlf, cl + hlf, n1 - hlf, vb, rb + hlf, cb, ncols, st);
}

// ---- rank-8 update of a narrow block ------------------------------------------------------------------
// C[m x nc] -= A[m x 8] * B[8 x nc], nc <= 8.  Only used inside panels taller than 189 440 rows, whose
// base width drops to 8 columns (the DMMA GEMM needs k >= 16).  One thread per row.
__global__ void __launch_bounds__(256) rank8_update_kernel(double *K, long long ld, long long ra, long long ca,
                                                           long long rb, long long cb, long long rc, long long cc,
                                                           long long m, int nc) {
  __shared__ double B[8][8];
  if (threadIdx.x < 64) {
    const int i = threadIdx.x >> 3, j = threadIdx.x & 7;
    B[i][j] = j < nc ? K[(rb + i) * ld + cb + j] : 0.0;
  }
  __syncthreads();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const double *arow = K + (ra + i) * ld + ca;
  double a[8];
#pragma unroll
  for (int k = 0; k < 8; k++) a[k] = arow[k];
  double *crow = K + (rc + i) * ld + cc;
#pragma unroll
  for (int j = 0; j < 8; j++) {
    if (j < nc) {
      double v = crow[j];
#pragma unroll
      for (int k = 0; k < 8; k++) v = fma(-a[k], B[k][j], v);
      crow[j] = v;
    }
  }
}

int rank8_update(UpdesLU *h, int v, int64_t ra, int64_t ca, int64_t rb, int64_t cb, int64_t rc, int64_t cc, int64_t m,
                 int nc, cudaStream_t st) {
  if (m <= 0 || nc <= 0) return 0;
  if (nc > 8) return -10;
  rank8_update_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(h->view[v].ptr, h->view[v].ld, ra, ca, rb, cb, rc, cc,
                                                                 m, nc);
  UPDES_LAUNCH_CHECK();
  return 0;
}

// ---- row equilibration ------------------------------------------------------------------------------
// The collocation matrix mixes rows of very different scales (operator rows ~ 9 r or 1/DT, boundary rows
// ~ r^3, polynomial rows ~ 1; SURVEY.md section 7, hard part 3).  Partial pivoting compares entries across rows, so
// the rows are first brought to max |entry| in [1, 2) by an exact power-of-two factor (no rounding: the
// scaled matrix is exactly diag(s) K); the solve scales the right-hand side by the same factors.
// One warp per row, 16-byte accesses; HBM-bound (the matrix is read once and read + written once).
__global__ void __launch_bounds__(256) row_absmax_kernel(const double *A, long long rows, long long cols, long long ld,
                                                         double *out) {
  const long long r = blockIdx.x * 8LL + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  const double *row = A + r * ld;
  double m = 0.0;
  const long long c2 = cols & ~1LL;
  for (long long c = 2LL * lane; c < c2; c += 64) {
    const double2 v = *reinterpret_cast<const double2 *>(row + c);
    const double a0 = fabs(v.x), a1 = fabs(v.y);
    m = a0 > m ? a0 : m;                // NaN never raises the maximum
    m = a1 > m ? a1 : m;
  }
  if (lane == 0 && c2 < cols) { const double a0 = fabs(row[c2]); m = a0 > m ? a0 : m; }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const double o = __shfl_xor_sync(0xffffffffu, m, off);
    m = o > m ? o : m;
  }
  if (lane == 0) out[r] = m;
}

// scale[r] = 2^-floor(log2(absmax[r])) (1 for zero / non-finite rows): exponent arithmetic only
__global__ void scale_from_absmax_kernel(const double *absmax, long long n, double *scale) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double m = absmax[i];
  double sc = 1.0;
  if (m > 0.0 && m < 1.7976931348623157e308) {
    int e;
    frexp(m, &e);                        // m = f * 2^e, f in [0.5, 1)  ->  m * 2^(1-e) in [1, 2)
    sc = ldexp(1.0, 1 - e);
  }
  scale[i] = sc;
}

__global__ void __launch_bounds__(256) row_scale_kernel(double *A, long long rows, long long cols, long long ld,
                                                        const double *scale) {
  const long long r = blockIdx.x * 8LL + (threadIdx.x >> 5);
  if (r >= rows) return;
  const double sc = scale[r];
  if (sc == 1.0) return;
  const int lane = threadIdx.x & 31;
  double *row = A + r * ld;
  const long long c2 = cols & ~1LL;
  for (long long c = 2LL * lane; c < c2; c += 64) {
    double2 v = *reinterpret_cast<double2 *>(row + c);
    v.x *= sc; v.y *= sc;
    *reinterpret_cast<double2 *>(row + c) = v;
  }
  if (lane == 0 && c2 < cols) row[c2] *= sc;
}

int row_absmax(const double *A, int64_t rows, int64_t cols, int64_t ld, double *out, cudaStream_t st) {
  if (rows <= 0) return 0;
  row_absmax_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(A, rows, cols, ld, out);
  UPDES_LAUNCH_CHECK();
  return 0;
}
int scale_from_absmax(const double *absmax, int64_t n, double *scale, cudaStream_t st) {
  if (n <= 0) return 0;
  scale_from_absmax_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(absmax, n, scale);
  UPDES_LAUNCH_CHECK();
  return 0;
}
int row_scale(double *A, int64_t rows, int64_t cols, int64_t ld, const double *scale, cudaStream_t st) {
  if (rows <= 0) return 0;
  row_scale_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(A, rows, cols, ld, scale);
  UPDES_LAUNCH_CHECK();
  return 0;
}

// ---- pivots -> permutation ------------------------------------------------------------------------
// perm[i] = index of the original row that the interchanges leave at position i.
__global__ void perm_init_kernel(int32_t *perm, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) perm[i] = i;
}
// One warp walks the pivot list 32 interchanges at a time: the 32 pivots of a chunk arrive in one coalesced load, their
// sequence is composed into a permutation of the <= 64 positions involved with register/shuffle work (the same
// composition as swap_rows_kernel), and the moved entries make ONE gather/scatter round trip.  The literal loop (one
// thread, two dependent global round trips per pivot) cost ~0.7 us per pivot: 6 ms of a 59 ms LU at n = 8192.
__global__ void __launch_bounds__(32) perm_apply_kernel(int32_t *perm, const int32_t *ipiv, int n) {
  const int lane = threadIdx.x;
  int nextpiv = lane < n ? ipiv[lane] : -1;
  for (int k0 = 0; k0 < n; k0 += 32) {
    const int npiv = n - k0 < 32 ? n - k0 : 32;
    const int mypiv = nextpiv;
    nextpiv = (k0 + 32 + lane) < n ? ipiv[k0 + 32 + lane] : -1;          // prefetch the next chunk
    const bool moved = lane < npiv && mypiv != k0 + lane;
    if (!__any_sync(0xffffffffu, moved)) continue;
    int row_lo = lane < npiv ? k0 + lane : -1, row_hi = -1;
    int src_lo = lane, src_hi = lane + 32;
    int nextra = 0;
    for (int t = 0; t < npiv; t++) {
      const int p = __shfl_sync(0xffffffffu, mypiv, t);
      if (p == k0 + t) continue;                                          // uniform
      int idx;
      if (p < k0 + npiv) {
        idx = p - k0;
      } else {
        const unsigned int hit = __ballot_sync(0xffffffffu, row_hi == p);
        if (hit) {
          idx = 32 + __ffs(hit) - 1;
        } else {
          idx = 32 + nextra;
          if (lane == nextra) row_hi = p;
          nextra++;
        }
      }
      const int st = __shfl_sync(0xffffffffu, src_lo, t);
      const int si = idx < 32 ? __shfl_sync(0xffffffffu, src_lo, idx) : __shfl_sync(0xffffffffu, src_hi, idx - 32);
      if (lane == t) src_lo = si;
      if (idx < 32) { if (lane == idx) src_lo = st; }
      else if (lane == idx - 32) src_hi = st;
    }
    // row of the slot whose original content ends up in my two slots
    const int r_lo_a = __shfl_sync(0xffffffffu, row_lo, src_lo & 31), r_lo_b = __shfl_sync(0xffffffffu, row_hi, src_lo & 31);
    const int r_hi_a = __shfl_sync(0xffffffffu, row_lo, src_hi & 31), r_hi_b = __shfl_sync(0xffffffffu, row_hi, src_hi & 31);
    const int from_lo = src_lo < 32 ? r_lo_a : r_lo_b, from_hi = src_hi < 32 ? r_hi_a : r_hi_b;
    const bool mv_lo = row_lo >= 0 && src_lo != lane, mv_hi = row_hi >= 0 && src_hi != lane + 32;
    int v_lo = 0, v_hi = 0;
    if (mv_lo) v_lo = __ldcg(perm + from_lo);
    if (mv_hi) v_hi = __ldcg(perm + from_hi);
    __syncwarp();
    if (mv_lo) __stcg(perm + row_lo, v_lo);
    if (mv_hi) __stcg(perm + row_hi, v_hi);
    __syncwarp();
  }
}

int build_permutation(UpdesLU *h, const int32_t *ipiv, cudaStream_t st) {
  const int n = (int)h->n;
  perm_init_kernel<<<(n + 255) / 256, 256, 0, st>>>(h->perm, n);
  UPDES_LAUNCH_CHECK();
  perm_apply_kernel<<<1, 32, 0, st>>>(h->perm, ipiv, n);
  UPDES_LAUNCH_CHECK();
  return 0;
}

}  // namespace updes

// Base panel of the blocked LU: partial-pivoting factorisation of a tall block of JB columns.
//
// The whole panel lives in REGISTERS for the duration of the kernel: each thread owns RPT rows of
// JB doubles (JB*RPT = 32), so a 90k x 32 panel is spread over 148 CTAs x 640 threads and global
// memory is touched exactly twice (load, store).  Per column j:
//   1. every CTA finds its local arg-max |a(i,j)| with warp shuffles (ties -> smallest row, as
//      LAPACK's idamax) and PUBLISHES the candidate's whole row (JB doubles) next to (|v|, row);
//      CTA 0 also publishes the row currently sitting at the diagonal position;
//   2. ONE grid barrier (monotonic atomic counter, acquire/release);
//   3. every CTA reduces the <=148 candidates, reads the winner's row, performs the row
//      interchange on the registers it owns and applies the rank-1 update to its rows.
// Publishing the candidate rows before the barrier is what keeps it at one barrier per column:
// nobody has to ask the pivot owner for its row afterwards.
// Launched cooperatively (all CTAs co-resident), grid <= number of SMs.
#include "lu.cuh"
#include <cooperative_groups.h>

namespace updes {

constexpr int PANEL_THREADS = 640;

struct PanelParams {
  double *K;
  long long ld, n, r0, c0;   // panel = rows [r0, n) x columns [c0, c0+jb)
  int jb;              // columns in this panel (<= JB)
  int rows_per_cta;    // R
  int32_t *ipiv, *info;
  double *cand, *top, *candval;
  int32_t *candrow;
  unsigned int *barrier;
  unsigned int barrier_base;   // counter value when this launch starts
  int num_ctas;
};

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int *p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// (value, row) arg-max with LAPACK tie-breaking (first = smallest row index)
__device__ __forceinline__ void argmax_combine(double &v, int &r, double ov, int orow) {
  if (ov > v || (ov == v && orow < r)) { v = ov; r = orow; }
}

template <int JB, int RPT>
__global__ void __launch_bounds__(PANEL_THREADS, 1) lu_panel_kernel(PanelParams P) {
  __shared__ double s_wv[32];
  __shared__ int s_wr[32];
  __shared__ double s_prow[JB], s_trow[JB];
  __shared__ int s_piv;

  const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
  const int b = blockIdx.x;
  const int jb = P.jb;
  const long long cta_row0 = P.r0 + (long long)b * P.rows_per_cta;

  // ---- load my rows ------------------------------------------------------------------
  double a[RPT][JB];
  long long myrow[RPT];
#pragma unroll
  for (int q = 0; q < RPT; q++) {
    const int li = tid + q * T;
    const long long row = cta_row0 + li;
    const bool valid = li < P.rows_per_cta && row < P.n;
    myrow[q] = valid ? row : -1;
    if (valid) {
      const double *src = P.K + row * P.ld + P.c0;
      if (jb == JB) {
#pragma unroll
        for (int c = 0; c < JB; c += 2) {
          const double2 v = *reinterpret_cast<const double2 *>(src + c);
          a[q][c] = v.x; a[q][c + 1] = v.y;
        }
      } else {
#pragma unroll
        for (int c = 0; c < JB; c++) a[q][c] = c < jb ? src[c] : 0.0;
      }
    } else {
#pragma unroll
      for (int c = 0; c < JB; c++) a[q][c] = 0.0;
    }
  }

  unsigned int target = P.barrier_base;

  // The column loop is unrolled 8 steps at a time and the register row is rotated left by 8 between
  // groups, so the current column is always one of registers 0..7 (static indexing) while the kernel
  // body stays ~4x smaller than a full 32-step unroll (ncu: instruction-cache misses were the second
  // largest stall).  Every thread applies the same rotation, so published rows are exchanged in
  // rotated form; after JB/8 groups the rotation is the identity again.
#pragma unroll 1
  for (int g = 0; g < JB / 8; g++) {
  const int live = JB - 8 * g;                 // registers [0, live) hold columns not yet eliminated
#pragma unroll
  for (int jl = 0; jl < 8; jl++) {
    const int j = 8 * g + jl;
    if (j < jb) {   // uniform
      const long long diag = P.r0 + j;
      const int par = j & 1;
      // ---- 1. local candidate -----------------------------------------------------------
      double bv = -1.0;
      int br = 0x7fffffff;
#pragma unroll
      for (int q = 0; q < RPT; q++)
        if (myrow[q] >= diag) argmax_combine(bv, br, fabs(a[q][jl]), (int)myrow[q]);
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bv, off);
        const int orow = __shfl_xor_sync(0xffffffffu, br, off);
        argmax_combine(bv, br, ov, orow);
      }
      if (lane == 0) { s_wv[warp] = bv; s_wr[warp] = br; }
      __syncthreads();
      double cv = lane < nwarps ? s_wv[lane] : -1.0;
      int cr = lane < nwarps ? s_wr[lane] : 0x7fffffff;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, cv, off);
        const int orow = __shfl_xor_sync(0xffffffffu, cr, off);
        argmax_combine(cv, cr, ov, orow);
      }
      // ---- publish -------------------------------------------------------------------------
#pragma unroll
      for (int q = 0; q < RPT; q++) {
        if (myrow[q] >= 0 && myrow[q] == (long long)cr) {
          double *dst = P.cand + ((size_t)par * P.num_ctas + b) * JB;
#pragma unroll
          for (int c = 0; c < JB; c++) dst[c] = a[q][c];
        }
        if (myrow[q] == diag) {
          double *dst = P.top + (size_t)par * JB;
#pragma unroll
          for (int c = 0; c < JB; c++) dst[c] = a[q][c];
        }
      }
      if (tid == 0) {
        P.candval[par * P.num_ctas + b] = cv;
        P.candrow[par * P.num_ctas + b] = cr;
      }
      __threadfence();
      __syncthreads();
      // ---- 2. grid barrier -------------------------------------------------------------------
      target += (unsigned int)P.num_ctas;
      if (tid == 0) {
        atomicAdd(P.barrier, 1u);
        // bounded spin: a lost arrival must not hang the device (info = -1 flags the failure)
        unsigned int polls = 0;
        while ((int)(ld_acquire_u32(P.barrier) - target) < 0) {
          if (++polls > (1u << 24)) { atomicExch(P.info, -1); break; }
        }
        __threadfence();
      }
      __syncthreads();
      // ---- 3. winner -----------------------------------------------------------------------------
      if (warp == 0) {
        double wv = -1.0;
        int wr = 0x7fffffff, wb = 0;
        for (int c = lane; c < P.num_ctas; c += 32) {
          const double ov = __ldcg(P.candval + par * P.num_ctas + c);
          const int orow = __ldcg(P.candrow + par * P.num_ctas + c);
          if (ov > wv || (ov == wv && orow < wr)) { wv = ov; wr = orow; wb = c; }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
          const double ov = __shfl_xor_sync(0xffffffffu, wv, off);
          const int orow = __shfl_xor_sync(0xffffffffu, wr, off);
          const int ob = __shfl_xor_sync(0xffffffffu, wb, off);
          if (ov > wv || (ov == wv && orow < wr)) { wv = ov; wr = orow; wb = ob; }
        }
        const bool none = (wr == 0x7fffffff);   // no comparable candidate (NaN column): keep the diagonal row
        if (none) { wr = (int)diag; wv = 0.0; }
        if (lane < JB) {
          const double t = __ldcg(P.top + (size_t)par * JB + lane);
          s_trow[lane] = t;
          s_prow[lane] = none ? t : __ldcg(P.cand + ((size_t)par * P.num_ctas + wb) * JB + lane);
        }
        if (lane == 0) {
          s_piv = wr;
          if (b == 0) {
            P.ipiv[diag] = wr;
            if (wv == 0.0) atomicCAS(P.info, 0, (int)diag + 1);
          }
        }
      }
      __syncthreads();
      // ---- 4. interchange + rank-1 update ------------------------------------------------------------
      const long long piv = s_piv;
      const double pval = s_prow[jl];
      const double rinv = pval != 0.0 ? 1.0 / pval : 0.0;
#pragma unroll
      for (int q = 0; q < RPT; q++) {
        if (myrow[q] < diag) continue;
        if (myrow[q] == diag) {
#pragma unroll
          for (int c = 0; c < JB; c++) a[q][c] = s_prow[c];
          continue;
        }
        if (myrow[q] == piv) {
#pragma unroll
          for (int c = 0; c < JB; c++) a[q][c] = s_trow[c];
        }
        if (pval != 0.0) {
          const double l = a[q][jl] * rinv;
          a[q][jl] = l;
#pragma unroll
          for (int c = jl + 1; c < JB; c++)
            if (c < live) a[q][c] = fma(-l, s_prow[c], a[q][c]);
        }
      }
    }
  }
  // rotate the register rows left by 8 columns
#pragma unroll
  for (int q = 0; q < RPT; q++) {
    double t8[8];
#pragma unroll
    for (int c = 0; c < 8; c++) t8[c] = a[q][c];
#pragma unroll
    for (int c = 0; c < JB - 8; c++) a[q][c] = a[q][c + 8];
#pragma unroll
    for (int c = 0; c < 8; c++) a[q][JB - 8 + c] = t8[c];
  }
  }

  // ---- store ---------------------------------------------------------------------------
#pragma unroll
  for (int q = 0; q < RPT; q++) {
    if (myrow[q] < 0) continue;
    double *dst = P.K + myrow[q] * P.ld + P.c0;
    if (jb == JB) {
#pragma unroll
      for (int c = 0; c < JB; c += 2) *reinterpret_cast<double2 *>(dst + c) = make_double2(a[q][c], a[q][c + 1]);
    } else {
#pragma unroll
      for (int c = 0; c < JB; c++)
        if (c < jb) dst[c] = a[q][c];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Cluster variant for short panels (m <= 16 x 512 rows): the CTAs of ONE thread-block cluster hold the
// panel, candidates are exchanged through distributed shared memory and the per-column barrier is a
// hardware cluster barrier (~0.6 us per column instead of ~5 us through global memory).  Same
// algorithm and tie-breaking as lu_panel_kernel; two candidate buffers alternate by column parity, so
// one cluster barrier per column is enough (a CTA can only overwrite buffer p two columns later, after
// every CTA has passed the barrier in between, i.e. finished reading it).
// ------------------------------------------------------------------------------------------------
namespace cg = cooperative_groups;

constexpr int CLUSTER_PANEL_THREADS = 512;   // 128 registers per thread: no spills with the DSMEM pointers live

template <int JB>
__global__ void __launch_bounds__(CLUSTER_PANEL_THREADS, 1) lu_panel_cluster_kernel(PanelParams P) {
  __shared__ double s_wv[32];
  __shared__ int s_wr[32];
  __shared__ double s_prow[JB], s_trow[JB];
  __shared__ int s_piv;
  // exchanged through DSMEM (double-buffered by column parity)
  __shared__ double x_cand[2][JB], x_top[2][JB], x_cval[2];
  __shared__ int x_crow[2];

  cg::cluster_group cluster = cg::this_cluster();
  const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
  const int b = (int)cluster.block_rank(), C = (int)cluster.num_blocks();
  const int jb = P.jb;
  const long long row = P.r0 + (long long)b * P.rows_per_cta + tid;
  const bool valid = tid < P.rows_per_cta && row < P.n;
  const long long myrow = valid ? row : -1;

  double a[JB];
  if (valid) {
    const double *src = P.K + row * P.ld + P.c0;
    if (jb == JB) {
#pragma unroll
      for (int c = 0; c < JB; c += 2) {
        const double2 v = *reinterpret_cast<const double2 *>(src + c);
        a[c] = v.x; a[c + 1] = v.y;
      }
    } else {
#pragma unroll
      for (int c = 0; c < JB; c++) a[c] = c < jb ? src[c] : 0.0;
    }
  } else {
#pragma unroll
    for (int c = 0; c < JB; c++) a[c] = 0.0;
  }

#pragma unroll 1
  for (int g = 0; g < JB / 8; g++) {           // 8 unrolled column steps per group, rows rotated by 8 in between
  const int live = JB - 8 * g;
#pragma unroll
  for (int jl = 0; jl < 8; jl++) {
    const int j = 8 * g + jl;
    if (j < jb) {   // uniform
      const long long diag = P.r0 + j;
      const int par = j & 1;
      double bv = -1.0;
      int br = 0x7fffffff;
      if (myrow >= diag) { bv = fabs(a[jl]); br = (int)myrow; }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bv, off);
        const int orow = __shfl_xor_sync(0xffffffffu, br, off);
        argmax_combine(bv, br, ov, orow);
      }
      if (lane == 0) { s_wv[warp] = bv; s_wr[warp] = br; }
      __syncthreads();
      double cv = lane < nwarps ? s_wv[lane] : -1.0;
      int cr = lane < nwarps ? s_wr[lane] : 0x7fffffff;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, cv, off);
        const int orow = __shfl_xor_sync(0xffffffffu, cr, off);
        argmax_combine(cv, cr, ov, orow);
      }
      if (myrow >= 0 && myrow == (long long)cr) {
#pragma unroll
        for (int c = 0; c < JB; c++) x_cand[par][c] = a[c];
      }
      if (myrow == diag) {
#pragma unroll
        for (int c = 0; c < JB; c++) x_top[par][c] = a[c];
      }
      if (tid == 0) { x_cval[par] = cv; x_crow[par] = cr; }
      cluster.sync();                                  // release/acquire across the cluster
      if (warp == 0) {
        double wv = -1.0;
        int wr = 0x7fffffff, wb = 0;
        for (int c = lane; c < C; c += 32) {
          const double ov = *cluster.map_shared_rank(&x_cval[par], c);
          const int orow = *cluster.map_shared_rank(&x_crow[par], c);
          if (ov > wv || (ov == wv && orow < wr)) { wv = ov; wr = orow; wb = c; }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
          const double ov = __shfl_xor_sync(0xffffffffu, wv, off);
          const int orow = __shfl_xor_sync(0xffffffffu, wr, off);
          const int ob = __shfl_xor_sync(0xffffffffu, wb, off);
          if (ov > wv || (ov == wv && orow < wr)) { wv = ov; wr = orow; wb = ob; }
        }
        const bool none = (wr == 0x7fffffff);
        if (none) { wr = (int)diag; wv = 0.0; }
        if (lane < JB) {
          const double t = cluster.map_shared_rank(&x_top[par][0], 0)[lane];     // diagonal rows live in CTA 0
          s_trow[lane] = t;
          s_prow[lane] = none ? t : cluster.map_shared_rank(&x_cand[par][0], wb)[lane];
        }
        if (lane == 0) {
          s_piv = wr;
          if (b == 0) {
            P.ipiv[diag] = wr;
            if (wv == 0.0) atomicCAS(P.info, 0, (int)diag + 1);
          }
        }
      }
      __syncthreads();
      const long long piv = s_piv;
      const double pval = s_prow[jl];
      const double rinv = pval != 0.0 ? 1.0 / pval : 0.0;
      if (myrow == diag) {
#pragma unroll
        for (int c = 0; c < JB; c++) a[c] = s_prow[c];
      } else if (myrow > diag) {
        if (myrow == piv) {
#pragma unroll
          for (int c = 0; c < JB; c++) a[c] = s_trow[c];
        }
        if (pval != 0.0) {
          const double l = a[jl] * rinv;
          a[jl] = l;
#pragma unroll
          for (int c = jl + 1; c < JB; c++)
            if (c < live) a[c] = fma(-l, s_prow[c], a[c]);
        }
      }
    }
  }
  {
    double t8[8];
#pragma unroll
    for (int c = 0; c < 8; c++) t8[c] = a[c];
#pragma unroll
    for (int c = 0; c < JB - 8; c++) a[c] = a[c + 8];
#pragma unroll
    for (int c = 0; c < 8; c++) a[JB - 8 + c] = t8[c];
  }
  }
  if (valid) {
    double *dst = P.K + row * P.ld + P.c0;
    if (jb == JB) {
#pragma unroll
      for (int c = 0; c < JB; c += 2) *reinterpret_cast<double2 *>(dst + c) = make_double2(a[c], a[c + 1]);
    } else {
#pragma unroll
      for (int c = 0; c < JB; c++)
        if (c < jb) dst[c] = a[c];
    }
  }
  cluster.sync();   // nobody may exit while a peer can still read its shared memory
}

// ------------------------------------------------------------------------------------------------
// Second-generation panel kernels (panel_variant 2, default).  Same arithmetic and the same pivots as
// the kernels above; what changes is the per-column latency chain:
//   * implicit pivoting: rows never move during the panel.  Each row carries its current logical
//     position; an interchange is two threads updating an integer, and the rows are written to their
//     final positions once, at the end.  (The diagonal row no longer has to be published per column.)
//   * arg-max on the integer pipe: |a| is compared as a 64-bit pattern with three REDUX instructions
//     (max of the high words, max of the low words among those, min position among the maxima: LAPACK's
//     first-index tie-break) instead of five shuffle rounds on (value, row) pairs, and every warp reduces
//     the per-warp candidates redundantly, so there is one block barrier per column, not two;
//   * cluster kernel: every CTA PUSHES its candidate (key, position, the whole row) into the shared
//     memory of all CTAs of the cluster, then one split cluster barrier (arrive.release / wait.acquire);
//     after the barrier everything a CTA needs is in its own shared memory -- no dependent remote reads;
//   * grid kernel: the thread that owns the CTA's candidate publishes it and arrives at the grid barrier
//     itself (no block barrier between publish and arrive).
// ------------------------------------------------------------------------------------------------
constexpr unsigned int FULL = 0xffffffffu;
constexpr int NOPOS = 0x7fffffff;

__device__ __forceinline__ unsigned long long abs_key(double v) {
  const double av = fabs(v);
  return av == av ? (unsigned long long)__double_as_longlong(av) : 0ull;    // NaN never wins a comparison
}

// warp-wide arg-max of (key, pos): largest key, ties -> smallest pos.  All lanes get the result.
__device__ __forceinline__ void warp_argmax(unsigned long long &key, int &pos) {
  const unsigned int hi = (unsigned int)(key >> 32), lo = (unsigned int)key;
  const unsigned int mhi = __reduce_max_sync(FULL, hi);
  const unsigned int mlo = __reduce_max_sync(FULL, hi == mhi ? lo : 0u);
  const bool match = hi == mhi && lo == mlo;
  const unsigned int mp = __reduce_min_sync(FULL, match ? (unsigned int)pos : (unsigned int)NOPOS);
  key = ((unsigned long long)mhi << 32) | mlo;
  pos = (int)mp;
}

__device__ __forceinline__ void cluster_arrive_release() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait_acquire() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

template <int JB>
__global__ void __launch_bounds__(CLUSTER_PANEL_THREADS, 1) lu_panel_cluster2_kernel(PanelParams P) {
  constexpr int MAXC = 16;
  __shared__ unsigned long long s_key[2][32];
  __shared__ int s_pos[2][32];
  __shared__ __align__(16) double s_stage[2][JB];
  __shared__ __align__(16) double x_cand[2][MAXC][JB];          // pushed by the CTAs of the cluster
  __shared__ __align__(16) unsigned long long x_kp[2][MAXC][2]; // (key, position) of each CTA's candidate

  cg::cluster_group cluster = cg::this_cluster();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int b = (int)cluster.block_rank(), C = (int)cluster.num_blocks();
  const int jb = P.jb;
  const long long row = P.r0 + (long long)b * P.rows_per_cta + tid;
  const bool valid = tid < P.rows_per_cta && row < P.n;
  int mypos = valid ? (int)row : NOPOS;       // current logical position of my row
  bool active = valid;                        // not yet chosen as a pivot row

  double a[JB];
  if (valid) {
    const double *src = P.K + row * P.ld + P.c0;
    if (jb == JB) {
#pragma unroll
      for (int c = 0; c < JB; c += 2) {
        const double2 v = *reinterpret_cast<const double2 *>(src + c);
        a[c] = v.x; a[c + 1] = v.y;
      }
    } else {
#pragma unroll
      for (int c = 0; c < JB; c++) a[c] = c < jb ? src[c] : 0.0;
    }
  } else {
#pragma unroll
    for (int c = 0; c < JB; c++) a[c] = 0.0;
  }

#pragma unroll 1
  for (int g = 0; g < JB / 8; g++) {           // 8 unrolled column steps per group, rows rotated by 8 in between
  const int live = JB - 8 * g;
#pragma unroll
  for (int jl = 0; jl < 8; jl++) {
    const int j = 8 * g + jl;
    if (j < jb) {   // uniform
      const int diag = (int)P.r0 + j;
      const int par = j & 1;
      // ---- candidate of this CTA (every warp ends up knowing it) ------------------------------------
      unsigned long long key = active ? abs_key(a[jl]) : 0ull;
      int pos = active ? mypos : NOPOS;
      warp_argmax(key, pos);
      if (lane == 0) { s_key[par][warp] = key; s_pos[par][warp] = pos; }
      __syncthreads();
      unsigned long long ckey = lane < nwarps ? s_key[par][lane] : 0ull;
      int cpos = lane < nwarps ? s_pos[par][lane] : NOPOS;
      warp_argmax(ckey, cpos);
      // ---- push it to every CTA of the cluster -----------------------------------------------------------
      const unsigned int own = __ballot_sync(FULL, active && mypos == cpos);
      if (own) {                                // this warp holds the candidate row
        if (lane == __ffs(own) - 1) {
#pragma unroll
          for (int c = 0; c < JB; c += 2) *reinterpret_cast<double2 *>(&s_stage[par][c]) = make_double2(a[c], a[c + 1]);
        }
        __syncwarp();
        constexpr int CH = JB / 2;              // 16-byte chunks per row
        const int chunk = lane % CH;
        const double2 v = *reinterpret_cast<const double2 *>(&s_stage[par][2 * chunk]);
        for (int d = lane / CH; d < C; d += 32 / CH) {
          double *remote = cluster.map_shared_rank(&x_cand[par][b][0], d);
          *reinterpret_cast<double2 *>(remote + 2 * chunk) = v;
        }
      }
      if (warp == 0 && lane < C) {
        unsigned long long *remote = cluster.map_shared_rank(&x_kp[par][b][0], lane);
        *reinterpret_cast<ulonglong2 *>(remote) = make_ulonglong2(ckey, (unsigned long long)(unsigned int)cpos);
      }
      cluster_arrive_release();
      cluster_wait_acquire();
      // ---- winner: everything is in local shared memory now ------------------------------------------------
      unsigned long long wkey = 0ull;
      int wpos = NOPOS;
      if (lane < C) {
        const ulonglong2 kp = *reinterpret_cast<const ulonglong2 *>(&x_kp[par][lane][0]);
        wkey = kp.x; wpos = (int)(unsigned int)kp.y;
      }
      const int lpos = wpos;
      warp_argmax(wkey, wpos);
      const int wb = (int)__reduce_min_sync(FULL, (lane < C && lpos == wpos) ? (unsigned int)lane : 99u);
      const double *prow = &x_cand[par][wb][0];
      if (b == 0 && tid == 0) {
        P.ipiv[diag] = wpos;
        if (wkey == 0ull) atomicCAS(P.info, 0, diag + 1);
      }
      // ---- interchange (positions only) + rank-1 update -------------------------------------------------------
      const double pval = prow[jl];
      const double rinv = pval != 0.0 ? 1.0 / pval : 0.0;
      if (active) {
        if (mypos == wpos) {
          active = false; mypos = diag;                      // pivot row: frozen, will land on the diagonal
        } else {
          if (mypos == diag) mypos = wpos;                   // displaced by the interchange
          if (pval != 0.0) {
            const double l = a[jl] * rinv;
            a[jl] = l;
#pragma unroll
            for (int c = jl + 1; c < JB; c++)
              if (c < live) a[c] = fma(-l, prow[c], a[c]);
          }
        }
      }
    }
  }
  {
    double t8[8];
#pragma unroll
    for (int c = 0; c < 8; c++) t8[c] = a[c];
#pragma unroll
    for (int c = 0; c < JB - 8; c++) a[c] = a[c + 8];
#pragma unroll
    for (int c = 0; c < 8; c++) a[JB - 8 + c] = t8[c];
  }
  }
  if (valid) {
    double *dst = P.K + (long long)mypos * P.ld + P.c0;
    if (jb == JB) {
#pragma unroll
      for (int c = 0; c < JB; c += 2) *reinterpret_cast<double2 *>(dst + c) = make_double2(a[c], a[c + 1]);
    } else {
#pragma unroll
      for (int c = 0; c < JB; c++)
        if (c < jb) dst[c] = a[c];
    }
  }
  // no trailing cluster barrier needed: a CTA leaves the last column's barrier only after every CTA has
  // arrived there, i.e. after the last remote store into its shared memory was issued and released
}

// Grid-wide variant: candidates through global memory, one grid barrier per column.
template <int JB, int RPT>
__global__ void __launch_bounds__(PANEL_THREADS, 1) lu_panel2_kernel(PanelParams P) {
  __shared__ unsigned long long s_key[2][32];
  __shared__ int s_pos[2][32];
  __shared__ __align__(16) double s_prow[JB];
  __shared__ int s_piv;

  const int tid = threadIdx.x, T = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarps = T >> 5;
  const int b = blockIdx.x;
  const int jb = P.jb;
  const long long cta_row0 = P.r0 + (long long)b * P.rows_per_cta;

  double a[RPT][JB];
  int mypos[RPT];
  bool active[RPT];
#pragma unroll
  for (int q = 0; q < RPT; q++) {
    const int li = tid + q * T;
    const long long row = cta_row0 + li;
    const bool valid = li < P.rows_per_cta && row < P.n;
    mypos[q] = valid ? (int)row : NOPOS;
    active[q] = valid;
    if (valid) {
      const double *src = P.K + row * P.ld + P.c0;
      if (jb == JB) {
#pragma unroll
        for (int c = 0; c < JB; c += 2) {
          const double2 v = *reinterpret_cast<const double2 *>(src + c);
          a[q][c] = v.x; a[q][c + 1] = v.y;
        }
      } else {
#pragma unroll
        for (int c = 0; c < JB; c++) a[q][c] = c < jb ? src[c] : 0.0;
      }
    } else {
#pragma unroll
      for (int c = 0; c < JB; c++) a[q][c] = 0.0;
    }
  }
  const bool valid0 = mypos[0] != NOPOS;      // RPT > 1: validity per slot is re-derived at the store

  unsigned int target = P.barrier_base;
  unsigned long long *candkey = reinterpret_cast<unsigned long long *>(P.candval);

#pragma unroll 1
  for (int g = 0; g < JB / 8; g++) {
  const int live = JB - 8 * g;
#pragma unroll
  for (int jl = 0; jl < 8; jl++) {
    const int j = 8 * g + jl;
    if (j < jb) {   // uniform
      const int diag = (int)P.r0 + j;
      const int par = j & 1;
      // ---- candidate of this CTA ------------------------------------------------------------------
      unsigned long long key = 0ull;
      int pos = NOPOS;
#pragma unroll
      for (int q = 0; q < RPT; q++) {
        if (active[q]) {
          const unsigned long long k = abs_key(a[q][jl]);
          if (k > key || (k == key && mypos[q] < pos)) { key = k; pos = mypos[q]; }
        }
      }
      warp_argmax(key, pos);
      if (lane == 0) { s_key[par][warp] = key; s_pos[par][warp] = pos; }
      __syncthreads();
      unsigned long long ckey = lane < nwarps ? s_key[par][lane] : 0ull;
      int cpos = lane < nwarps ? s_pos[par][lane] : NOPOS;
      warp_argmax(ckey, cpos);
      // ---- publish + arrive: done by the one thread that owns the candidate row ---------------------------
      target += (unsigned int)P.num_ctas;
      bool arrive = (cpos == NOPOS) && tid == 0;             // a CTA without active rows still has to arrive
#pragma unroll
      for (int q = 0; q < RPT; q++) {
        if (active[q] && mypos[q] == cpos) {
          double *dst = P.cand + ((size_t)par * P.num_ctas + b) * JB;
#pragma unroll
          for (int c = 0; c < JB; c += 2) __stcg(reinterpret_cast<double2 *>(dst + c), make_double2(a[q][c], a[q][c + 1]));
          arrive = true;
        }
      }
      if (arrive) {
        __stcg(candkey + par * P.num_ctas + b, ckey);
        __stcg(P.candrow + par * P.num_ctas + b, cpos);
        __threadfence();
        atomicAdd(P.barrier, 1u);
      }
      if (tid == 0) {
        // bounded spin: a lost arrival must not hang the device (info = -1 flags the failure)
        unsigned int polls = 0;
        while ((int)(ld_acquire_u32(P.barrier) - target) < 0) {
          if (++polls > (1u << 24)) { atomicExch(P.info, -1); break; }
        }
      }
      __syncthreads();
      // ---- winner ------------------------------------------------------------------------------------
      if (warp == 0) {
        unsigned long long wkey = 0ull;
        int wpos = NOPOS, wb = 0;
        constexpr int PER = 5;                                 // up to 160 CTAs
        unsigned long long k5[PER];
        int p5[PER];
#pragma unroll
        for (int u = 0; u < PER; u++) {
          const int c = lane + 32 * u;
          const bool in = c < P.num_ctas;
          k5[u] = in ? __ldcg(candkey + par * P.num_ctas + c) : 0ull;
          p5[u] = in ? __ldcg(P.candrow + par * P.num_ctas + c) : NOPOS;
        }
#pragma unroll
        for (int u = 0; u < PER; u++)
          if (k5[u] > wkey || (k5[u] == wkey && p5[u] < wpos)) { wkey = k5[u]; wpos = p5[u]; wb = lane + 32 * u; }
        const int lpos = wpos;
        warp_argmax(wkey, wpos);
        wb = (int)__reduce_min_sync(FULL, (lpos == wpos && wpos != NOPOS) ? (unsigned int)wb : 0xffffu);
        if (wb == 0xffff) wb = 0;
        if (lane < JB) s_prow[lane] = __ldcg(P.cand + ((size_t)par * P.num_ctas + wb) * JB + lane);
        if (lane == 0) {
          s_piv = wpos;
          if (b == 0) {
            P.ipiv[diag] = wpos;
            if (wkey == 0ull) atomicCAS(P.info, 0, diag + 1);
          }
        }
      }
      __syncthreads();
      // ---- interchange (positions only) + rank-1 update ------------------------------------------------------------
      const int wpos = s_piv;
      const double pval = s_prow[jl];
      const double rinv = pval != 0.0 ? 1.0 / pval : 0.0;
#pragma unroll
      for (int q = 0; q < RPT; q++) {
        if (!active[q]) continue;
        if (mypos[q] == wpos) { active[q] = false; mypos[q] = diag; continue; }
        if (mypos[q] == diag) mypos[q] = wpos;
        if (pval != 0.0) {
          const double l = a[q][jl] * rinv;
          a[q][jl] = l;
#pragma unroll
          for (int c = jl + 1; c < JB; c++)
            if (c < live) a[q][c] = fma(-l, s_prow[c], a[q][c]);
        }
      }
    }
  }
  // rotate the register rows left by 8 columns
#pragma unroll
  for (int q = 0; q < RPT; q++) {
    double t8[8];
#pragma unroll
    for (int c = 0; c < 8; c++) t8[c] = a[q][c];
#pragma unroll
    for (int c = 0; c < JB - 8; c++) a[q][c] = a[q][c + 8];
#pragma unroll
    for (int c = 0; c < 8; c++) a[q][JB - 8 + c] = t8[c];
  }
  }
  (void)valid0;
#pragma unroll
  for (int q = 0; q < RPT; q++) {
    if (mypos[q] == NOPOS) continue;
    double *dst = P.K + (long long)mypos[q] * P.ld + P.c0;
    if (jb == JB) {
#pragma unroll
      for (int c = 0; c < JB; c += 2) *reinterpret_cast<double2 *>(dst + c) = make_double2(a[q][c], a[q][c + 1]);
    } else {
#pragma unroll
      for (int c = 0; c < JB; c++)
        if (c < jb) dst[c] = a[q][c];
    }
  }
}

// largest cluster size (<= 16) this device can co-schedule for the cluster panel kernel; 0 = unsupported
typedef void (*PanelKernelFn)(PanelParams);

static int cluster_panel_max_for(PanelKernelFn fn, int &cached) {
  if (cached >= 0) return cached;
  cached = 0;
  if (cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
    cudaGetLastError();
  }
  for (int c = 16; c >= 1; c >>= 1) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(c); cfg.blockDim = dim3(CLUSTER_PANEL_THREADS); cfg.dynamicSmemBytes = 0;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = c; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int nclusters = 0;
    if (cudaOccupancyMaxActiveClusters(&nclusters, fn, &cfg) == cudaSuccess && nclusters >= 1) {
      cached = c;
      break;
    }
    cudaGetLastError();
  }
  return cached;
}

static int cluster_panel_max(UpdesLU *h) {
  static int cached1 = -1, cached2 = -1;
  return h->panel_variant == 2 ? cluster_panel_max_for(lu_panel_cluster2_kernel<32>, cached2)
                               : cluster_panel_max_for(lu_panel_cluster_kernel<32>, cached1);
}

static int launch_panel_cluster(UpdesLU *h, PanelParams &P, int ctas, int threads, cudaStream_t st) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = 0; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = ctas; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  prof_begin(PROF_PANEL, (double)(P.n - P.r0) * P.jb * P.jb, st);
  cudaError_t e = h->panel_variant == 2 ? cudaLaunchKernelEx(&cfg, lu_panel_cluster2_kernel<32>, P)
                                        : cudaLaunchKernelEx(&cfg, lu_panel_cluster_kernel<32>, P);
  prof_end(st);
  if (e != cudaSuccess) return (int)e;
  ++g_launch_count;
  return 0;
}

template <int JB, int RPT>
static int launch_panel(UpdesLU *h, PanelParams &P, int threads, cudaStream_t st) {
  void *args[] = {&P};
  prof_begin(PROF_PANEL, (double)(P.n - P.r0) * P.jb * P.jb, st);
  void *fn = h->panel_variant == 2 ? (void *)lu_panel2_kernel<JB, RPT> : (void *)lu_panel_kernel<JB, RPT>;
  cudaError_t le = cudaLaunchCooperativeKernel(fn, dim3(P.num_ctas), dim3(threads), args, 0, st);
  prof_end(st);
  if (le != cudaSuccess) return (int)le;
  ++g_launch_count;
  return 0;
}

// widest base panel the register-resident kernel can take for a panel of m rows
int panel_width_for(const UpdesLU *h, int64_t m) {
  const int64_t cap = h->panel_cap > 0 ? h->panel_cap : (int64_t)h->num_sms * PANEL_THREADS;
  if (m <= cap) return 32;
  if (m <= 2 * cap) return 16;
  if (m <= 4 * cap) return 8;
  return 0;
}

int lu_panel_base(UpdesLU *h, int v, int64_t r0, int64_t c0, int jb, int32_t *ipiv, int32_t *info, cudaStream_t st) {
  const MatView &V = h->view[v];
  if (!V.ptr) return -2;
  if (c0 & 1) return -4;
  const int64_t m = V.rows - r0;
  if (m <= 0 || jb <= 0) return 0;
  const int JBmax = panel_width_for(h, m);
  if (JBmax == 0) return -3;
  if (jb > JBmax) return -4;
  // short panels: one thread-block cluster, DSMEM exchange, hardware cluster barrier
  if (h->panel_variant >= 1 && JBmax == 32) {
    const int cmax = cluster_panel_max(h);
    if (cmax > 0 && m <= (int64_t)cmax * CLUSTER_PANEL_THREADS) {
      int c = 1;
      while ((int64_t)c * CLUSTER_PANEL_THREADS < m) c <<= 1;
      int64_t Rc = (m + c - 1) / c;
      Rc = (Rc + 31) / 32 * 32;
      if (Rc < 32) Rc = 32;
      PanelParams Pc;
      Pc.K = V.ptr; Pc.ld = V.ld; Pc.n = V.rows; Pc.r0 = r0; Pc.c0 = c0; Pc.jb = jb; Pc.rows_per_cta = (int)Rc;
      Pc.ipiv = ipiv; Pc.info = info; Pc.cand = nullptr; Pc.top = nullptr; Pc.candval = nullptr; Pc.candrow = nullptr;
      Pc.barrier = nullptr; Pc.barrier_base = 0; Pc.num_ctas = c;
      return launch_panel_cluster(h, Pc, c, (int)Rc, st);
    }
  }
  const int rpt = 32 / JBmax;
  // rows per CTA: enough CTAs to fill the machine, at least 64 rows each, a multiple of 32
  int ctas = (int)((m + 63) / 64);
  if (ctas > h->num_sms) ctas = h->num_sms;
  int64_t R = (m + ctas - 1) / ctas;
  R = (R + 31) / 32 * 32;
  ctas = (int)((m + R - 1) / R);
  int threads = (int)((R + rpt - 1) / rpt);
  threads = (threads + 31) / 32 * 32;
  if (threads > PANEL_THREADS) return -3;
  if (R < jb && ctas > 1) return -3;
  PanelParams P;
  P.K = V.ptr; P.ld = V.ld; P.n = V.rows; P.r0 = r0; P.c0 = c0; P.jb = jb; P.rows_per_cta = (int)R;
  P.ipiv = ipiv; P.info = info; P.cand = h->cand; P.top = h->top; P.candval = h->candval; P.candrow = h->candrow;
  P.barrier = h->barrier; P.barrier_base = h->barrier_count; P.num_ctas = ctas;
  h->barrier_count += (unsigned int)ctas * (unsigned int)jb;
  if (JBmax == 32) return launch_panel<32, 1>(h, P, threads, st);
  if (JBmax == 16) return launch_panel<16, 2>(h, P, threads, st);
  return launch_panel<8, 4>(h, P, threads, st);
}

}  // namespace updes

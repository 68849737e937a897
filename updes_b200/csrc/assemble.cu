// Dense assembly of the collocation system K (replaces updes/assembly.py:10-362).
//
//   K = [[op(Phi) op(P)],      rows 0..Ni-1     internal operator rows   assembly.py:93-137
//        [bd(Phi) bd(P)],      rows Ni..N-1     boundary rows            assembly.py:141-362
//        [  P^T     0   ]]     rows N..N+M-1    polynomial constraints   assembly.py:39-59,:80-83
//
// Every entry is an independent closed form, so the kernel is a pure HBM-write stream:
// algorithmic bytes = 8 per entry.  One CTA owns a TR x 512 or 1024 tile: each thread keeps
// the coordinates of its two or four adjacent centres in registers, row descriptors (evaluation point,
// pre-combined coefficients, optional periodic partner) are staged in shared memory once per
// tile, and every row is written with 16-byte streaming stores (one warp = 512 B or 1 KB contiguous).
#include "common.cuh"

namespace updes {

long long g_launch_count = 0;

constexpr int ASM_THREADS = 256;
// adjacent columns per thread: 2 for the closed-form Laplacian rows (HBM-bound: 12.5 ms vs 14.2 ms with 4 at
// n = 90 003), 4 for general jets (FP64/LSU-bound: fewer row-descriptor loads per entry, 18.3 ms vs 20.5 ms)
constexpr int ASM_TR = 32;               // rows per tile

struct alignas(16) RowStage {
  RowPoint a, b;
  int has_b;
  int skip;
  int pad[2];
};

struct AsmParams {
  const double *centres;
  UpdesRows rows;
  long long row0, nrows;   // collocation rows handled (all < N)
  long long col0, ncols;   // RBF columns handled (all < N)
  double *out;             // out[(r-row0)*ld + (c-col0)]
  long long ld;
  int ip;
  double e2;
};

// stage one evaluation point of a row: coordinates + coefficients with the kernel's constants folded in
template <int MASK>
__device__ __forceinline__ void load_point(RowPoint &rp, const double *pts, int p, const double *c, double gamma,
                                           double eta, double lapfac) {
  rp.x = pts[2 * (size_t)p];
  rp.y = pts[2 * (size_t)p + 1];
  rp.c0 = c[0];
  if (MASK & JET_ISO) {
    rp.cg1 = 0.0; rp.cg2 = 0.0; rp.cg34 = 0.0; rp.ch4 = 0.0;
    rp.ch3 = lapfac * c[3];
  } else {
    rp.cg1 = gamma * c[1]; rp.cg2 = gamma * c[2]; rp.cg34 = gamma * (c[3] + c[4]);
    rp.ch3 = eta * c[3]; rp.ch4 = eta * c[4];
  }
}

template <int KIND, int MASK, int ASM_CPT, int PFIX>
__global__ void __launch_bounds__(ASM_THREADS) assemble_phi_kernel(AsmParams P) {
  constexpr int ASM_TC = ASM_THREADS * ASM_CPT;   // columns per tile
  __shared__ RowStage stage[ASM_TR];
  const long long r_tile = P.row0 + (long long)blockIdx.y * ASM_TR;
  const int nr = (int)min((long long)ASM_TR, P.row0 + P.nrows - r_tile);
  double gamma, eta, lapfac;
  hat_constants<KIND>(P.ip, P.e2, gamma, eta, lapfac);
  const double gamma2 = 2.0 * gamma;
  if (threadIdx.x < nr) {
    const long long r = r_tile + threadIdx.x;
    RowStage st;
    load_point<MASK>(st.a, P.rows.pts, P.rows.p1[r], P.rows.cphi1 + 5 * r, gamma, eta, lapfac);
    const int p2 = P.rows.p2 ? P.rows.p2[r] : -1;
    st.has_b = p2 >= 0;
    if (st.has_b) load_point<MASK>(st.b, P.rows.pts, p2, P.rows.cphi2 + 5 * r, gamma, eta, lapfac);
    else st.b = st.a;
    st.skip = P.rows.skip ? P.rows.skip[r] : -1;
    stage[threadIdx.x] = st;
  }
  const long long j = P.col0 + (long long)blockIdx.x * ASM_TC + ASM_CPT * threadIdx.x;
  const long long jend = P.col0 + P.ncols;
  double cx[ASM_CPT], cy[ASM_CPT];
#pragma unroll
  for (int u = 0; u < ASM_CPT; u += 2) {
    cx[u] = cy[u] = cx[u + 1] = cy[u + 1] = 0.0;
    if (j + u + 1 < jend) {
      const double4 c = *reinterpret_cast<const double4 *>(P.centres + 2 * (j + u));  // 32-byte aligned: j even
      cx[u] = c.x; cy[u] = c.y; cx[u + 1] = c.z; cy[u + 1] = c.w;
    } else if (j + u < jend) {
      cx[u] = P.centres[2 * (j + u)]; cy[u] = P.centres[2 * (j + u) + 1];
    }
  }
  __syncthreads();
  if (j >= jend) return;
  double *o = P.out + (r_tile - P.row0) * P.ld + (j - P.col0);
  const int ncol = (int)min((long long)ASM_CPT, jend - j);
  const int j32 = (int)j;
#pragma unroll 2
  for (int t = 0; t < nr; t++) {
    const RowStage &st = stage[t];
    double v[ASM_CPT];
#pragma unroll
    for (int u = 0; u < ASM_CPT; u++) v[u] = entry_one_point<KIND, MASK, PFIX>(st.a, cx[u], cy[u], P.ip, P.e2, gamma2, eta);
    if (st.has_b) {
#pragma unroll
      for (int u = 0; u < ASM_CPT; u++) v[u] += entry_one_point<KIND, MASK, PFIX>(st.b, cx[u], cy[u], P.ip, P.e2, gamma2, eta);
    }
    if (ncol == ASM_CPT) {
#pragma unroll
      for (int u = 0; u < ASM_CPT; u += 2) __stcs(reinterpret_cast<double2 *>(o + u), make_double2(v[u], v[u + 1]));
    } else {
#pragma unroll
      for (int u = 0; u < ASM_CPT; u++)
        if (u < ncol) __stcs(o + u, v[u]);
    }
    // skipped column (the row's own node, cloud.py:110-112): the one thread that holds it overwrites the
    // entry with 0 after its vector store (same thread, same address: program order) -- one predicated
    // store instead of a select per entry
    const unsigned int sk = (unsigned int)(st.skip - j32);
    if (sk < (unsigned int)ncol) __stcs(o + sk, 0.0);
    o += P.ld;
  }
}

// Monomial columns N..N+M-1 of collocation rows, plus zero padding up to the block edge.
struct PolyParams {
  UpdesRows rows;
  long long row0, nrows, col0, ncols;  // columns >= N (global numbering); rows < N
  double *out;
  long long ld;
  int N, M;
};

__global__ void assemble_poly_cols_kernel(PolyParams P) {
  const long long r = P.row0 + blockIdx.x * (long long)blockDim.y + threadIdx.y;
  if (r >= P.row0 + P.nrows) return;
  for (long long c = P.col0 + threadIdx.x; c < P.col0 + P.ncols; c += blockDim.x) {
    double v = 0.0;
    const int m = (int)(c - P.N);
    if (m < P.M) {
      double jet[5];
      const int p1 = P.rows.p1[r];
      monomial_jet(m, P.rows.pts[2 * (size_t)p1], P.rows.pts[2 * (size_t)p1 + 1], jet);
      const double *c1 = P.rows.cpol1 + 5 * r;
      v = c1[0] * jet[0] + c1[1] * jet[1] + c1[2] * jet[2] + c1[3] * jet[3] + c1[4] * jet[4];
      const int p2 = P.rows.p2 ? P.rows.p2[r] : -1;
      if (p2 >= 0) {
        monomial_jet(m, P.rows.pts[2 * (size_t)p2], P.rows.pts[2 * (size_t)p2 + 1], jet);
        const double *c2 = P.rows.cpol2 + 5 * r;
        v += c2[0] * jet[0] + c2[1] * jet[1] + c2[2] * jet[2] + c2[3] * jet[3] + c2[4] * jet[4];
      }
    }
    P.out[(r - P.row0) * P.ld + (c - P.col0)] = v;
  }
}

// Rows N..N+M-1:  K[N+m][j] = monomial_m(centre_j) for j < N, 0 otherwise (assembly.py:83).
struct PtParams {
  const double *centres;
  long long row0, nrows, col0, ncols;  // rows >= N (global numbering)
  double *out;
  long long ld;
  int N;
};

__global__ void assemble_pt_rows_kernel(PtParams P) {
  const long long c = P.col0 + blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (c >= P.col0 + P.ncols) return;
  double x = 0, y = 0;
  if (c < P.N) { x = P.centres[2 * c]; y = P.centres[2 * c + 1]; }
  for (long long r = P.row0; r < P.row0 + P.nrows; r++) {
    double v = 0.0;
    if (c < P.N) {
      double jet[5];
      monomial_jet((int)(r - P.N), x, y, jet);
      v = jet[0];
    }
    P.out[(r - P.row0) * P.ld + (c - P.col0)] = v;
  }
}

// columns per thread: g_asm_variant bit 0 -> closed-form Laplacian rows use 4 (default 2);
//                     bit 1 -> general jets use 2 (default 4)
static int g_asm_variant = 0;

template <int KIND, int MASK, int CPT>
static void launch_phi_one(const AsmParams &P, cudaStream_t st) {
  dim3 grid((unsigned)((P.ncols + ASM_THREADS * CPT - 1) / (ASM_THREADS * CPT)), (unsigned)((P.nrows + ASM_TR - 1) / ASM_TR));
  // the default kernel r^3 (polyharmonic a = 1) gets its exponent at compile time: no per-entry branches on `a`
  if (KIND == UPDES_RBF_POLYHARMONIC && P.ip == 1) assemble_phi_kernel<KIND, MASK, CPT, 3><<<grid, ASM_THREADS, 0, st>>>(P);
  else assemble_phi_kernel<KIND, MASK, CPT, 0><<<grid, ASM_THREADS, 0, st>>>(P);
}

template <int KIND>
static int launch_phi(int mask, const AsmParams &P, cudaStream_t st) {
  if (mask & JET_ISO) {
    const bool wide = g_asm_variant & 1;
    if (mask & JET_VAL) {
      if (wide) launch_phi_one<KIND, JET_ISO | JET_VAL, 4>(P, st); else launch_phi_one<KIND, JET_ISO | JET_VAL, 2>(P, st);
    } else {
      if (wide) launch_phi_one<KIND, JET_ISO, 4>(P, st); else launch_phi_one<KIND, JET_ISO, 2>(P, st);
    }
    UPDES_LAUNCH_CHECK();
    return 0;
  }
  const bool narrow = g_asm_variant & 2;
#define UPDES_PHI_CASE(M)                                                          \
  case M:                                                                          \
    if (narrow) launch_phi_one<KIND, M, 2>(P, st); else launch_phi_one<KIND, M, 4>(P, st); \
    break;
  switch (mask & 7) {
    UPDES_PHI_CASE(1)
    UPDES_PHI_CASE(2)
    UPDES_PHI_CASE(3)
    UPDES_PHI_CASE(6)
    default:
      if (narrow) launch_phi_one<KIND, 7, 2>(P, st); else launch_phi_one<KIND, 7, 4>(P, st);
      break;
  }
#undef UPDES_PHI_CASE
  UPDES_LAUNCH_CHECK();
  return 0;
}

static int assemble_block_impl(int kind, double param, int N, int M, const double *centres, const UpdesRows *rows,
                               long long row0, long long nrows, long long col0, long long ncols, int jet_mask,
                               double *out, long long ld, cudaStream_t st) {
  if (N <= 0) return -3;
  if (M < 0 || M > 15) return -4;
  if (!centres) return -5;
  if (!rows) return -6;
  if (row0 < 0 || nrows < 0 || row0 + nrows > (long long)N + M) return -7;
  if (col0 < 0 || ncols < 0 || (col0 & 1)) return -9;
  if (!out) return -12;
  if (ld < ncols || (ld & 1)) return -13;
  if (nrows == 0 || ncols == 0) return 0;
  if ((jet_mask & 15) == 0) jet_mask = 7;
  if (jet_mask & JET_ISO) jet_mask &= (JET_ISO | JET_VAL);
  else if (jet_mask & JET_H) jet_mask |= JET_G;   // second derivatives carry the g term

  const long long rc_end = row0 + nrows < N ? row0 + nrows : N;   // collocation rows [row0, rc_end)
  const long long cphi_end = col0 + ncols < N ? col0 + ncols : N;  // rbf columns [col0, cphi_end)
  if (rc_end > row0) {
    if (cphi_end > col0) {
      AsmParams P;
      P.centres = centres; P.rows = *rows; P.row0 = row0; P.nrows = rc_end - row0;
      P.col0 = col0; P.ncols = cphi_end - col0; P.out = out; P.ld = ld;
      P.ip = (int)param; P.e2 = param * param;
      int rc = 0;
      prof_begin(PROF_ASSEMBLE, 8.0 * (double)P.nrows * (double)P.ncols, st);
      UPDES_DISPATCH_KIND(kind, rc = launch_phi<KIND>(jet_mask, P, st));
      prof_end(st);
      if (rc) return rc;
    }
    if (col0 + ncols > N) {
      PolyParams Q;
      Q.rows = *rows; Q.row0 = row0; Q.nrows = rc_end - row0;
      Q.col0 = col0 > N ? col0 : N; Q.ncols = col0 + ncols - Q.col0;
      Q.out = out + (Q.col0 - col0); Q.ld = ld; Q.N = N; Q.M = M;
      dim3 block(32, 8);
      assemble_poly_cols_kernel<<<(unsigned)((Q.nrows + 7) / 8), block, 0, st>>>(Q);
      UPDES_LAUNCH_CHECK();
    }
  }
  if (row0 + nrows > N) {
    PtParams T;
    T.centres = centres; T.row0 = row0 > N ? row0 : N; T.nrows = row0 + nrows - T.row0;
    T.col0 = col0; T.ncols = ncols; T.out = out + (T.row0 - row0) * ld; T.ld = ld; T.N = N;
    assemble_pt_rows_kernel<<<(unsigned)((ncols + 255) / 256), 256, 0, st>>>(T);
    UPDES_LAUNCH_CHECK();
  }
  return 0;
}

}  // namespace updes

extern "C" int updes_assemble_rows(int rbf_kind, double rbf_param, int N, int M, const double *centres,
                                   const UpdesRows *rows, int64_t row0, int64_t nrows, int jet_mask,
                                   double *out, int64_t ld, void *stream) {
  return updes::assemble_block_impl(rbf_kind, rbf_param, N, M, centres, rows, row0, nrows, 0, ld, jet_mask, out, ld,
                                    (cudaStream_t)stream);
}

extern "C" int updes_assemble_block(int rbf_kind, double rbf_param, int N, int M, const double *centres,
                                    const UpdesRows *rows, int64_t row0, int64_t nrows, int64_t col0,
                                    int64_t ncols, int jet_mask, double *out, int64_t ld, void *stream) {
  return updes::assemble_block_impl(rbf_kind, rbf_param, N, M, centres, rows, row0, nrows, col0, ncols, jet_mask, out,
                                    ld, (cudaStream_t)stream);
}

/* tuning hook (tools/, bench experiments): columns per thread of the assembly kernel, see g_asm_variant */
extern "C" int updes_assemble_set_variant(int variant) {
  if (variant < 0 || variant > 3) return -1;
  updes::g_asm_variant = variant;
  return 0;
}

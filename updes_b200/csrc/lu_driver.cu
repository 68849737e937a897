// Host driver of the dense LU with partial pivoting (P K = L U, in place, row-major).
//
// Replaces the reference's inverse + GEMM + QR chain (updes/assembly.py:87-90,:398-401,
// updes/operators.py:612-616) by one factorisation of the collocation system.
//
// Recursive (Toledo-style) blocking: a block of nc columns is split in two; the left half is
// factored recursively, the right half gets U12 = L11^-1 A12 (recursive triangular solve) and the
// Schur update A22 -= L21 U12 (DMMA GEMM), then the right half is factored recursively.  All
// O(n^3) work lands in the GEMM with the largest possible inner dimension.  Row interchanges are
// applied to the rest of the swap range right after each base panel (32 pivots at a time), which
// keeps every interchange kernel short and fully coalesced in the row-major layout.
//
// The same routine factors (a) the whole matrix on one GPU (swap range = all columns) and (b) one
// column panel of a column-block-cyclic local matrix on the multi-GPU path (swap range = the panel
// only; the other columns, local and remote, get the interchanges when the panel is applied).
#include "lu.cuh"

namespace updes {

int lu_recursive(UpdesLU *h, int v, int64_t r0, int64_t c0, int64_t nc, int64_t swap_lo, int64_t swap_hi,
                 int32_t *ipiv, int32_t *info, cudaStream_t st) {
  if (nc <= 0) return 0;
  const int64_t rows = h->view[v].rows;
  const int W = panel_width_for(h, rows - r0);
  if (W < 8) return -2;    // panel taller than the register-resident kernel supports (> 378 880 rows)
  if (nc <= W) {
    int rc = lu_panel_base(h, v, r0, c0, (int)nc, ipiv, info, st);
    if (rc) return rc;
    return swap_rows_hole(h, v, swap_lo, swap_hi - swap_lo, c0, nc, r0, nc, ipiv, st);   // left and right of the panel at once
  }
  int64_t n1;
  if (nc <= 16) n1 = 8;           // only reached when W == 8 (very tall panels)
  else if (nc <= 32) n1 = 16;
  else n1 = (nc / 2 + 31) / 32 * 32;
  int rc = lu_recursive(h, v, r0, c0, n1, swap_lo, swap_hi, ipiv, info, st);
  if (rc) return rc;
  const int64_t n2 = nc - n1;
  rc = trsm_unit_lower(h, v, r0, c0, n1, v, r0, c0 + n1, n2, st);
  if (rc) return rc;
  if (n1 == 8) rc = rank8_update(h, v, r0 + n1, c0, r0, c0 + n1, r0 + n1, c0 + n1, rows - (r0 + n1), (int)n2, st);
  else rc = dgemm_sub(h, v, r0 + n1, c0, v, r0, c0 + n1, v, r0 + n1, c0 + n1, rows - (r0 + n1), n2, n1, st);
  if (rc) return rc;
  return lu_recursive(h, v, r0 + n1, c0 + n1, n2, swap_lo, swap_hi, ipiv, info, st);
}

}  // namespace updes

extern "C" int updes_lu_create(UpdesLU **handle, int64_t n, int64_t ld) {
  if (!handle) return -1;
  if (n <= 0 || n > 0x7fffffff) return -2;
  if (ld < 16 || (ld % 16)) return -3;
  UpdesLU *h = new UpdesLU();
  h->n = n; h->ld = ld;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, dev);
  // ONE device allocation for the whole workspace (cudaMalloc / cudaFree synchronise the device and cost ~0.1-1 ms
  // each: eleven of them were a visible share of a small factorisation), carved at 256-byte boundaries
  const size_t W = updes::PANEL_W;
  const size_t nflags = (size_t)((n + 127) / 128) + 1;
  size_t off = 0;
  auto carve = [&off](size_t bytes) { const size_t o = off; off += (bytes + 255) / 256 * 256; return o; };
  const size_t o_cand = carve(sizeof(double) * 2 * h->num_sms * W);
  const size_t o_top = carve(sizeof(double) * 2 * W);
  const size_t o_candval = carve(sizeof(double) * 2 * h->num_sms);
  const size_t o_candrow = carve(sizeof(int32_t) * 2 * h->num_sms);
  const size_t o_barrier = carve(sizeof(unsigned int));
  const size_t o_counters = carve(sizeof(unsigned int) * UPDES_GEMM_COUNTERS);
  const size_t o_flags = carve(sizeof(unsigned int) * nflags);
  const size_t o_ticket = carve(sizeof(unsigned int));
  const size_t o_err = carve(sizeof(int));
  const size_t zero_bytes = off;                       // everything above starts at zero
  const size_t o_perm = carve(sizeof(int32_t) * n);
  const size_t o_xbuf = carve(sizeof(double) * (n + 1) * 8);   // X and Y, 4 right-hand sides each (even stride)
  char *base = nullptr;
  if (e == cudaSuccess) e = cudaMalloc(&base, off);
  if (e == cudaSuccess) e = cudaMemset(base, 0, zero_bytes);
  if (e != cudaSuccess) {
    delete h;
    return (int)e;
  }
  h->workspace = base;
  h->cand = (double *)(base + o_cand); h->top = (double *)(base + o_top); h->candval = (double *)(base + o_candval);
  h->candrow = (int32_t *)(base + o_candrow); h->barrier = (unsigned int *)(base + o_barrier);
  h->gemm_counters = (unsigned int *)(base + o_counters); h->sweep_flags = (unsigned int *)(base + o_flags);
  h->sweep_ticket = (unsigned int *)(base + o_ticket); h->sweep_err = (int *)(base + o_err);
  h->perm = (int32_t *)(base + o_perm); h->xbuf = (double *)(base + o_xbuf);
  *handle = h;
  return 0;
}

extern "C" int updes_lu_destroy(UpdesLU *h) {
  if (!h) return 0;
  cudaFree(h->workspace);
  delete h;
  return 0;
}

extern "C" int updes_lu_bind(UpdesLU *h, int slot, double *ptr, int64_t rows, int64_t ld) {
  if (!h) return -1;
  return updes::lu_bind_view(h, slot, ptr, rows, ld);
}

static int lu_factor_impl(UpdesLU *h, double *K, int32_t *ipiv, double *scale, int32_t *info, cudaStream_t st) {
  if (!h) return -1;
  if (!K || (((uintptr_t)K) & 127)) return -2;   // rows must be whole 128-byte lines
  if (!ipiv) return -3;
  if (!info) return -5;
  if (h->ld < h->n) return -1;
  UPDES_CUDA_TRY(cudaMemsetAsync(info, 0, sizeof(int32_t), st));
  int rc = updes::lu_bind_view(h, 0, K, h->n, h->ld);
  if (rc) return rc;
  h->row_scale = nullptr;
  if (scale) {
    // row equilibration: scale[] doubles as the scratch for the row maxima
    rc = updes::row_absmax(K, h->n, h->n, h->ld, scale, st);
    if (rc) return rc;
    rc = updes::scale_from_absmax(scale, h->n, scale, st);
    if (rc) return rc;
    rc = updes::row_scale(K, h->n, h->n, h->ld, scale, st);
    if (rc) return rc;
    h->row_scale = scale;
  }
  rc = updes::lu_recursive(h, 0, 0, 0, h->n, 0, h->n, ipiv, info, st);
  if (rc) return rc;
  return updes::build_permutation(h, ipiv, st);
}

extern "C" int updes_lu_factor(UpdesLU *h, double *K, int32_t *ipiv, int32_t *info, void *stream) {
  int rc = lu_factor_impl(h, K, ipiv, nullptr, info, (cudaStream_t)stream);
  return rc == -5 ? -4 : rc;
}

extern "C" int updes_lu_factor_scaled(UpdesLU *h, double *K, int32_t *ipiv, double *scale, int32_t *info, void *stream) {
  if (!scale) return -4;
  return lu_factor_impl(h, K, ipiv, scale, info, (cudaStream_t)stream);
}

extern "C" int updes_lu_set_row_scale(UpdesLU *h, const double *scale) {
  if (!h) return -1;
  h->row_scale = scale;
  return 0;
}

extern "C" int updes_row_absmax(const double *A, int64_t rows, int64_t cols, int64_t ld, double *out, void *stream) {
  if (!A) return -1;
  if (rows < 0) return -2;
  if (cols < 0 || cols > ld) return -3;
  if (ld & 1) return -4;
  if (!out) return -5;
  return updes::row_absmax(A, rows, cols, ld, out, (cudaStream_t)stream);
}

extern "C" int updes_scale_from_absmax(const double *absmax, int64_t n, double *scale, void *stream) {
  if (!absmax) return -1;
  if (n < 0) return -2;
  if (!scale) return -3;
  return updes::scale_from_absmax(absmax, n, scale, (cudaStream_t)stream);
}

extern "C" int updes_row_scale(double *A, int64_t rows, int64_t cols, int64_t ld, const double *scale, void *stream) {
  if (!A) return -1;
  if (rows < 0) return -2;
  if (cols < 0 || cols > ld) return -3;
  if (ld & 1) return -4;
  if (!scale) return -5;
  return updes::row_scale(A, rows, cols, ld, scale, (cudaStream_t)stream);
}

/* Internal-failure flags of the handle's kernels, read back synchronously on `stream`: bit 0 = a wait inside a
 * triangular sweep timed out.  (A timed-out grid barrier of the panel kernel is reported as info = -1.) */
extern "C" int updes_lu_status(UpdesLU *h, int32_t *host_flags, void *stream) {
  if (!h) return -1;
  if (!host_flags) return -2;
  int v = 0;
  UPDES_CUDA_TRY(cudaMemcpyAsync(&v, h->sweep_err, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  UPDES_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
  *host_flags = v;
  return 0;
}

extern "C" int updes_lu_panel(UpdesLU *h, double *K, int64_t r0, int64_t nc, int32_t *ipiv, int32_t *info,
                              void *stream) {
  if (!h) return -1;
  if (!K) return -2;
  int rc = updes::lu_bind_view(h, 0, K, h->n, h->ld);
  if (rc) return rc;
  return updes::lu_panel_base(h, 0, r0, r0, (int)nc, ipiv, info, (cudaStream_t)stream);
}

// ---- building blocks of the distributed (column-block-cyclic) factorisation ------------------------
extern "C" int updes_lu_panel_factor(UpdesLU *h, int slot, int64_t r0, int64_t c0, int64_t nc, int32_t *ipiv,
                                     int32_t *info, void *stream) {
  if (!h) return -1;
  if (slot < 0 || slot >= UPDES_MAX_VIEWS || !h->view[slot].ptr) return -2;
  if (r0 < 0 || r0 >= h->view[slot].rows) return -3;
  if (c0 < 0 || (c0 & 1) || c0 + nc > h->view[slot].ld) return -4;
  if (!ipiv) return -6;
  if (!info) return -7;
  return updes::lu_recursive(h, slot, r0, c0, nc, c0, c0 + nc, ipiv, info, (cudaStream_t)stream);
}

extern "C" int updes_lu_apply_swaps(UpdesLU *h, int slot, int64_t c_lo, int64_t c_hi, int64_t k0, int64_t npiv,
                                    const int32_t *ipiv, void *stream) {
  if (!h) return -1;
  if (slot < 0 || slot >= UPDES_MAX_VIEWS || !h->view[slot].ptr) return -2;
  if (c_lo < 0 || (c_lo & 1) || c_hi > h->view[slot].ld) return -3;
  if (!ipiv) return -7;
  return updes::swap_rows(h, slot, c_lo, c_hi - c_lo, k0, npiv, ipiv, (cudaStream_t)stream);
}

extern "C" int updes_lu_trsm(UpdesLU *h, int slot_l, int64_t rl, int64_t cl, int64_t n1, int slot_b, int64_t rb,
                             int64_t cb, int64_t ncols, void *stream) {
  if (!h) return -1;
  if (slot_l < 0 || slot_l >= UPDES_MAX_VIEWS) return -2;
  if (slot_b < 0 || slot_b >= UPDES_MAX_VIEWS) return -6;
  if (n1 > 32 && (n1 % 32)) return -5;
  if (n1 <= 32 && n1 != 32 && n1 != 16) return -5;
  return updes::trsm_unit_lower(h, slot_l, rl, cl, n1, slot_b, rb, cb, ncols, (cudaStream_t)stream);
}

extern "C" const char *updes_b200_version(void) { return "updes_b200 0.2 (sm_100a)"; }
extern "C" int64_t updes_launch_count(void) { return updes::g_launch_count; }

// Internal declarations of the dense LU (handle, building blocks).  Public ABI: include/updes_b200.h
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "common.cuh"

// A row-major buffer the LU kernels can address: the (local) matrix itself (slot 0) or an
// auxiliary panel buffer (slots 1..3) that holds a broadcast panel on the multi-GPU path.
struct MatView {
  double *ptr = nullptr;
  int64_t ld = 0, rows = 0;
  CUtensorMap mapA;        // 2D, box {16 k, 128 rows}, 128B swizzle: left (L) operand tiles
  CUtensorMap mapB;        // 2D, box {16 cols, 16 k-rows}, 128B swizzle: right (U) operand tiles
};

constexpr int UPDES_MAX_VIEWS = 4;
constexpr int UPDES_GEMM_COUNTERS = 1024;

struct UpdesLU {
  int64_t n = 0, ld = 0;   // rows and leading dimension of slot 0
  int num_sms = 148;
  int gemm_ctas = 0;       // 0 = one CTA per SM; smaller leaves SMs free for concurrent NCCL kernels
  int64_t panel_cap = 0;   // rows a 32-wide register-resident panel can hold (0 = num_sms * 640); test hook
  int panel_variant = 2;   // 2 (default): implicit-pivoting kernels (cluster push exchange for <= 8 192 rows, grid kernel above);
                           // 1: first-generation cluster + grid kernels; 0: first-generation grid kernel only
  int trsm_base_rows = 128; // largest block of the triangular solve handled by one substitution kernel (32: first generation)
  int gemm_kdeep = 1;      // 1 (default): 32-deep pipeline stages (two 16-k sub-tiles per barrier round) when k % 32 == 0
  int gemm_epilogue = 1;   // 1 (default): fire-and-forget red.global.add.f64 (the SM never waits for C); 0: read-modify-write
  int gemm_c_prefetch = 1; // read-modify-write epilogue only: pull each C tile into L2 one pipeline stage ahead
  int gemm_variant = 1;    // 1 (default): ping-pong, two 128x64 CTAs per SM; 0: one 128x128 CTA per SM
  MatView view[UPDES_MAX_VIEWS];
  void *workspace = nullptr;       // the one device allocation every pointer below points into
  // device workspace of the panel kernel
  double *cand = nullptr;          // [2][num_sms][PANEL_W] candidate pivot rows
  double *top = nullptr;           // [2][PANEL_W] row currently at the diagonal position
  double *candval = nullptr;       // [2][num_sms]
  int32_t *candrow = nullptr;      // [2][num_sms]
  unsigned int *barrier = nullptr; // grid barrier counter (monotonic)
  unsigned int barrier_count = 0;  // host mirror of the counter after all enqueued panels
  unsigned int *gemm_counters = nullptr;   // ring of per-launch tile counters (dynamic scheduler)
  unsigned long long gemm_launch_id = 0;
  int solve_variant = 2;           // 2: row-block streaming sweeps (default); 1: step-synchronous persistent sweeps; 0: one launch per 128-row block
  unsigned int *sweep_ticket = nullptr;  // dynamic row-block counter of the streaming sweep (monotonic)
  unsigned int sweep_ticket_count = 0;   // host mirror
  unsigned int *sweep_flags = nullptr;   // [ceil(n/128)] publication flags of the persistent sweep
  unsigned int sweep_epoch = 0;
  int *sweep_err = nullptr;
  int32_t *perm = nullptr;         // [n] row permutation of the last factorisation (for solves)
  const double *row_scale = nullptr;   // [n] caller-owned row equilibration factors applied to right-hand sides, or null
  double *xbuf = nullptr;          // solve scratch
};

namespace updes {

constexpr int PANEL_W = 32;   // widest base panel

int lu_bind_view(UpdesLU *h, int slot, const double *ptr, int64_t rows, int64_t ld);
// C[rc.., cc..] -= A[ra.., ca..] (m x k, view va) * B[rb.., cb..] (k x n, view vb); C in view vc
int dgemm_sub(UpdesLU *h, int va, int64_t ra, int64_t ca, int vb, int64_t rb, int64_t cb, int vc, int64_t rc,
              int64_t cc, int64_t m, int64_t n, int64_t k, cudaStream_t st);
int panel_width_for(const UpdesLU *h, int64_t m);
// base panel: rows [r0, rows) x columns [c0, c0+jb) of view v; pivots -> ipiv[r0 .. r0+jb)
int lu_panel_base(UpdesLU *h, int v, int64_t r0, int64_t c0, int jb, int32_t *ipiv, int32_t *info, cudaStream_t st);
int swap_rows(UpdesLU *h, int v, int64_t c0, int64_t ncols, int64_t k0, int64_t npiv, const int32_t *ipiv,
              cudaStream_t st);
int swap_rows_hole(UpdesLU *h, int v, int64_t c0, int64_t ncols, int64_t hole0, int64_t holew, int64_t k0, int64_t npiv,
                   const int32_t *ipiv, cudaStream_t st);
// B <- L^-1 B, L = unit-lower n1 x n1 at (rl, cl) of view vl, B = n1 x ncols at (rb, cb) of view vb
int trsm_unit_lower(UpdesLU *h, int vl, int64_t rl, int64_t cl, int64_t n1, int vb, int64_t rb, int64_t cb,
                    int64_t ncols, cudaStream_t st);
// recursive LU of rows [r0, rows) x columns [c0, c0+nc) of view v; interchanges are applied to the
// columns [swap_lo, swap_hi) of the same view (the panel itself included)
int lu_recursive(UpdesLU *h, int v, int64_t r0, int64_t c0, int64_t nc, int64_t swap_lo, int64_t swap_hi,
                 int32_t *ipiv, int32_t *info, cudaStream_t st);
int build_permutation(UpdesLU *h, const int32_t *ipiv, cudaStream_t st);
int row_absmax(const double *A, int64_t rows, int64_t cols, int64_t ld, double *out, cudaStream_t st);
int scale_from_absmax(const double *absmax, int64_t n, double *scale, cudaStream_t st);
int row_scale(double *A, int64_t rows, int64_t cols, int64_t ld, const double *scale, cudaStream_t st);
int rank8_update(UpdesLU *h, int v, int64_t ra, int64_t ca, int64_t rb, int64_t cb, int64_t rc, int64_t cc, int64_t m,
                 int nc, cudaStream_t st);

}  // namespace updes

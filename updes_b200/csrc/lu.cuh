// Internal declarations of the dense LU (handle, building blocks).  Public ABI: include/updes_b200.h
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "common.cuh"

struct UpdesLU {
  int64_t n = 0, ld = 0;
  int num_sms = 148;
  // TMA descriptors over the whole matrix, rebuilt when the bound pointer changes
  const double *bound = nullptr;
  CUtensorMap mapA;        // 2D, box {16 k, 128 rows}, 128B swizzle: L operand tiles
  CUtensorMap mapB;        // 2D, box {16 cols, 16 k-rows}, 128B swizzle: U operand tiles
  // device workspace of the panel kernel
  double *cand = nullptr;          // [2][num_sms][PANEL_W] candidate pivot rows
  double *top = nullptr;           // [2][PANEL_W] row currently at the diagonal position
  double *candval = nullptr;       // [2][num_sms]
  int32_t *candrow = nullptr;      // [2][num_sms]
  unsigned int *barrier = nullptr; // grid barrier counter (monotonic)
  unsigned int barrier_count = 0;  // host mirror of the counter after all enqueued panels
  int32_t *perm = nullptr;         // [n] scratch for solves
  double *xbuf = nullptr;          // solve scratch
};

namespace updes {

constexpr int PANEL_W = 32;   // widest base panel

int lu_bind(UpdesLU *h, const double *K);
int dgemm_sub(UpdesLU *h, double *K, int64_t rc, int64_t cc, int64_t ra, int64_t ca, int64_t rb, int64_t cb,
              int64_t m, int64_t n, int64_t k, cudaStream_t st);
int panel_width_for(const UpdesLU *h, int64_t m);
int lu_panel_base(UpdesLU *h, double *K, int64_t r0, int jb, int32_t *ipiv, int32_t *info, cudaStream_t st);
int swap_rows(UpdesLU *h, double *K, int64_t c0, int64_t ncols, int64_t k0, int64_t npiv, const int32_t *ipiv,
              cudaStream_t st);
int trsm_unit_lower(UpdesLU *h, double *K, int64_t r0, int64_t n1, int64_t c0, int64_t ncols, cudaStream_t st);
int lu_recursive(UpdesLU *h, double *K, int64_t r0, int64_t nc, int32_t *ipiv, int32_t *info, cudaStream_t st);

}  // namespace updes

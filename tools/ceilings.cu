// Ceiling probe for the B200 box (not product code): measures the FP64 roofs the
// hot path is judged against.  Build: see tools/Makefile.  Output: one JSON object.
//   - cuBLAS DGEMM (the realistic FP64 tensor ceiling at sustained clocks)
//   - raw DMMA (mma.sync m8n8k4 f64) issue rate and raw DFMA issue rate
//   - HBM write-only bandwidth (fill kernel, 16-byte stores)
//   - cuSOLVER Dgetrf at a few n (the library comparator for the LU)
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cublas_v2.h>
#include <cusolverDn.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__global__ void dmma_kernel(double* out, int iters) {
  double a = threadIdx.x * 1e-9, b = 1.0 + threadIdx.x * 1e-9;
  double c[16][2];
#pragma unroll
  for (int i = 0; i < 16; i++) { c[i][0] = 0.0; c[i][1] = 0.0; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) {
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dfma_kernel(double* out, int iters) {
  double a = 1.0 + threadIdx.x * 1e-12, b = threadIdx.x * 1e-9;
  double c[16];
#pragma unroll
  for (int i = 0; i < 16; i++) c[i] = i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void fill_kernel(double2* p, size_t n2, double v) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  double2 val = make_double2(v, v);
  for (; i < n2; i += stride) __stcs(p + i, val);
}

__global__ void init_kernel(double* p, size_t n, size_t ld, size_t rows) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    unsigned long long h = i * 0x9E3779B97F4A7C15ull; h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
    double v = (double)(h & 0xFFFFF) / 1048576.0 - 0.5;
    size_t r = i % ld, c = i / ld;
    if (r == c) v += 4.0;
    p[i] = v;
  }
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms; CK(cudaEventElapsedTime(&ms, a, b)); return ms; }

// cuSOLVER Dgetrf alone at one (large) size: the library comparator at the headline n.
static int getrf_only(size_t n) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  cusolverDnHandle_t so; cusolverDnCreate(&so);
  double* A; int* ipiv; int* info; double* work; int lwork = 0;
  CK(cudaMalloc(&A, n * n * 8)); CK(cudaMalloc(&ipiv, n * 4)); CK(cudaMalloc(&info, 4));
  cusolverDnDgetrf_bufferSize(so, n, n, A, n, &lwork);
  CK(cudaMalloc(&work, (size_t)lwork * 8));
  init_kernel<<<148 * 8, 256>>>(A, n * n, n, n);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  cusolverDnDgetrf(so, n, n, A, n, work, ipiv, info);
  CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
  float ms = time_ms(e0, e1);
  int hinfo = -1; CK(cudaMemcpy(&hinfo, info, 4, cudaMemcpyDeviceToHost));
  printf("{\"cusolver_getrf_n%zu\": {\"ms\": %.1f, \"tflops\": %.2f, \"info\": %d, \"lwork_doubles\": %d}}\n", n, ms,
         2.0 / 3.0 * n * n * n / ms * 1e-9, hinfo, lwork);
  return 0;
}

int main(int argc, char** argv) {
  int big = argc > 1 ? atoi(argv[1]) : 16384;
  if (argc > 2) return getrf_only((size_t)atol(argv[2]));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d,\n", prop.name, prop.multiProcessorCount, prop.clockRate);

  // ---- raw DMMA / DFMA issue rates ------------------------------------------------
  {
    double* out; CK(cudaMalloc(&out, sizeof(double) * 148 * 8 * 1024));
    for (int warps = 4; warps <= 16; warps *= 2) {
      int iters = 4096;
      dmma_kernel<<<148 * 2, warps * 32>>>(out, 16); CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e0)); dmma_kernel<<<148 * 2, warps * 32>>>(out, iters); CK(cudaEventRecord(e1));
      CK(cudaDeviceSynchronize());
      float ms = time_ms(e0, e1);
      double flops = 2.0 * 8 * 8 * 4 * 16.0 * iters * warps * 148 * 2;
      printf(" \"dmma_tflops_%dwarps\": %.2f,\n", warps, flops / ms * 1e-9);
    }
    for (int warps = 8; warps <= 32; warps *= 2) {
      int iters = 8192;
      dfma_kernel<<<148 * 2, warps * 32>>>(out, 16); CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(e0)); dfma_kernel<<<148 * 2, warps * 32>>>(out, iters); CK(cudaEventRecord(e1));
      CK(cudaDeviceSynchronize());
      float ms = time_ms(e0, e1);
      double flops = 2.0 * 16.0 * iters * warps * 32 * 148 * 2;
      printf(" \"dfma_tflops_%dwarps\": %.2f,\n", warps, flops / ms * 1e-9);
    }
    CK(cudaFree(out));
  }

  // ---- HBM write-only ---------------------------------------------------------------
  {
    size_t bytes = (size_t)16 << 30;
    double2* p; CK(cudaMalloc(&p, bytes));
    size_t n2 = bytes / 16;
    float best = 1e30f;
    for (int rep = 0; rep < 6; rep++) {
      CK(cudaEventRecord(e0)); fill_kernel<<<148 * 16, 512>>>(p, n2, 1.0 + rep); CK(cudaEventRecord(e1));
      CK(cudaDeviceSynchronize());
      float ms = time_ms(e0, e1); if (rep > 0 && ms < best) best = ms;
    }
    printf(" \"hbm_write_gbs\": %.1f,\n", bytes / best * 1e-6);
    float bestm = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
      CK(cudaEventRecord(e0)); CK(cudaMemsetAsync(p, rep, bytes)); CK(cudaEventRecord(e1));
      CK(cudaDeviceSynchronize());
      float ms = time_ms(e0, e1); if (rep > 0 && ms < bestm) bestm = ms;
    }
    printf(" \"hbm_memset_gbs\": %.1f,\n", bytes / bestm * 1e-6);
    CK(cudaFree(p));
  }

  // ---- cuBLAS DGEMM -----------------------------------------------------------------
  cublasHandle_t bl; cublasCreate(&bl);
  {
    int ns[3] = {4096, 8192, big};
    for (int t = 0; t < 3; t++) {
      size_t n = ns[t];
      double *A, *B, *C;
      CK(cudaMalloc(&A, n * n * 8)); CK(cudaMalloc(&B, n * n * 8)); CK(cudaMalloc(&C, n * n * 8));
      init_kernel<<<148 * 8, 256>>>(A, n * n, n, n); init_kernel<<<148 * 8, 256>>>(B, n * n, n, n);
      CK(cudaMemset(C, 0, n * n * 8));
      double al = -1.0, be = 1.0;
      cublasDgemm(bl, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &al, A, n, B, n, &be, C, n);
      CK(cudaDeviceSynchronize());
      int reps = n >= 16384 ? 3 : 6;
      CK(cudaEventRecord(e0));
      for (int r = 0; r < reps; r++) cublasDgemm(bl, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &al, A, n, B, n, &be, C, n);
      CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
      float ms = time_ms(e0, e1) / reps;
      printf(" \"cublas_dgemm_tflops_n%zu\": %.2f,\n", n, 2.0 * n * n * n / ms * 1e-9);
      // rank-512 update shape (the LU trailing update): m=n=big, k=512
      if (t == 2) {
        int k = 512;
        cublasDgemm(bl, CUBLAS_OP_N, CUBLAS_OP_N, n, n, k, &al, A, n, B, n, &be, C, n);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        for (int r = 0; r < 10; r++) cublasDgemm(bl, CUBLAS_OP_N, CUBLAS_OP_N, n, n, k, &al, A, n, B, n, &be, C, n);
        CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
        ms = time_ms(e0, e1) / 10;
        printf(" \"cublas_dgemm_tflops_k512\": %.2f,\n", 2.0 * n * n * k / ms * 1e-9);
      }
      CK(cudaFree(A)); CK(cudaFree(B)); CK(cudaFree(C));
    }
  }

  // ---- cuSOLVER Dgetrf ----------------------------------------------------------------
  {
    cusolverDnHandle_t so; cusolverDnCreate(&so);
    int ns[3] = {8192, 16384, 32768};
    for (int t = 0; t < 3; t++) {
      size_t n = ns[t];
      double* A; int* ipiv; int* info; double* work; int lwork = 0;
      CK(cudaMalloc(&A, n * n * 8)); CK(cudaMalloc(&ipiv, n * 4)); CK(cudaMalloc(&info, 4));
      cusolverDnDgetrf_bufferSize(so, n, n, A, n, &lwork);
      CK(cudaMalloc(&work, (size_t)lwork * 8));
      float best = 1e30f;
      for (int rep = 0; rep < 2; rep++) {
        init_kernel<<<148 * 8, 256>>>(A, n * n, n, n);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        cusolverDnDgetrf(so, n, n, A, n, work, ipiv, info);
        CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
        float ms = time_ms(e0, e1); if (ms < best) best = ms;
      }
      printf(" \"cusolver_getrf_n%zu\": {\"ms\": %.1f, \"tflops\": %.2f},\n", n, best, 2.0 / 3.0 * n * n * n / best * 1e-9);
      CK(cudaFree(A)); CK(cudaFree(ipiv)); CK(cudaFree(info)); CK(cudaFree(work));
    }
    cusolverDnDestroy(so);
  }
  cublasDestroy(bl);
  printf(" \"done\": true}\n");
  return 0;
}

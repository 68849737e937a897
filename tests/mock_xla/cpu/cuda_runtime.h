// CPU stand-in for the three CUDA runtime names integration/updes_jax_ffi.cc uses, so that the adapter can be RUN in a
// container without a GPU against the emulated C-ABI (tests/cpu_abi_emulation.py).  Test infrastructure.
#ifndef MOCK_CUDA_RUNTIME_H_
#define MOCK_CUDA_RUNTIME_H_
#include <cstddef>
#include <cstring>
typedef void *cudaStream_t;
enum cudaMemcpyKind { cudaMemcpyDeviceToDevice = 3 };
static inline int cudaMemcpyAsync(void *dst, const void *src, size_t n, cudaMemcpyKind, cudaStream_t) { std::memcpy(dst, src, n); return 0; }
static inline int cudaStreamSynchronize(cudaStream_t) { return 0; }
#endif

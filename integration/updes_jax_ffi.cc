// XLA FFI handlers over libupdes_b200.so.  No jaxlib / XLA FFI headers exist in this environment: the file is compiled
// and run by the tests against a mock of the binding API (tests/mock_xla/xla/ffi/api/ffi.h), not against real XLA.
// See integration/README.md for the build line and INTEGRATION.md for context.
// Replaces, on the reference side: assemble_op_Phi_P / assemble_bd_Phi_P / assemble_A block assembly
// (updes/assembly.py:10-362), inv + GEMM + lineax QR (updes/assembly.py:87-90,:398-401,
// updes/operators.py:612-616) and the field evaluators (updes/operators.py:118-351).
#include <cuda_runtime.h>

#include "updes_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

static ffi::Error AssembleImpl(cudaStream_t stream, int32_t kind, double param, int32_t M, int32_t mask_internal,
                               int32_t mask_boundary, int32_t Ni, ffi::Buffer<ffi::F64> centres,
                               ffi::Buffer<ffi::S32> p1, ffi::Buffer<ffi::S32> p2, ffi::Buffer<ffi::F64> cphi1,
                               ffi::Buffer<ffi::F64> cphi2, ffi::Buffer<ffi::F64> cpol1, ffi::Buffer<ffi::F64> cpol2,
                               ffi::Buffer<ffi::S32> skip, ffi::ResultBuffer<ffi::F64> K) {
  const int N = static_cast<int>(centres.dimensions()[0]);
  const int64_t n = K->dimensions()[0], ld = K->dimensions()[1];
  UpdesRows rows{centres.typed_data(), p1.typed_data(), p2.typed_data(), cphi1.typed_data(), cphi2.typed_data(),
                 cpol1.typed_data(), cpol2.typed_data(), skip.typed_data()};
  // three row ranges so each gets the narrowest jet specialisation (updes_b200/assembly.py:assemble_system)
  const int64_t r0[3] = {0, Ni, N}, nr[3] = {Ni, N - Ni, M};
  const int mask[3] = {mask_internal, mask_boundary, 7};
  for (int t = 0; t < 3; t++) {
    if (nr[t] <= 0) continue;
    int rc = updes_assemble_rows(kind, param, N, M, centres.typed_data(), &rows, r0[t], nr[t], mask[t],
                                 K->typed_data() + r0[t] * ld, ld, stream);
    if (rc) return ffi::Error::Internal("updes_assemble_rows failed");
  }
  (void)n;
  return ffi::Error::Success();
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(
    UpdesAssemble, AssembleImpl,
    ffi::Ffi::Bind()
        .Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Attr<int32_t>("kind").Attr<double>("param").Attr<int32_t>("M")
        .Attr<int32_t>("mask_internal").Attr<int32_t>("mask_boundary").Attr<int32_t>("Ni")
        .Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::S32>>().Arg<ffi::Buffer<ffi::S32>>()
        .Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>()
        .Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::S32>>()
        .Ret<ffi::Buffer<ffi::F64>>());

// In-place LU + one solve.  K is donated (input_output_aliases={0: 0}) so XLA does not copy 65 GB.
static ffi::Error FactorSolveImpl(cudaStream_t stream, ffi::Buffer<ffi::F64> K, ffi::Buffer<ffi::F64> rhs,
                                  ffi::ResultBuffer<ffi::F64> LU, ffi::ResultBuffer<ffi::F64> x,
                                  ffi::ResultBuffer<ffi::S32> ipiv, ffi::ResultBuffer<ffi::S32> info) {
  const int64_t n = K.dimensions()[0], ld = K.dimensions()[1];
  UpdesLU* h = nullptr;
  if (updes_lu_create(&h, n, ld)) return ffi::Error::Internal("updes_lu_create failed");
  cudaMemcpyAsync(x->typed_data(), rhs.typed_data(), sizeof(double) * n, cudaMemcpyDeviceToDevice, stream);
  int rc = updes_lu_factor(h, LU->typed_data(), ipiv->typed_data(), info->typed_data(), stream);
  if (!rc) rc = updes_lu_solve(h, LU->typed_data(), ipiv->typed_data(), x->typed_data(), n, 1, 0, stream);
  cudaStreamSynchronize(stream);   // the handle owns scratch used by the enqueued kernels
  updes_lu_destroy(h);
  return rc == 0 ? ffi::Error::Success() : ffi::Error::Internal("updes LU failed");
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(UpdesFactorSolve, FactorSolveImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::S32>>().Ret<ffi::Buffer<ffi::S32>>());

// Matrix-free jets: value / gradient / laplacian of fields given by coefficients.
static ffi::Error EvalJetsImpl(cudaStream_t stream, int32_t kind, double param, int32_t M,
                               ffi::Buffer<ffi::F64> centres, ffi::Buffer<ffi::F64> coeffs, ffi::Buffer<ffi::F64> pts,
                               ffi::ResultBuffer<ffi::F64> jphi, ffi::ResultBuffer<ffi::F64> jpol,
                               ffi::ResultBuffer<ffi::F64> workspace) {
  const int N = static_cast<int>(centres.dimensions()[0]);
  const int nf = static_cast<int>(coeffs.dimensions()[0]);
  const int64_t ldc = coeffs.dimensions()[1];
  const int npts = static_cast<int>(pts.dimensions()[0]);
  int rc = updes_eval_jets(kind, param, N, M, centres.typed_data(), coeffs.typed_data(), ldc, nf, pts.typed_data(), npts,
                           nullptr, jphi->typed_data(), jpol->typed_data(), workspace->typed_data(), stream);
  return rc == 0 ? ffi::Error::Success() : ffi::Error::Internal("updes_eval_jets failed");
}

XLA_FFI_DEFINE_HANDLER_SYMBOL(UpdesEvalJets, EvalJetsImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int32_t>("kind").Attr<double>("param").Attr<int32_t>("M")
                                  .Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>());

// Plain C helper for the Python side: size of the workspace result buffer UpdesEvalJets needs.
extern "C" size_t UpdesEvalJetsWorkspaceBytes(int N, int npts, int nf) { return updes_eval_jets_workspace_bytes(N, npts, nf); }

// Matrix-free jet sums: for every evaluation point, sum_j c_j * jet(x_i, centre_j).
//
// Replaces the reference's field evaluators value / gradient / laplacian / divergence
// (updes/operators.py:118-351), which vmap a per-(x, centre) autodiff call over all centres, and
// gives K.c and [Phi P].c without a stored matrix (residuals, SteadySol.vals).
//
// Compute-bound FP64 (no matrix traffic): one thread per evaluation point, centres and their
// coefficients streamed through shared memory in chunks, the centre range optionally split across
// CTAs (partials reduced in a fixed order so results are deterministic).
#include "common.cuh"

namespace updes {

constexpr int EJ_THREADS = 128;
constexpr int EJ_CHUNK = 256;

struct EjParams {
  const double *centres;
  const double *coeffs;   // [nf][ldc]
  long long ldc;
  const double *pts;
  const int32_t *skip;
  double *partial;        // [nsplit][nf][npts][5]
  int N, npts, nsplit, per_split;
  int ip;
  double e2;
};

template <int KIND, int NF>
__global__ void __launch_bounds__(EJ_THREADS) eval_jets_kernel(EjParams P, int f0) {
  __shared__ double sx[EJ_CHUNK], sy[EJ_CHUNK], sc[NF][EJ_CHUNK];
  const int i = blockIdx.x * EJ_THREADS + threadIdx.x;
  const bool live = i < P.npts;
  const double x = live ? P.pts[2 * (size_t)i] : 0.0, y = live ? P.pts[2 * (size_t)i + 1] : 0.0;
  const int skip = (live && P.skip) ? P.skip[i] : -1;
  const int jbeg = blockIdx.y * P.per_split;
  const int jend = min(P.N, jbeg + P.per_split);
  double acc[NF][5];
#pragma unroll
  for (int f = 0; f < NF; f++)
#pragma unroll
    for (int k = 0; k < 5; k++) acc[f][k] = 0.0;

  for (int j0 = jbeg; j0 < jend; j0 += EJ_CHUNK) {
    const int cnt = min(EJ_CHUNK, jend - j0);
    __syncthreads();
    for (int t = threadIdx.x; t < cnt; t += EJ_THREADS) {
      sx[t] = P.centres[2 * (size_t)(j0 + t)];
      sy[t] = P.centres[2 * (size_t)(j0 + t) + 1];
#pragma unroll
      for (int f = 0; f < NF; f++) sc[f][t] = P.coeffs[(size_t)(f0 + f) * P.ldc + j0 + t];
    }
    __syncthreads();
    if (!live) continue;
    for (int t = 0; t < cnt; t++) {
      const double dx = x - sx[t], dy = y - sy[t];
      const double dx2 = dx * dx, dy2 = dy * dy, s = dx2 + dy2;
      double phi, g, h;
      radial<KIND>(s, P.ip, P.e2, phi, g, h);
      if (s == 0.0) { phi = phi_at_zero<KIND>(); g = 0.0; h = 0.0; }
      if (j0 + t == skip) { phi = 0.0; g = 0.0; h = 0.0; }
      const double jx = g * dx, jy = g * dy, jxx = fma(h, dx2, g), jyy = fma(h, dy2, g);
#pragma unroll
      for (int f = 0; f < NF; f++) {
        const double c = sc[f][t];
        acc[f][0] = fma(c, phi, acc[f][0]);
        acc[f][1] = fma(c, jx, acc[f][1]);
        acc[f][2] = fma(c, jy, acc[f][2]);
        acc[f][3] = fma(c, jxx, acc[f][3]);
        acc[f][4] = fma(c, jyy, acc[f][4]);
      }
    }
  }
  if (!live) return;
  // partial layout: [split][field-in-call (NF)][npts][5]
#pragma unroll
  for (int f = 0; f < NF; f++) {
    double *dst = P.partial + (((size_t)blockIdx.y * NF + f) * P.npts + i) * 5;
#pragma unroll
    for (int k = 0; k < 5; k++) dst[k] = acc[f][k];
  }
}

// Sum the splits (fixed order) into jphi, and evaluate the polynomial part into jpol.
__global__ void eval_jets_finish_kernel(const double *partial, int nsplit, int nfc, int npts, int f0, double *jphi,
                                        double *jpol, const double *coeffs, long long ldc, const double *pts, int N,
                                        int M) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;   // over nfc * npts
  if (idx >= nfc * npts) return;
  const int f = idx / npts, i = idx - f * npts;
  double acc[5] = {0, 0, 0, 0, 0};
  for (int s = 0; s < nsplit; s++) {
    const double *src = partial + (((size_t)s * nfc + f) * npts + i) * 5;
    for (int k = 0; k < 5; k++) acc[k] += src[k];
  }
  double *o = jphi + ((size_t)(f0 + f) * npts + i) * 5;
  for (int k = 0; k < 5; k++) o[k] = acc[k];
  double pol[5] = {0, 0, 0, 0, 0};
  const double x = pts[2 * (size_t)i], y = pts[2 * (size_t)i + 1];
  for (int m = 0; m < M; m++) {
    double jet[5];
    monomial_jet(m, x, y, jet);
    const double c = coeffs[(size_t)(f0 + f) * ldc + N + m];
    for (int k = 0; k < 5; k++) pol[k] = fma(c, jet[k], pol[k]);
  }
  double *q = jpol + ((size_t)(f0 + f) * npts + i) * 5;
  for (int k = 0; k < 5; k++) q[k] = pol[k];
}

static void choose_split(int N, int npts, int &nsplit, int &per_split) {
  const int row_blocks = (npts + EJ_THREADS - 1) / EJ_THREADS;
  int want = (148 * 8 + row_blocks - 1) / row_blocks;         // aim at ~8 CTAs per SM in flight
  const int max_split = (N + EJ_CHUNK - 1) / EJ_CHUNK;
  nsplit = want < 1 ? 1 : (want > max_split ? max_split : want);
  per_split = (N + nsplit - 1) / nsplit;
  per_split = ((per_split + EJ_CHUNK - 1) / EJ_CHUNK) * EJ_CHUNK;
  nsplit = (N + per_split - 1) / per_split;
}

}  // namespace updes

extern "C" size_t updes_eval_jets_workspace_bytes(int N, int npts, int nf) {
  int nsplit, per;
  updes::choose_split(N, npts, nsplit, per);
  const int nfc = nf >= 2 ? 2 : 1;
  return (size_t)nsplit * nfc * (size_t)npts * 5 * sizeof(double);
}

extern "C" int updes_eval_jets(int rbf_kind, double rbf_param, int N, int M, const double *centres,
                               const double *coeffs, int64_t ldc, int nf, const double *pts, int npts,
                               const int32_t *skip, double *jphi, double *jpol, void *workspace, void *stream) {
  using namespace updes;
  if (N <= 0) return -3;
  if (M < 0 || M > 15) return -4;
  if (!centres) return -5;
  if (!coeffs) return -6;
  if (ldc < (int64_t)N + M) return -7;
  if (nf <= 0) return -8;
  if (!pts) return -9;
  if (npts <= 0) return 0;
  if (!jphi) return -12;
  if (!jpol) return -13;
  if (!workspace) return -14;
  cudaStream_t st = (cudaStream_t)stream;
  EjParams P;
  P.centres = centres; P.coeffs = coeffs; P.ldc = ldc; P.pts = pts; P.skip = skip;
  P.partial = (double *)workspace; P.N = N; P.npts = npts;
  choose_split(N, npts, P.nsplit, P.per_split);
  P.ip = (int)rbf_param; P.e2 = rbf_param * rbf_param;
  dim3 grid((npts + EJ_THREADS - 1) / EJ_THREADS, P.nsplit);
  for (int f0 = 0; f0 < nf; f0 += 2) {
    const int nfc = (nf - f0) >= 2 ? 2 : 1;
    if (nfc == 2) {
      UPDES_DISPATCH_KIND(rbf_kind, (eval_jets_kernel<KIND, 2><<<grid, EJ_THREADS, 0, st>>>(P, f0)));
    } else {
      UPDES_DISPATCH_KIND(rbf_kind, (eval_jets_kernel<KIND, 1><<<grid, EJ_THREADS, 0, st>>>(P, f0)));
    }
    UPDES_LAUNCH_CHECK();
    const int tot = nfc * npts;
    eval_jets_finish_kernel<<<(tot + 255) / 256, 256, 0, st>>>(P.partial, P.nsplit, nfc, npts, f0, jphi, jpol, coeffs,
                                                               ldc, pts, N, M);
    UPDES_LAUNCH_CHECK();
  }
  return 0;
}

// Shared device code of libupdes_b200: closed-form RBF jets, monomial jets, launch helpers.
//
// Reference semantics reproduced here (paths under /root/reference/updes):
//   kernels         utils.py:30-69      multiquadric, inverse_multiquadric, gaussian, polyharmonic, thin_plate
//   monomials       utils.py:92-134     15 monomials of degree <= 4
//   term set        operators.py:15-111 nodal_value / nodal_gradient / nodal_laplacian / nodal_div_grad
// The reference differentiates with JAX autodiff and applies nan_to_num, which turns the NaN that
// autodiff produces at r == 0 into 0 (operators.py:58,:83,:109).  The closed forms below therefore
// return zero first and second derivatives at r == 0 and the plain value phi(0) (0 for the
// log-kernel, utils.py:65).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/updes_b200.h"

namespace updes {

extern long long g_launch_count;   // kernels launched by this library (bench.py gpu_launches)

// optional event timing per kernel class (profile.cu); work = flops or bytes of the launch
enum { PROF_GEMM = 0, PROF_PANEL = 1, PROF_SWAP = 2, PROF_TRSM = 3, PROF_ASSEMBLE = 4, PROF_SOLVE = 5, PROF_JETS = 6 };
void prof_begin(int cat, double work, cudaStream_t st);
void prof_end(cudaStream_t st);

#define UPDES_CUDA_TRY(expr)                       \
  do {                                             \
    cudaError_t _e = (expr);                       \
    if (_e != cudaSuccess) return (int)_e;         \
  } while (0)

#define UPDES_LAUNCH_CHECK()                       \
  do {                                             \
    ++::updes::g_launch_count;                     \
    cudaError_t _e = cudaGetLastError();           \
    if (_e != cudaSuccess) return (int)_e;         \
  } while (0)

constexpr int JET_VAL = 1;   // phi needed
constexpr int JET_G = 2;     // first derivatives (and/or the g part of second derivatives) needed
constexpr int JET_H = 4;     // second derivatives needed
constexpr int JET_ISO = 8;   // rows only use c0*phi + c3*(phi_xx + phi_yy): no gradient term, c3 == c4

__device__ __forceinline__ bool is_zero_bits(double s) { return __double_as_longlong(s) == 0ll; }   // s >= 0 here

__device__ __forceinline__ double ipow_u(double x, int e) {
  double v = 1.0;
  for (int k = 0; k < e; k++) v *= x;
  return v;
}

// 1/sqrt(s) in FP64 from the FP32 SFU estimate plus one third-order (Halley) correction:
// 23 bits -> ~69 bits before rounding, about 1 ulp, in 5 FP64 operations instead of the ~12 of
// rsqrt().  Falls back to rsqrt() outside the float range.
__device__ __forceinline__ double fast_rsqrt(double s) {
  // range test on the exponent bits (integer pipe, keeps the FP64 pipe for arithmetic):
  // 2^-100 <= s < 2^100, which also excludes 0, negatives, inf and NaN
  const unsigned int hi = (unsigned int)__double2hiint(s);
  if (hi - 0x39B00000u >= 0x0C800000u) return rsqrt(s);
  const double y = (double)rsqrtf((float)s);
  const double t = s * y;
  const double e = fma(-t, y, 1.0);
  const double q = e * fma(0.375, e, 0.5);
  return fma(y, q, y);
}

// Radial triple of a kernel as a function of s = r^2 > 0:
//   phi(r),  g = phi'(r)/r,  h = (phi''(r) - g)/r^2
// so that  phi_x = g dx,  phi_xx = g + h dx^2,  phi_yy = g + h dy^2.
// ip: integer parameter (a for polyharmonic / thin_plate); e2: eps^2 for the others.
template <int KIND>
__device__ __forceinline__ void radial(double s, int ip, double e2, double &phi, double &g, double &h) {
  if (KIND == UPDES_RBF_POLYHARMONIC) {
    // phi = r^p, p = 2a+1:  g = p r^(p-2),  h = p (p-2) r^(p-4)
    const int p = 2 * ip + 1;
    const double rs = fast_rsqrt(s);
    double base;                       // r^(p-4)
    if (p >= 5) base = ipow_u(s, (p - 5) >> 1) * (s * rs);
    else if (p == 3) base = rs;
    else base = rs * rs * rs;          // p == 1
    h = (double)(p * (p - 2)) * base;
    const double bs = base * s;        // r^(p-2)
    g = (double)p * bs;
    phi = bs * s;
  } else if (KIND == UPDES_RBF_THIN_PLATE) {
    // phi = r^q log r, q = 2a:  g = r^(q-2) (q L + 1),  h = r^(q-4) ((q-2)(q L + 1) + q)
    const int q = 2 * ip;
    const double L = 0.5 * log(s);
    double base;                       // r^(q-4) = s^(a-2)
    if (ip >= 2) base = ipow_u(s, ip - 2);
    else if (ip == 1) base = 1.0 / s;
    else base = 1.0 / (s * s);
    const double t = (double)q * L + 1.0;
    h = base * ((double)(q - 2) * t + (double)q);
    const double bs = base * s;
    g = bs * t;
    phi = bs * s * L;
  } else if (KIND == UPDES_RBF_GAUSSIAN) {
    const double E = exp(-e2 * s);
    phi = E;
    g = -2.0 * e2 * E;
    h = 4.0 * e2 * e2 * E;
  } else if (KIND == UPDES_RBF_MULTIQUADRIC) {
    const double w = fma(e2, s, 1.0);
    const double rw = rsqrt(w);
    phi = w * rw;
    g = e2 * rw;
    h = -(e2 * e2) * (rw * rw * rw);
  } else {  // inverse multiquadric
    const double w = fma(e2, s, 1.0);
    const double rw = rsqrt(w);
    const double rw3 = rw * rw * rw;
    phi = rw;
    g = -e2 * rw3;
    h = 3.0 * (e2 * e2) * (rw3 * rw * rw);
  }
}

// Radial Laplacian  lap(r) = phi'' + phi'/r = 2 g + h s  in closed form (2-D), with phi.
template <int KIND>
__device__ __forceinline__ void radial_lap(double s, int ip, double e2, double &phi, double &lap) {
  if (KIND == UPDES_RBF_POLYHARMONIC) {
    const int p = 2 * ip + 1;                 // lap = p^2 r^(p-2)
    const double rs = fast_rsqrt(s);
    const double r = s * rs;
    double rp2;                               // r^(p-2)
    if (p == 3) rp2 = r;
    else if (p == 1) rp2 = rs;
    else rp2 = ipow_u(s, (p - 3) >> 1) * r;
    lap = (double)(p * p) * rp2;
    phi = rp2 * s;
  } else {
    double g, h;
    radial<KIND>(s, ip, e2, phi, g, h);
    lap = fma(h, s, 2.0 * g);
  }
}

template <int KIND>
__device__ __forceinline__ double phi_at_zero() {
  return (KIND == UPDES_RBF_POLYHARMONIC || KIND == UPDES_RBF_THIN_PLATE) ? 0.0 : 1.0;
}

// Pre-combined coefficients of one evaluation point of one row:
//   entry = c0 phi + g (c1 dx + c2 dy + c34) + h (c3 dx^2 + c4 dy^2),  c34 = c3 + c4
struct RowPoint {
  double x, y;
  double c0, c1, c2, c3, c4, c34;
};

template <int KIND, int MASK>
__device__ __forceinline__ double entry_one_point(const RowPoint &rp, double cx, double cy, int ip, double e2) {
  const double dx = rp.x - cx, dy = rp.y - cy;
  const double dx2 = dx * dx, dy2 = dy * dy;
  const double s = dx2 + dy2;
  if (MASK & JET_ISO) {
    double phi, lap;
    radial_lap<KIND>(s, ip, e2, phi, lap);
    double v = rp.c3 * lap;
    if (MASK & JET_VAL) v = fma(rp.c0, phi, v);
    return is_zero_bits(s) ? ((MASK & JET_VAL) ? rp.c0 * phi_at_zero<KIND>() : 0.0) : v;
  }
  double phi, g, h;
  radial<KIND>(s, ip, e2, phi, g, h);
  double v = 0.0;
  if (MASK & JET_H) v = h * fma(rp.c3, dx2, rp.c4 * dy2);
  if (MASK & JET_G) v = fma(g, fma(rp.c1, dx, fma(rp.c2, dy, rp.c34)), v);
  if (MASK & JET_VAL) v = fma(rp.c0, phi, v);
  // r == 0: derivatives -> 0 (nan_to_num), value -> phi(0)
  return is_zero_bits(s) ? ((MASK & JET_VAL) ? rp.c0 * phi_at_zero<KIND>() : 0.0) : v;
}

// jet of monomial id (utils.py:92-134) at (x, y): value, d/dx, d/dy, d2/dx2, d2/dy2
__device__ __forceinline__ void monomial_jet(int id, double x, double y, double *jet) {
  const int ex[15] = {0, 1, 0, 2, 1, 0, 3, 2, 1, 0, 4, 3, 2, 1, 0};
  const int ey[15] = {0, 0, 1, 0, 1, 2, 0, 1, 2, 3, 0, 1, 2, 3, 4};
  const int a = ex[id], b = ey[id];
  const double xa = ipow_u(x, a), yb = ipow_u(y, b);
  jet[0] = xa * yb;
  jet[1] = a >= 1 ? a * ipow_u(x, a - 1) * yb : 0.0;
  jet[2] = b >= 1 ? b * xa * ipow_u(y, b - 1) : 0.0;
  jet[3] = a >= 2 ? a * (a - 1) * ipow_u(x, a - 2) * yb : 0.0;
  jet[4] = b >= 2 ? b * (b - 1) * xa * ipow_u(y, b - 2) : 0.0;
}

// Dispatch a callable templated on the rbf kind.
#define UPDES_DISPATCH_KIND(kind, ...)                                                     \
  switch (kind) {                                                                          \
    case UPDES_RBF_POLYHARMONIC: { constexpr int KIND = UPDES_RBF_POLYHARMONIC; __VA_ARGS__; break; } \
    case UPDES_RBF_THIN_PLATE: { constexpr int KIND = UPDES_RBF_THIN_PLATE; __VA_ARGS__; break; }     \
    case UPDES_RBF_GAUSSIAN: { constexpr int KIND = UPDES_RBF_GAUSSIAN; __VA_ARGS__; break; }         \
    case UPDES_RBF_MULTIQUADRIC: { constexpr int KIND = UPDES_RBF_MULTIQUADRIC; __VA_ARGS__; break; } \
    case UPDES_RBF_INVERSE_MULTIQUADRIC: { constexpr int KIND = UPDES_RBF_INVERSE_MULTIQUADRIC; __VA_ARGS__; break; } \
    default: return -1;                                                                    \
  }

}  // namespace updes

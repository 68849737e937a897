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

// 1/sqrt(s) in FP64 from the SFU's FP64 estimate (MUFU.RSQ64H: works on the high word of the double, so
// no FP64<->FP32 conversions; ~2^-20 relative after the truncation) plus one third-order (Halley)
// correction: (5/16) e^3 ~ 2^-63 before rounding, i.e. about 1 ulp, in 5 FP64 operations instead of the
// ~12 of rsqrt().  `s` must be a positive normal double.
__device__ __forceinline__ double fast_rsqrt_pos(double s) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(s));
  const double t = s * y;
  const double e = fma(-t, y, 1.0);
  const double q = e * fma(0.375, e, 0.5);
  return fma(y, q, y);
}
// Same for s >= 0 with the seed's input clamped to >= 2^-1021 on the integer pipe: the result stays finite at
// s == 0 (~2^511), so expressions of the form s * rsqrt(s) evaluate to exactly 0 there without a select.
__device__ __forceinline__ double fast_rsqrt_clamped(double s) {
  const int hi = max(__double2hiint(s), 0x00200000);
  const double sc = __hiloint2double(hi, 0);
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(sc));
  const double t = s * y;
  const double e = fma(-t, y, 1.0);
  const double q = e * fma(0.375, e, 0.5);
  return fma(y, q, y);
}
// s >= 0 that is zero or below 2^-1021 (integer pipe test on the exponent bits)
__device__ __forceinline__ bool tiny_or_zero(double s) { return (unsigned int)__double2hiint(s) < 0x00200000u; }
__device__ __forceinline__ double fast_rsqrt(double s) { return tiny_or_zero(s) ? rsqrt(s) : fast_rsqrt_pos(s); }

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
    const double rw = fast_rsqrt_pos(w);
    phi = w * rw;
    g = e2 * rw;
    h = -(e2 * e2) * (rw * rw * rw);
  } else {  // inverse multiquadric
    const double w = fma(e2, s, 1.0);
    const double rw = fast_rsqrt_pos(w);
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

// ---- assembly entry with constants folded into the row coefficients ---------------------------------
// For every kernel the radial triple factors as  g = gamma * ghat,  h = eta * hhat  with constants
// gamma, eta (polyharmonic p = 2a+1: gamma = p, eta = p(p-2), ghat = r^(p-2), hhat = r^(p-4); gaussian:
// ghat = hhat = phi; ...).  The assembly kernel folds gamma / eta into the staged row coefficients once
// per tile, so the per-entry work is only the "hat" triple and three fused dot products:
//   entry = c0 phi + ghat (cg1 dx + cg2 dy + cg34) + hhat (ch3 dx^2 + ch4 dy^2)
// Rows of the form c0 phi + c3 (phi_xx + phi_yy) (JET_ISO) use the radial Laplacian
//   lap = 2 g + h s;   polyharmonic: lap = p^2 r^(p-2), p^2 folded into the coefficient.
template <int KIND>
__device__ __forceinline__ void hat_constants(int ip, double e2, double &gamma, double &eta, double &lapfac) {
  if (KIND == UPDES_RBF_POLYHARMONIC) {
    const int p = 2 * ip + 1;
    gamma = (double)p; eta = (double)(p * (p - 2)); lapfac = (double)(p * p);
  } else if (KIND == UPDES_RBF_GAUSSIAN) {
    gamma = -2.0 * e2; eta = 4.0 * e2 * e2; lapfac = 1.0;
  } else if (KIND == UPDES_RBF_MULTIQUADRIC) {
    gamma = e2; eta = -(e2 * e2); lapfac = 1.0;
  } else if (KIND == UPDES_RBF_INVERSE_MULTIQUADRIC) {
    gamma = -e2; eta = 3.0 * (e2 * e2); lapfac = 1.0;
  } else {
    gamma = 1.0; eta = 1.0; lapfac = 1.0;
  }
}

// (phi, ghat, hhat) for s > 0 (s not tiny)
// PFIX: compile-time polyharmonic exponent (3 = the default r^3 kernel), 0 = runtime `ip`
template <int KIND, int PFIX>
__device__ __forceinline__ void radial_hat(double s, int ip, double e2, double &phi, double &gh, double &hh) {
  if (KIND == UPDES_RBF_POLYHARMONIC) {
    const int p = PFIX ? PFIX : 2 * ip + 1;
    const double rs = PFIX == 3 ? fast_rsqrt_clamped(s) : fast_rsqrt_pos(s);
    if (p == 3) hh = rs;
    else if (p == 1) hh = rs * rs * rs;
    else hh = ipow_u(s, (p - 5) >> 1) * (s * rs);
    gh = hh * s;
    phi = gh * s;
  } else if (KIND == UPDES_RBF_GAUSSIAN) {
    phi = exp(-e2 * s); gh = phi; hh = phi;
  } else if (KIND == UPDES_RBF_MULTIQUADRIC) {
    const double w = fma(e2, s, 1.0);
    const double rw = fast_rsqrt_pos(w);
    phi = w * rw; gh = rw; hh = rw * rw * rw;
  } else if (KIND == UPDES_RBF_INVERSE_MULTIQUADRIC) {
    const double w = fma(e2, s, 1.0);
    const double rw = fast_rsqrt_pos(w);
    const double rw2 = rw * rw;
    phi = rw; gh = rw2 * rw; hh = gh * rw2;
  } else {
    radial<KIND>(s, ip, e2, phi, gh, hh);      // thin plate: no constant to fold
  }
}

// One evaluation point of one row, coefficients already scaled by gamma / eta / lapfac.
struct alignas(16) RowPoint {
  double x, y;
  double ch3, c0;                           // JET_ISO rows: ch3 holds lapfac * c3 (one 16-byte load with c0)
  double cg1, cg2, cg34, ch4;
};

template <int KIND, int MASK, int PFIX>
__device__ __forceinline__ double entry_one_point(const RowPoint &rp, double cx, double cy, int ip, double e2,
                                                  double gamma2, double eta) {
  const double dx = rp.x - cx, dy = rp.y - cy;
  const double dy2 = dy * dy;
  if (MASK & JET_ISO) {
    const double s = fma(dx, dx, dy2);
    double v;
    if (KIND == UPDES_RBF_POLYHARMONIC) {
      const int p = PFIX ? PFIX : 2 * ip + 1;
      const double rs = PFIX == 3 ? fast_rsqrt_clamped(s) : fast_rsqrt_pos(s);
      double l;                                 // r^(p-2)
      if (p == 3) l = s * rs;
      else if (p == 1) l = rs;
      else l = ipow_u(s, (p - 3) >> 1) * (s * rs);
      v = rp.ch3 * l;
      if (MASK & JET_VAL) v = fma(rp.c0 * s, l, v);
    } else {
      double phi, gh, hh;
      radial_hat<KIND, PFIX>(s, ip, e2, phi, gh, hh);
      v = rp.ch3 * fma(eta * hh, s, gamma2 * gh);   // lap = 2 g + h s
      if (MASK & JET_VAL) v = fma(rp.c0, phi, v);
    }
    // r == 0: derivatives -> 0 (nan_to_num), value -> phi(0).  r^3: already exactly 0 (clamped seed).
    if (KIND == UPDES_RBF_POLYHARMONIC && PFIX == 3) return v;
    return tiny_or_zero(s) ? ((MASK & JET_VAL) ? rp.c0 * phi_at_zero<KIND>() : 0.0) : v;
  }
  const double dx2 = dx * dx;
  const double s = dx2 + dy2;
  double phi, gh, hh;
  radial_hat<KIND, PFIX>(s, ip, e2, phi, gh, hh);
  double v = 0.0;
  if (MASK & JET_H) v = hh * fma(rp.ch3, dx2, rp.ch4 * dy2);
  if (MASK & JET_G) v = fma(gh, fma(rp.cg1, dx, fma(rp.cg2, dy, rp.cg34)), v);
  if (MASK & JET_VAL) v = fma(rp.c0, phi, v);
  if (KIND == UPDES_RBF_POLYHARMONIC && PFIX == 3) return v;    // every term is exactly 0 at s == 0
  return tiny_or_zero(s) ? ((MASK & JET_VAL) ? rp.c0 * phi_at_zero<KIND>() : 0.0) : v;
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

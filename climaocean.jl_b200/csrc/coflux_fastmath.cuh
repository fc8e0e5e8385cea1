// coflux_fastmath.cuh — lean Float64 elementary functions for the similarity hot loop (sm_100a).
//
// The converged Float64 solve is bound by the FP64 pipe, so the lever is FP64 instructions
// per pass.  The CUDA math library spends 30–50 FP64 instructions plus range/special-case branches per
// log/exp/cbrt call and ~15 per division; the functions below spend 8–12 and have no branches.  They are NOT
// general-purpose: arguments must be positive, finite, normal numbers far from the overflow/underflow
// thresholds (the callers guarantee it and fall back to the exact path otherwise).  Accuracy is at rounding
// level — measured against 50-digit references in tests/test_fastmath.py (host build of this same header)
// and against the CUDA math library on the device by tools/fm_check.cu:
//     rcp, div, sqrt ≤ 2 ulp;  cbrt ≤ 2 ulp;  exp ≤ 2 ulp (relative);  log ≤ 2 ulp of max(1, |ln x|) (absolute).
// Seeds come from the SFU (MUFU.RCP64H / MUFU.RSQ64H / MUFU.LG2 / MUFU.EX2, all off the FP64 pipe) and are
// refined with FP64 FMAs; log and exp use small tables (2 KB + 512 B, copied to shared memory by the kernel).
//
// The header also compiles as plain C++ (seeds emulated at 20-bit accuracy) so that the CPU test-suite can
// pin the polynomials and tables without a GPU.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>
#include "coflux_math_tables.h"

#ifdef __CUDACC__
#define COFLUX_FM __device__ __forceinline__
#else
#define COFLUX_FM static inline
#endif

namespace coflux {
namespace fm {

// Literals of the hot loop.  On the device they live in constant memory so that FP64 instructions take them as
// c[bank][offset] operands: a 64-bit immediate would cost two extra move instructions per use, and the loop
// is issue-bound as much as FP64-bound (profiles/).
struct Consts {
  double third, ln2, l6, l5, l4, l3, l2;            // cbrt, log
  double magic, k64ln2, ln2_64_hi, ln2_64_lo;      // exp argument reduction
  double e5, e4, e3, e2;                           // exp polynomial
};
#define COFLUX_FM_CONSTS { 0.33333333333333333, 0.6931471805599453, -1.0 / 6.0, 0.2, -0.25, 1.0 / 3.0, -0.5, \
                           6755399441055744.0, 92.33248261689366, -0.010830424696249145, -3.623510646634843e-19, \
                           1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0, 0.5 }
#ifdef __CUDACC__
__constant__ Consts KC = COFLUX_FM_CONSTS;
#else
static const Consts KC = COFLUX_FM_CONSTS;
#endif

COFLUX_FM double fma_(double a, double b, double c) { return ::fma(a, b, c); }

COFLUX_FM int64_t bits_of(double x) {
#ifdef __CUDA_ARCH__
  return __double_as_longlong(x);
#else
  int64_t b; memcpy(&b, &x, 8); return b;
#endif
}
COFLUX_FM double from_bits(int64_t b) {
#ifdef __CUDA_ARCH__
  return __longlong_as_double(b);
#else
  double x; memcpy(&x, &b, 8); return x;
#endif
}

#ifndef __CUDA_ARCH__
// host emulation of an SFU seed: keep `nbits` mantissa bits of the exact value
COFLUX_FM double coarse(double v, int nbits = 20) {
  int64_t b = bits_of(v);
  b &= ~((int64_t(1) << (52 - nbits)) - 1);
  return from_bits(b);
}
#endif

// ---- seeds ------------------------------------------------------------------------------------------
COFLUX_FM double rcp_seed(double x) {
#ifdef __CUDA_ARCH__
  double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); return r;
#else
  return coarse(1.0 / x);
#endif
}
COFLUX_FM double rsqrt_seed(double x) {
#ifdef __CUDA_ARCH__
  double r; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); return r;
#else
  return coarse(1.0 / ::sqrt(x));
#endif
}

// ---- 1/x, a/b, √x -----------------------------------------------------------------------------------
COFLUX_FM double rcp(double b) {
  double r = rcp_seed(b);
  double e = fma_(-b, r, 1.0);
  r = fma_(r, e, r);
  e = fma_(-b, r, 1.0);
  return fma_(r, e, r);
}
COFLUX_FM double div(double a, double b) {
  const double r = rcp(b);
  const double q = a * r;
  return fma_(fma_(-b, q, a), r, q);
}
COFLUX_FM double sqrt(double x) {      // x > 0
  const double y = rsqrt_seed(x);
  double g = x * y, h = 0.5 * y;
  double r = fma_(-h, g, 0.5);
  g = fma_(g, r, g); h = fma_(h, r, h);
  r = fma_(-h, g, 0.5);
  g = fma_(g, r, g); h = fma_(h, r, h);
  return fma_(fma_(-g, g, x), h, g);   // final residual correction
}

// ---- ∛x, x > 0 ---------------------------------------------------------------------------------------
COFLUX_FM double cbrt(double x) {
#ifdef __CUDA_ARCH__
  const float xf = __double2float_rn(x);
  float lg; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(xf));
  float cf; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(cf) : "f"(lg * 0.333333343f));
  float rf; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rf) : "f"(cf));
  double c = (double)cf, r = (double)rf;
#else
  double c = coarse(::cbrt(x), 19), r = coarse(1.0 / c, 19);
#endif
  const double k = (r * r) * KC.third;   // ≈ 1/(3c²); only scales the Newton correction
  double e = fma_(c * c, c, -x);                    // seed 2⁻²⁰ → 2⁻⁴⁰ → 2⁻⁶⁰ (k is only 2⁻²⁰-accurate, hence not 2⁻⁸⁰)
  c = fma_(-e, k, c);
  e = fma_(c * c, c, -x);
  return fma_(-e, k, c);
}

// ---- ln x, x > 0 normal ------------------------------------------------------------------------------
// tab: COFLUX_LOG_TABLE (shared-memory copy on the device)
COFLUX_FM double log(double x, const double* tab) {
  const int64_t b = bits_of(x);
  const int hi = (int)(b >> 32);
  const int e = (hi >> 20) - 1023;
  const int j = (hi >> 13) & 127;
  const double m = from_bits((b & 0x000fffffffffffffLL) | 0x3ff0000000000000LL);
#ifdef __CUDA_ARCH__
  const double2 t = *reinterpret_cast<const double2*>(tab + 2 * j);
  const double rj = t.x, lj = t.y;
#else
  const double rj = tab[2 * j], lj = tab[2 * j + 1];
#endif
  const double r = fma_(m, rj, -1.0);                 // |r| ≤ 2⁻⁸
  // −1/6 and 1/5 rounded to 21 bits (32-bit immediates; they multiply r⁶, r⁵ ≤ 2⁻⁴⁰, so 2⁻²² relative is plenty)
  double q = fma_(r, -0.16666662693023682, 0.20000004768371582);
  q = fma_(q, r, -0.25);
  q = fma_(q, r, KC.l3);
  q = fma_(q, r, -0.5);
  const double p = fma_(q, r * r, r);                 // log1p(r), truncation r⁷/7 < 2e-18
  return fma_((double)e, KC.ln2, lj) + p;
}

// ---- eˣ, |x| < 700 -----------------------------------------------------------------------------------
// tab: COFLUX_EXP_TABLE (shared-memory copy on the device)
COFLUX_FM double exp(double x, const double* tab) {
  const double t = fma_(x, KC.k64ln2, KC.magic);      // magic = 1.5·2⁵²: the integer n = rint(64x/ln2) lands in the low mantissa bits
  const int n = (int)(uint32_t)bits_of(t);
  const double nf = t - KC.magic;
  double r = fma_(nf, KC.ln2_64_hi, x);               // −double(ln2/64): the fma rounds the exact x − n·hi once
  r = fma_(nf, KC.ln2_64_lo, r);                      // −(ln2/64 − double(ln2/64))
  double q = fma_(r, 0.00833333283662796, KC.e4);     // 1/120 rounded to 21 bits (multiplies r⁵ < 2⁻³⁷)
  q = fma_(q, r, KC.e3);
  q = fma_(q, r, 0.5);
  q = fma_(q, r, 1.0);
  const double p = q * r;                             // e^r − 1, truncation r⁶/720 < 4e-17
  const double T = from_bits(bits_of(tab[n & 63]) + ((int64_t)(n >> 6) << 52));
  return fma_(T, p, T);
}

// ---- Float32 overloads: the CUDA single-precision functions (already SFU-based and short), so that the same pass
// template serves both precisions -----------------------------------------------------------------------
COFLUX_FM float fma_(float a, float b, float c) { return ::fmaf(a, b, c); }
// COFLUX_F32_FAST (see below): 1/x and √x of the Float32 PASS from the SFU — rcp.approx + one Newton step (≤ 1 ulp; the
// IEEE __frcp_rn is 8 instructions with a slow-path branch and was 6 % of the Float32 kernel's instructions, whose issue
// slots are 82 % busy) and sqrt.approx (≈ 1 ulp; sqrtf: 10 instructions).  The phase-A thermodynamics do not come through
// here (M<float>: IEEE division, powf, expf), because Δq amplifies their errors a hundredfold.
#ifndef COFLUX_F32_FAST
#define COFLUX_F32_FAST 1
#endif
COFLUX_FM float rcp(float x) {
#if defined(__CUDA_ARCH__) && COFLUX_F32_FAST
  float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return ::fmaf(r, ::fmaf(-x, r, 1.0f), r);
#elif defined(__CUDA_ARCH__)
  return __frcp_rn(x);
#else
  return 1.0f / x;
#endif
}
COFLUX_FM float div(float a, float b) { return a / b; }
COFLUX_FM float sqrt(float x) {
#if defined(__CUDA_ARCH__) && COFLUX_F32_FAST
  float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#else
  return ::sqrtf(x);
#endif
}
// COFLUX_F32_FAST: SFU forms of the three transcendental functions of the Float32 pass — lg2.approx / ex2.approx (MUFU.LG2 /
// MUFU.EX2) with one Newton step for the cube root — instead of the CUDA library's logf / expf / cbrtf (20–30 instructions
// each).  Accuracy: ≈ 2⁻²¹ absolute in log₂, i.e. ≤ 3e-7 relative in ln(h/ℓ) ≈ 10.  Measured on B200 (round 2): 2.35 → 2.15 ms at
// 1/12°, and the CUDA-vs-oracle error distribution is unchanged (max|d|/max|b| 3.3e-7, p99 5.4e-6: profiles/README.md).
COFLUX_FM float cbrt(float x) {
#if defined(__CUDA_ARCH__) && COFLUX_F32_FAST
  float lg; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(x));
  float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(lg * 0.333333343f));
  float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(y * y));
  return ::fmaf(-(::fmaf(y * y, y, -x)) * 0.333333343f, r, y);      // one Newton step: relative error ≈ 2⁻⁴⁰ → rounding
#else
  return ::cbrtf(x);
#endif
}
COFLUX_FM float log(float x, const double*) {
#if defined(__CUDA_ARCH__) && COFLUX_F32_FAST
  return __logf(x);
#else
  return ::logf(x);
#endif
}
COFLUX_FM float exp(float x, const double*) {
#if defined(__CUDA_ARCH__) && COFLUX_F32_FAST
  return __expf(x);
#else
  return ::expf(x);
#endif
}

}  // namespace fm
}  // namespace coflux

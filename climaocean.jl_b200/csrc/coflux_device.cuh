// coflux_device.cuh — device-side physics of the surface-flux path (sm_100a).
//
// One thread owns one surface cell; the whole similarity iteration lives in registers.
// This is the product implementation (not shared with oracle/): compared with the literal
// restatement it hoists everything that is invariant under the fixed-point iteration
// (atmosphere/surface thermodynamic states, surface humidity, Δθ, Δq, ν(T_s), ln-free
// constants), evaluates only the taken branch of each stability function, and evaluates the
// scalar similarity profile once when the θ and q roughness parameterisations coincide.
// None of these change a rounded value that the reference formulas (SURVEY.md Appendix A;
// parameter surface /root/reference/src/OMIPConfigurations/omip_simulation.jl:40-113) define.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>
#include "../../include/coflux.h"
#include "coflux_fastmath.cuh"
#include "coflux_psi_table.h"

namespace coflux {

// ---------------------------------------------------------------------------------------------
// libm dispatch (IEEE-accurate CUDA math library entry points; no fast-math anywhere).
// The transcendental entry points are deliberately NOT inlined: with every call site inlined the
// similarity loop was 60+ KB of SASS and the kernel stalled on instruction fetch (ncu: stall
// 'no_instruction' 4.6 per issue, profiles/r01_flux_tile_v2a_*); one shared copy of each keeps the
// loop inside the instruction cache.
// ---------------------------------------------------------------------------------------------
#ifndef COFLUX_MATH_INLINE_MASK
#define COFLUX_MATH_INLINE_MASK 0      /* bit0 log, bit1 exp, bit2 cbrt, bit3 atan, bit4 pow: 1 = force inline (tuning) */
#endif
#if COFLUX_MATH_INLINE_MASK & 1
#define COFLUX_INL_LOG __forceinline__
#else
#define COFLUX_INL_LOG __noinline__
#endif
#if COFLUX_MATH_INLINE_MASK & 2
#define COFLUX_INL_EXP __forceinline__
#else
#define COFLUX_INL_EXP __noinline__
#endif
#if COFLUX_MATH_INLINE_MASK & 4
#define COFLUX_INL_CBRT __forceinline__
#else
#define COFLUX_INL_CBRT __noinline__
#endif
#if COFLUX_MATH_INLINE_MASK & 8
#define COFLUX_INL_ATAN __forceinline__
#else
#define COFLUX_INL_ATAN __noinline__
#endif
#if COFLUX_MATH_INLINE_MASK & 16
#define COFLUX_INL_POW __forceinline__
#else
#define COFLUX_INL_POW __noinline__
#endif
template <typename FT> struct M;
template <> struct M<double> {
  static __device__ COFLUX_INL_LOG double log(double x) { return ::log(x); }
  static __device__ COFLUX_INL_EXP double exp(double x) { return ::exp(x); }
  static __device__ __forceinline__ double sqrt(double x) { return ::sqrt(x); }
  static __device__ COFLUX_INL_CBRT double cbrt(double x) { return ::cbrt(x); }
  static __device__ COFLUX_INL_ATAN double atan(double x) { return ::atan(x); }
  static __device__ COFLUX_INL_POW double pow(double x, double y) { return ::pow(x, y); }
  static __device__ __forceinline__ double div(double a, double b) { return a / b; }
  static __device__ __forceinline__ double abs(double x) { return ::fabs(x); }
  static __device__ __forceinline__ double floor(double x) { return ::floor(x); }
  static __device__ __forceinline__ double trunc(double x) { return ::trunc(x); }
  static __device__ __forceinline__ double min(double a, double b) { return ::fmin(a, b); }
  static __device__ __forceinline__ double max(double a, double b) { return ::fmax(a, b); }
  static __device__ __forceinline__ double inf() { return CUDART_INF; }
  static __device__ __forceinline__ double pi() { return 3.14159265358979323846; }
};
template <> struct M<float> {
  static __device__ COFLUX_INL_LOG float log(float x) { return ::logf(x); }
  static __device__ COFLUX_INL_EXP float exp(float x) { return ::expf(x); }
  static __device__ __forceinline__ float sqrt(float x) { return ::sqrtf(x); }
  static __device__ COFLUX_INL_CBRT float cbrt(float x) { return ::cbrtf(x); }
  static __device__ COFLUX_INL_ATAN float atan(float x) { return ::atanf(x); }
  static __device__ COFLUX_INL_POW float pow(float x, float y) { return ::powf(x, y); }
  static __device__ __forceinline__ float div(float a, float b) { return a / b; }
  static __device__ __forceinline__ float abs(float x) { return ::fabsf(x); }
  static __device__ __forceinline__ float floor(float x) { return ::floorf(x); }
  static __device__ __forceinline__ float trunc(float x) { return ::truncf(x); }
  static __device__ __forceinline__ float min(float a, float b) { return ::fminf(a, b); }
  static __device__ __forceinline__ float max(float a, float b) { return ::fmaxf(a, b); }
  static __device__ __forceinline__ float inf() { return CUDART_INF_F; }
  static __device__ __forceinline__ float pi() { return 3.14159265358979323846f; }
};

// Lean Float64 math policy: pow, exp, division and square root from coflux_fastmath.cuh (tables read from global
// memory), everything else — including the ORDER of operations of the callers — unchanged.  Measured on B200 against the
// CUDA math library / IEEE operations on 4.2 M arguments each (tools/fm_check.cu): div and sqrt agree bit for bit, exp
// and log (hence pow) to ≤ 1 ulp.  Against the oracle on 256×128 cells q★ deviates 1.6e-13 (relative, floor 1e-3) with this
// policy; a fused single-exponential form of p_sat that is mathematically identical deviates 5.2e-13 and pushes the
// salt flux past the 1e-12 bar — hence the policy swaps functions, never formulas.
struct MLeanD {
  static __device__ __forceinline__ double pow(double x, double y) { return fm::exp(y * fm::log(x, &COFLUX_LOG_TABLE[0][0]), COFLUX_EXP_TABLE); }
  static __device__ __forceinline__ double exp(double x) { return fm::exp(x, COFLUX_EXP_TABLE); }
  static __device__ __forceinline__ double log(double x) { return fm::log(x, &COFLUX_LOG_TABLE[0][0]); }
  static __device__ __forceinline__ double div(double a, double b) { return fm::div(a, b); }
  static __device__ __forceinline__ double sqrt(double x) { return (x > 0.0) ? fm::sqrt(x) : ::sqrt(x); }
};
#ifndef COFLUX_LEAN_V1
#define COFLUX_LEAN_V1 1      /* one-cell-per-thread kernels (Large–Yeager, sea ice): lean policy in Float64 */
#endif
template <typename FT> struct DefaultMP { using type = M<FT>; };
#if COFLUX_LEAN_V1
template <> struct DefaultMP<double> { using type = MLeanD; };
#endif

// ---------------------------------------------------------------------------------------------
// Device parameter block (built once per context on the host, in FT arithmetic)
// ---------------------------------------------------------------------------------------------
template <typename FT> struct ThermoC {
  FT R_d, R_v, eps, cp_d, cp_v, cp_l, cp_i, LH_v0, LH_s0, T_0, T_tr, p_tr, T_fr, T_in;
  FT Rd_over_Rv;
  // constants of the Clausius–Clapeyron form for the pure phases, divided once on the host exactly as psat_generic divides
  // them per call (same operands, same roundings, IEEE division): Δcp/R_v, (LH_0 − Δcp·T_0)/R_v, 1/T_tr
  FT a_liq, b_liq, a_ice, b_ice, inv_T_tr;
};
template <typename FT> struct Visc { int kind; FT nu, c0, c1, c2, c3; };
template <typename FT> struct MomRough {
  int kind, waves;
  FT fixed, alpha, a1, a2, umax, amin, beta_s, lmax, g;
  Visc<FT> visc;
};
template <typename FT> struct ScaRough { int kind; FT fixed, A, b, lmax; Visc<FT> visc; };
template <typename FT> struct FluxP {
  int formulation, stability, form, velocity, stop_kind, maxit, itemp, same_scalar, same_visc, skin_update;
  FT tol, kappa, beta, ugmin, init, ly_umin, skin_max_dT;
  MomRough<FT> mr;
  ScaRough<FT> tr, qr;
};
template <typename FT> struct IceOceanP {
  int heat_flux, friction;
  FT um_star, T0, slope, alpha_h, alpha_s, ustar_const, ustar_min, rho_i, L_f, Cd, k_ice, h_c;
};
// constants of the fast iteration (coflux_solve_tile.cuh), built on the host
template <typename FT> struct FastConsts {
  FT lnh;
  FT lnhA_q, lnhl_q, lrclip_q;     // ln(h/A), ln(h/ℓmax or fixed), ln(A/ℓmax)/b  (water vapour)
  FT lnhA_t, lnhl_t, lrclip_t;     // same for temperature
  int edson, gust_skip, fast_q, fast_t;
  // reciprocals / products hoisted for the lean Float64 pass (coflux_solve_tile.cuh::iterate_lean)
  FT alpha_g, inv_g, bnu, inv_nu;   // bnu, inv_nu: constant-viscosity case
};
// CCSM3 sea-ice albedo (coflux_ccsm3_albedo; atmosphere.jl:31-44 `SeaIceAlbedo(hi, hs, Ts)`)
template <typename FT> struct Ccsm3P {
  FT ice_v, ice_n, snow_v, snow_n, hmax, dT, d_ice, d_snow_v, d_snow_n, hpatch, alb_o, fvis, Tmelt;
};
template <typename FT> struct DevParams {
  ThermoC<FT> th;
  FastConsts<FT> K;                // for the atmosphere–ocean parameters
  FT h, hbl, g;
  FT rho0, c0, rhof, Smin, wmf_alpha, T_offset;  // T_offset: 273.15 if ocean T in °C else 0
  FT rho0inv, rhofinv;                           // 1/ρ₀, 1/ρ_f: divided once on the host (IEEE, the bits a per-cell division gives)
  FT sigma, alb_o, emis_o, emis_i, alb_i;
  int sw_pen, ice_albedo_kind;
  Ccsm3P<FT> ccsm3;
  FluxP<FT> ao, ai;
  IceOceanP<FT> io;
};

// ---------------------------------------------------------------------------------------------
// Sea-ice albedo seen by the radiation terms of rows a7 / f1: a prescribed plane or constant, or the CCSM3 form evaluated
// from the live ice thickness, snow thickness and top temperature (include/coflux.h: coflux_ccsm3_albedo)
// ---------------------------------------------------------------------------------------------
template <typename FT> __device__ __forceinline__ FT ccsm3_albedo(const Ccsm3P<FT>& c, FT hi, FT hs, FT TsK) {
  const FT fh = M<FT>::min(M<FT>::atan(FT(4) * hi) / M<FT>::atan(FT(4) * c.hmax), FT(1));
  const FT fT = M<FT>::min(M<FT>::max(FT(1) - (c.Tmelt - TsK) / c.dT, FT(0)), FT(1));
  const FT aiv = c.ice_v * fh + c.alb_o * (FT(1) - fh) - c.d_ice * fT;
  const FT ain = c.ice_n * fh + c.alb_o * (FT(1) - fh) - c.d_ice * fT;
  const FT asv = c.snow_v - c.d_snow_v * fT;
  const FT asn = c.snow_n - c.d_snow_n * fT;
  const FT fs = hs / (hs + c.hpatch);
  const FT av = (FT(1) - fs) * aiv + fs * asv;
  const FT an = (FT(1) - fs) * ain + fs * asn;
  return c.fvis * av + (FT(1) - c.fvis) * an;
}

// ---------------------------------------------------------------------------------------------
// Thermodynamics (A1, A2)
// ---------------------------------------------------------------------------------------------
template <typename FT> struct Thermo { FT rho, cp_m, q_vap, T_v; };

// MP: the math policy supplying pow and exp (M<FT>: CUDA math library; the Float64 tile kernel passes a policy
// built on coflux_fastmath.cuh).  The ORDER of operations is the reference's in either case: 1/T_tr − 1/T cancels
// three digits, so any reformulation of this expression moves q_sat by ~10 ulp and Δq = q_a − q_s by 100× that.
template <typename FT, class MP = M<FT>>
__device__ __forceinline__ FT psat_generic(const ThermoC<FT>& c, FT T, FT LH_0, FT dcp) {
  return c.p_tr * MP::pow(MP::div(T, c.T_tr), MP::div(dcp, c.R_v)) *
         MP::exp(MP::div(LH_0 - dcp * c.T_0, c.R_v) * (c.inv_T_tr - MP::div(FT(1), T)));
}
// the same for a pure phase, with the two quotients of constants taken from the parameter block (a = Δcp/R_v,
// b = (LH_0 − Δcp·T_0)/R_v): the same numbers psat_generic would form, three divisions less per call
template <typename FT, class MP = M<FT>>
__device__ __forceinline__ FT psat_pure(const ThermoC<FT>& c, FT T, FT a, FT b) {
  return c.p_tr * MP::pow(MP::div(T, c.T_tr), a) * MP::exp(b * (c.inv_T_tr - MP::div(FT(1), T)));
}
template <typename FT, class MP = M<FT>> __device__ __forceinline__ FT liquid_fraction(const ThermoC<FT>& c, FT T) {
  if (T > c.T_fr) return FT(1);
  if (T <= c.T_in) return FT(0);
  return MP::div(T - c.T_in, c.T_fr - c.T_in);
}
// MLeanD forms pow(x, y) as exp(y·log x): the two saturation pressures of one temperature (pure phase at the surface, mixed
// phase inside phase_equil_pTq) then share ln(T/T_tr) and 1/T_tr − 1/T — computed once by the caller, same operations,
// same bits, one logarithm and two divisions less.  Other policies (powf) keep their own pow.
template <class MP> struct PowIsExpLog { static constexpr bool value = false; };
template <> struct PowIsExpLog<MLeanD> { static constexpr bool value = true; };
template <typename FT> struct PsatT { FT lnT, rT; };
// ps_liquid (optional): the saturation pressure over LIQUID water at this T, if the caller has it.  For T above freezing
// the liquid fraction is exactly 1, LH_0 and Δcp below are then exactly the liquid constants, and the saturation pressure
// computed here would be the same number, bit for bit — so it is reused instead of recomputed (a pow and an exp).
template <typename FT, class MP = M<FT>> __device__ __forceinline__ Thermo<FT> phase_equil_pTq(const ThermoC<FT>& c, FT p, FT T, FT q,
                                                                                               const FT* ps_liquid = nullptr,
                                                                                               const PsatT<FT>* tt = nullptr) {
  FT lam = liquid_fraction<FT, MP>(c, T);
  FT LH_0 = lam * c.LH_v0 + (FT(1) - lam) * c.LH_s0;
  FT dcp = lam * (c.cp_v - c.cp_l) + (FT(1) - lam) * (c.cp_v - c.cp_i);
  FT ps;
  if (ps_liquid && lam == FT(1)) ps = *ps_liquid;
  else if (PowIsExpLog<MP>::value && tt) ps = c.p_tr * MP::exp(MP::div(dcp, c.R_v) * tt->lnT) * MP::exp(MP::div(LH_0 - dcp * c.T_0, c.R_v) * tt->rT);
  else ps = psat_generic<FT, MP>(c, T, LH_0, dcp);
  FT denom = p - ps;
  FT q_vs = (denom > FT(0)) ? MP::div(c.Rd_over_Rv * (FT(1) - q) * ps, denom) : M<FT>::inf();
  FT q_c = M<FT>::max(q - q_vs, FT(0));
  FT q_liq = lam * q_c, q_ice = (FT(1) - lam) * q_c;
  FT R_m = c.R_d * (FT(1) + (c.eps - FT(1)) * q - c.eps * q_c);
  Thermo<FT> s;
  s.rho = MP::div(p, R_m * T);
  s.cp_m = c.cp_d + (c.cp_v - c.cp_d) * q + (c.cp_l - c.cp_v) * q_liq + (c.cp_i - c.cp_v) * q_ice;
  s.q_vap = q - q_liq - q_ice;
  s.T_v = MP::div(T * R_m, c.R_d);
  return s;
}

// ---------------------------------------------------------------------------------------------
// Stability functions (A5) — only the taken branch is evaluated
// ---------------------------------------------------------------------------------------------
// LM: the math policy (M<FT>: CUDA math library; MLeanD in Float64: lean log / cbrt / sqrt / division, ≤ 1 ulp apart)
template <typename FT> struct LMath { using P = typename DefaultMP<FT>::type;
  static __device__ __forceinline__ FT log(FT x) { return P::log(x); }
  static __device__ __forceinline__ FT sqrt(FT x) { return P::sqrt(x); }
  static __device__ __forceinline__ FT div(FT a, FT b) { return P::div(a, b); }
  static __device__ __forceinline__ FT cbrt(FT x);
};
template <> __device__ __forceinline__ double LMath<double>::cbrt(double x) {
#if COFLUX_LEAN_V1
  return (x > 1e-30 && x < 1e30) ? fm::cbrt(x) : ::cbrt(x);
#else
  return ::cbrt(x);
#endif
}
template <> __device__ __forceinline__ float LMath<float>::cbrt(float x) { return ::cbrtf(x); }

// Table evaluation of a stability function of the sea-ice / Large–Yeager solves (Paulson, SHEBA: coflux_psi_table.h):
// z ∈ [2^KMIN, 2^ICE_KMAX) positive (−ζ for the unstable table, ζ for the stable one); which = 0 momentum, 1 scalar.
// Degree-7 piecewise polynomials, error ≤ 3e-16·max(1,|ψ|).  The range reaches 2²⁶: calm, strongly stratified ice cells
// iterate through |ζ| of 10⁴ … 10⁷, and the closed forms they used to fall back to (cube root, two arctangents, three
// logarithms, with a handful of lanes active) cost a fifth of the whole solve (measured, round 2).
#define COFLUX_PSI_ROWS ((COFLUX_PSI_ICE_KMAX - COFLUX_PSI_KMIN) * COFLUX_PSI_NS)
__device__ __forceinline__ bool psi_tab_in_range(double z) { return z >= 9.313225746154785e-10 && z < 67108864.0; }
__device__ __forceinline__ bool psi_tab_in_range(float z) { return z >= 9.313225746154785e-10f && z < 67108864.0f; }
static_assert(COFLUX_PSI_KMIN == -30 && COFLUX_PSI_ICE_KMAX == 26, "psi_tab_in_range assumes [2^-30, 2^26)");
__device__ __forceinline__ double psi_tab_eval(const double (*tab)[2][8], double z, int which) {
  const long long bits = __double_as_longlong(z);
  const int hi = (int)(bits >> 32);
  int row = (hi >> 16) - ((1023 + COFLUX_PSI_KMIN) << 4);
  row = max(0, min(row, COFLUX_PSI_ROWS - 1));
  const double t = fma(2.0, __longlong_as_double(((bits & 0x0000ffffffffffffLL) << 4) | 0x3ff0000000000000LL), -3.0);
  // two 256-bit loads (LDG.E.256, sm_100): a lane's polynomial is 64 B of its own table row, so 32 lanes read 32 different
  // rows and the L1 data pipe, not latency, is what these gathers cost — half the requests of four 128-bit loads
  const double* q = &tab[row][which][0];
  double k0, k1, k2, k3, k4, k5, k6, k7;
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(k0), "=d"(k1), "=d"(k2), "=d"(k3) : "l"(q));
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(k4), "=d"(k5), "=d"(k6), "=d"(k7) : "l"(q + 4));
  double a = fma(k7, t, k6);
  a = fma(a, t, k5); a = fma(a, t, k4); a = fma(a, t, k3); a = fma(a, t, k2); a = fma(a, t, k1);
  return fma(a, t, k0);
}
__device__ __forceinline__ float psi_tab_eval(const float (*tab)[2][8], float z, int which) {
  const int bits = __float_as_int(z);
  int row = (bits >> 19) - ((127 + COFLUX_PSI_KMIN) << 4);
  row = max(0, min(row, COFLUX_PSI_ROWS - 1));
  const float t = fmaf(2.0f, __int_as_float(((bits & 0x0007ffff) << 4) | 0x3f800000), -3.0f);
  const float4* q = reinterpret_cast<const float4*>(&tab[row][which][0]);
  const float4 c0 = __ldg(q), c1 = __ldg(q + 1);
  float a = fmaf(c1.w, t, c1.z);
  a = fmaf(a, t, c1.y); a = fmaf(a, t, c1.x); a = fmaf(a, t, c0.w); a = fmaf(a, t, c0.z); a = fmaf(a, t, c0.y);
  return fmaf(a, t, c0.x);
}
// both functions of one row at once (momentum and scalar share the argument, hence the row and t): one index computation,
// the loads of both polynomials in flight together, two independent Horner chains.  Same operations per function as
// psi_tab_eval → same bits.
__device__ __forceinline__ void psi_tab_eval_pair(const double (*tab)[2][8], double z, double& pm, double& ps) {
  const long long bits = __double_as_longlong(z);
  const int hi = (int)(bits >> 32);
  int row = (hi >> 16) - ((1023 + COFLUX_PSI_KMIN) << 4);
  row = max(0, min(row, COFLUX_PSI_ROWS - 1));
  const double t = fma(2.0, __longlong_as_double(((bits & 0x0000ffffffffffffLL) << 4) | 0x3ff0000000000000LL), -3.0);
  const double* q = &tab[row][0][0];
  double m0, m1, m2, m3, m4, m5, m6, m7, s0, s1, s2, s3, s4, s5, s6, s7;
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(m0), "=d"(m1), "=d"(m2), "=d"(m3) : "l"(q));
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(m4), "=d"(m5), "=d"(m6), "=d"(m7) : "l"(q + 4));
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(s0), "=d"(s1), "=d"(s2), "=d"(s3) : "l"(q + 8));
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(s4), "=d"(s5), "=d"(s6), "=d"(s7) : "l"(q + 12));
  double a = fma(m7, t, m6), b = fma(s7, t, s6);
  a = fma(a, t, m5); b = fma(b, t, s5); a = fma(a, t, m4); b = fma(b, t, s4); a = fma(a, t, m3); b = fma(b, t, s3);
  a = fma(a, t, m2); b = fma(b, t, s2); a = fma(a, t, m1); b = fma(b, t, s1);
  pm = fma(a, t, m0); ps = fma(b, t, s0);
}
__device__ __forceinline__ void psi_tab_eval_pair(const float (*tab)[2][8], float z, float& pm, float& ps) {
  pm = psi_tab_eval(tab, z, 0); ps = psi_tab_eval(tab, z, 1);
}
// one function each at TWO arguments (ψ_m(ℓ_u/L★) and ψ_h(ℓ_q/L★) of the logarithmic profile form): both rows requested
// before either polynomial is evaluated, two independent Horner chains.  Same operations per function → same bits.
__device__ __forceinline__ void psi_tab_eval_two(const double (*tab)[2][8], double za, double zb, double& pa, double& pb) {
  const long long ba = __double_as_longlong(za), bb = __double_as_longlong(zb);
  int ra = ((int)(ba >> 32) >> 16) - ((1023 + COFLUX_PSI_KMIN) << 4), rb = ((int)(bb >> 32) >> 16) - ((1023 + COFLUX_PSI_KMIN) << 4);
  ra = max(0, min(ra, COFLUX_PSI_ROWS - 1)); rb = max(0, min(rb, COFLUX_PSI_ROWS - 1));
  const double ta = fma(2.0, __longlong_as_double(((ba & 0x0000ffffffffffffLL) << 4) | 0x3ff0000000000000LL), -3.0);
  const double tb = fma(2.0, __longlong_as_double(((bb & 0x0000ffffffffffffLL) << 4) | 0x3ff0000000000000LL), -3.0);
  const double* qa = &tab[ra][0][0];
  const double* qb = &tab[rb][1][0];
  double m0, m1, m2, m3, m4, m5, m6, m7, s0, s1, s2, s3, s4, s5, s6, s7;
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(m0), "=d"(m1), "=d"(m2), "=d"(m3) : "l"(qa));
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(m4), "=d"(m5), "=d"(m6), "=d"(m7) : "l"(qa + 4));
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(s0), "=d"(s1), "=d"(s2), "=d"(s3) : "l"(qb));
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(s4), "=d"(s5), "=d"(s6), "=d"(s7) : "l"(qb + 4));
  double a = fma(m7, ta, m6), b = fma(s7, tb, s6);
  a = fma(a, ta, m5); b = fma(b, tb, s5); a = fma(a, ta, m4); b = fma(b, tb, s4); a = fma(a, ta, m3); b = fma(b, tb, s3);
  a = fma(a, ta, m2); b = fma(b, tb, s2); a = fma(a, ta, m1); b = fma(b, tb, s1);
  pa = fma(a, ta, m0); pb = fma(b, tb, s0);
}
__device__ __forceinline__ void psi_tab_eval_two(const float (*tab)[2][8], float za, float zb, float& pa, float& pb) {
  pa = psi_tab_eval(tab, za, 0); pb = psi_tab_eval(tab, zb, 1);
}
template <typename FT> struct PsiTabs;
template <> struct PsiTabs<double> {
  static __device__ __forceinline__ const double (*paulson())[2][8] { return COFLUX_PSI_PAULSON_F64; }
  static __device__ __forceinline__ const double (*sheba())[2][8] { return COFLUX_PSI_SHEBA_F64; }
};
template <> struct PsiTabs<float> {
  static __device__ __forceinline__ const float (*paulson())[2][8] { return COFLUX_PSI_PAULSON_F32; }
  static __device__ __forceinline__ const float (*sheba())[2][8] { return COFLUX_PSI_SHEBA_F32; }
};
#ifndef COFLUX_ICE_COARE
#define COFLUX_ICE_COARE 1            /* the compact sea-ice pass also takes the COARE log form (`:corrected`, `:ncar` ice: 13.0 / 12.2 ms
                                         instead of 34 ms at 1/12°); validated on B200 in round 2: all sea-ice parity tests, both
                                         precisions, three parameter sets (gpurun_out/r2_pytest6_icecoare.log → profiles/README.md) */
#endif
#ifndef COFLUX_PSI_TABLES_V1
#define COFLUX_PSI_TABLES_V1 1      /* Paulson / SHEBA ψ from tables in the one-cell-per-thread and refill solves */
#endif

template <typename FT> __device__ __forceinline__ FT psi_conv_cbrt(FT y) {  // convective (cube-root) limb shared by Edson ψu, ψθ
  using L = LMath<FT>;
  const FT rt3 = FT(1.7320508075688772935);
  return FT(1.5) * L::log(L::div(FT(1) + y + y * y, FT(3))) - rt3 * M<FT>::atan(L::div(FT(1) + FT(2) * y, rt3)) + L::div(M<FT>::pi(), rt3);
}
template <typename FT> __device__ __forceinline__ FT psi_businger_momentum(FT x) {  // Kansas limb, x = (1 − γ ζ)^{1/4}
  using L = LMath<FT>;
  return FT(2) * L::log((FT(1) + x) / FT(2)) + L::log((FT(1) + x * x) / FT(2)) - FT(2) * M<FT>::atan(x) + M<FT>::pi() / FT(2);
}
template <typename FT> __device__ __noinline__ FT psi_momentum(int kind, FT z) {
  using L = LMath<FT>;
  if (kind == COFLUX_STABILITY_EDSON) {
    if (z >= FT(0)) {
      FT dz = M<FT>::min(FT(50), FT(0.35) * z);
      return -FT(0.7) * z - FT(0.75) * (z - FT(5) / FT(0.35)) * M<FT>::exp(-dz) - FT(0.75) * FT(5) / FT(0.35);
    }
    FT x = L::sqrt(L::sqrt(FT(1) - FT(15) * z));
    FT psik = psi_businger_momentum(x);
    FT psic = psi_conv_cbrt(L::cbrt(FT(1) - FT(10.15) * z));
    FT f = L::div(z * z, FT(1) + z * z);
    return (FT(1) - f) * psik + f * psic;
  }
  if (kind == COFLUX_STABILITY_NEUTRAL) return FT(0);
  // SHEBA_PAULSON and LARGE_YEAGER share the Paulson unstable limb
  if (z < FT(0)) {
    if (COFLUX_PSI_TABLES_V1 && psi_tab_in_range(-z)) return psi_tab_eval(PsiTabs<FT>::paulson(), -z, 0);
    return psi_businger_momentum(L::sqrt(L::sqrt(FT(1) - FT(16) * z)));
  }
  if (kind == COFLUX_STABILITY_LARGE_YEAGER) return -FT(5) * z;
  if (COFLUX_PSI_TABLES_V1 && psi_tab_in_range(z)) return psi_tab_eval(PsiTabs<FT>::sheba(), z, 0);
  // Grachev et al. (2007) SHEBA, stable.  B = ∛((1 − b_m)/b_m) and the constant arctangent depend only on a_m, b_m:
  // written as literals (the correctly rounded values of the FT expressions) instead of a ∛ and an atan per call.
  const FT a = FT(5), b = FT(5) / FT(6.5);   // a_m, b_m = a_m/6.5
  const FT rt3 = FT(1.7320508075688772935);
  FT x = L::cbrt(FT(1) + z);
  const FT B = (sizeof(FT) == 8) ? FT(0.6694329500821695) : FT(0.6694329380989075);
  const FT atanB = (sizeof(FT) == 8) ? FT(0.8539936329836121) : FT(0.8539936542510986);   // atan((2 − B)/(√3 B))
  FT p1 = L::div(-FT(3) * a * (x - FT(1)), b);
  FT p2 = L::div(a * B, FT(2) * b) *
          (FT(2) * L::log(L::div(x + B, FT(1) + B)) - L::log(L::div(x * x - B * x + B * B, FT(1) - B + B * B)) +
           FT(2) * rt3 * (M<FT>::atan(L::div(FT(2) * x - B, rt3 * B)) - atanB));
  return p1 + p2;
}
template <typename FT> __device__ __noinline__ FT psi_scalar(int kind, FT z) {
  using L = LMath<FT>;
  if (kind == COFLUX_STABILITY_EDSON) {
    if (z >= FT(0)) {
      FT dz = M<FT>::min(FT(50), FT(0.35) * z);
      return -M<FT>::pow(FT(1) + FT(2) / FT(3) * z, FT(1.5)) - FT(2) / FT(3) * (z - FT(14.28)) * M<FT>::exp(-dz) - FT(8.525);
    }
    FT x = L::sqrt(FT(1) - FT(15) * z);
    FT psik = FT(2) * L::log((FT(1) + x) / FT(2));
    FT psic = psi_conv_cbrt(L::cbrt(FT(1) - FT(34.15) * z));
    FT f = L::div(z * z, FT(1) + z * z);
    return (FT(1) - f) * psik + f * psic;
  }
  if (kind == COFLUX_STABILITY_NEUTRAL) return FT(0);
  if (z < FT(0)) {
    if (COFLUX_PSI_TABLES_V1 && psi_tab_in_range(-z)) return psi_tab_eval(PsiTabs<FT>::paulson(), -z, 1);
    return FT(2) * L::log((FT(1) + L::sqrt(FT(1) - FT(16) * z)) / FT(2));
  }
  if (kind == COFLUX_STABILITY_LARGE_YEAGER) return -FT(5) * z;
  if (COFLUX_PSI_TABLES_V1 && psi_tab_in_range(z)) return psi_tab_eval(PsiTabs<FT>::sheba(), z, 1);
  const FT a = FT(5), b = FT(5), c = FT(3);
  const FT B = (sizeof(FT) == 8) ? FT(2.23606797749979) : FT(2.2360680103302);              // √(c² − 4)
  const FT logB = (sizeof(FT) == 8) ? FT(-1.9248473002384139) : FT(-1.9248473644256592);    // ln((c − B)/(c + B))
  FT p1 = -b / FT(2) * L::log(FT(1) + c * z + z * z);
  FT p2 = (L::div(-a, B) + L::div(b * c, FT(2) * B)) *
          (L::log(L::div(FT(2) * z + c - B, FT(2) * z + c + B)) - logB);
  return p1 + p2;
}

// ---------------------------------------------------------------------------------------------
// Roughness lengths (A6)
// ---------------------------------------------------------------------------------------------
template <typename FT> __device__ __forceinline__ FT air_viscosity(const Visc<FT>& v, FT T) {
  if (v.kind == COFLUX_VISCOSITY_CONSTANT) return v.nu;
  FT Tp = T - FT(273.15);
  return v.c0 + v.c1 * Tp + v.c2 * Tp * Tp + v.c3 * Tp * Tp * Tp;
}
template <typename FT>
__device__ __forceinline__ FT momentum_roughness(const MomRough<FT>& r, FT ustar, FT U, FT nu) {
  if (r.kind == COFLUX_ROUGHNESS_FIXED) return r.fixed;
  FT alpha = r.alpha;
  if (r.waves == COFLUX_WAVES_WIND_DEPENDENT) alpha = M<FT>::max(r.a1 * M<FT>::min(U, r.umax) + r.a2, r.amin);
  FT lR = (ustar == FT(0)) ? r.lmax : r.beta_s * nu / ustar;
  return M<FT>::min(alpha * ustar * ustar / r.g + lR, r.lmax);
}
template <typename FT>
__device__ __forceinline__ FT scalar_roughness(const ScaRough<FT>& r, FT lu, FT ustar, FT nu) {
  if (r.kind == COFLUX_ROUGHNESS_FIXED) return r.fixed;
  FT Rstar = lu * ustar / nu;
  FT lq = (Rstar == FT(0)) ? FT(0) : r.A / M<FT>::pow(Rstar, r.b);
  return M<FT>::min(lq, r.lmax);
}

// χ = ln(h/ℓ) − ψ(h/L) [+ ψ(ℓ/L)], with ψ(h/L) supplied (it is shared between θ and q)
template <typename FT, bool SCALAR>
__device__ __forceinline__ FT similarity_profile(int form, int stab, FT h, FT l, FT L, FT psi_h) {
  FT chi = M<FT>::log(h / l) - psi_h;
  if (form == COFLUX_PROFILE_LOGARITHMIC) {
    FT th = l / L;
    chi += SCALAR ? psi_scalar(stab, th) : psi_momentum(stab, th);
  }
  return chi;
}

template <typename FT, class MP = M<FT>> __device__ __forceinline__ FT ly_cdn(FT U) {
  if (U >= FT(33)) return FT(2.34e-3);
  FT U2 = U * U, U6 = U2 * U2 * U2;
  return FT(1e-3) * (MP::div(FT(2.7), U) + FT(0.142) + MP::div(U, FT(13.09)) - FT(3.14807e-10) * U6);
}

// ---------------------------------------------------------------------------------------------
// Per-cell interface solve (A3, A4, A7; a7 with SKIN temperature)
// ---------------------------------------------------------------------------------------------
template <typename FT> struct CellIn {
  FT ua, va, Ta, pa, qa, Qs, Ql;
  FT us, vs;      // surface (ocean or ice) velocity at the cell centre
  FT Ts0;         // initial interface temperature [K] (ocean: bulk T; ice: previous T_top)
  FT So;          // ocean salinity (Raoult factor); unused over ice
  FT h_ice, S_ice, albedo;
};
template <typename FT> struct CellOut {
  FT ustar, tstar, qstar, Ts;
  FT rho_a, cp_a, du, dv;
  int it;
};

template <typename FT> struct SurfaceState { FT qs, dq, dtheta, T_v, q_vap, nu_m, nu_t, nu_q; };

template <typename FT, int SURF, class MP = M<FT>>
__device__ __forceinline__ SurfaceState<FT> surface_state(const DevParams<FT>& P, const FluxP<FT>& F,
                                                          const Thermo<FT>& atm, FT pa, FT theta_a, FT x, FT Ts,
                                                          bool need_viscosity = true) {
  const ThermoC<FT>& c = P.th;
  SurfaceState<FT> s;
  FT ps;
  PsatT<FT> tt{FT(0), FT(0)};
  if constexpr (PowIsExpLog<MP>::value) {
    tt.lnT = MP::log(MP::div(Ts, c.T_tr)); tt.rT = c.inv_T_tr - MP::div(FT(1), Ts);
    ps = c.p_tr * MP::exp(((SURF == 0) ? c.a_liq : c.a_ice) * tt.lnT) * MP::exp(((SURF == 0) ? c.b_liq : c.b_ice) * tt.rT);
  } else {
    ps = (SURF == 0) ? psat_pure<FT, MP>(c, Ts, c.a_liq, c.b_liq) : psat_pure<FT, MP>(c, Ts, c.a_ice, c.b_ice);
  }
  FT qstar = MP::div(ps, atm.rho * c.R_v * Ts);
  s.qs = qstar * x;
  s.dq = atm.q_vap - s.qs;
  s.dtheta = theta_a - Ts;
  Thermo<FT> surf = phase_equil_pTq<FT, MP>(c, pa, Ts, s.qs, (SURF == 0) ? &ps : nullptr, PowIsExpLog<MP>::value ? &tt : nullptr);
  s.T_v = surf.T_v;
  s.q_vap = surf.q_vap;
  if (need_viscosity) {            // (fixed roughness lengths never look at it: the sea-ice pass skips the three polynomials)
    s.nu_m = air_viscosity(F.mr.visc, Ts);
    s.nu_t = air_viscosity(F.tr.visc, Ts);
    s.nu_q = air_viscosity(F.qr.visc, Ts);
  } else { s.nu_m = s.nu_t = s.nu_q = FT(0); }
  return s;
}

// One cell's interface solve as an object: init() hoists everything invariant under the iteration, pass() is one
// fixed-point pass (sets `go`), finish() hands the scales back.  solve_cell() runs it start to finish in one thread
// (flux_kernel); flux_refill_kernel keeps one solver per lane and refills a lane as soon as its cell has converged.
#ifndef COFLUX_ICE_PSI_SERIES
#define COFLUX_ICE_PSI_SERIES 0    /* sea-ice pass: ψ(ℓ/L★) by a Taylor polynomial instead of a table row.  A/B knob, OFF: measured on B200
                                      at 1/12° with 92 % ice cover, 15.86 ms with the series against 15.63 ms with the table rows (they are
                                      the same few rows for every cell, i.e. L1-resident), Float32 9.90 against 9.41 ms */
#endif
template <typename FT, int SURF> struct CellSolver {
  using MP = typename DefaultMP<FT>::type;
  CellIn<FT> in;
  Thermo<FT> atm;
  SurfaceState<FT> S;
  FT du, dv, x, theta_a, delta, du2dv2, lnh10, U_ly;
  FT lnh_lu, lnh_lt, lnh_lq;                   // ln(h/ℓ) of the fixed roughness lengths (ice_fast)
  FT Ts, ustar, tstar, qstar, rcdn_ly;
  FT su, st, sq, sT, sr;                       // Brent snapshot, refreshed after 1, 2, 4, … passes
  int it, snap_it, window, stop_at;
  bool ly, fixed, go, ice_fast;

  __device__ __forceinline__ void init(const DevParams<FT>& P, const FluxP<FT>& F, const CellIn<FT>& in_) {
    in = in_;
    const ThermoC<FT>& c = P.th;
    const FT g = P.g, h = P.h;
    atm = phase_equil_pTq<FT, MP>(c, in.pa, in.Ta, in.qa);
    if (F.velocity == COFLUX_VELOCITY_RELATIVE) { du = in.ua - in.us; dv = in.va - in.vs; }
    else { du = in.ua; dv = in.va; }
    x = FT(1);
    if (SURF == 0) { FT s = MP::div(in.So, FT(1000)); x = MP::div(FT(1) - s, FT(1) - s + P.wmf_alpha * s); }
    theta_a = in.Ta + MP::div(g * h, atm.cp_m);
    delta = c.eps - FT(1);
    du2dv2 = du * du + dv * dv;

    Ts = in.Ts0;
    S = surface_state<FT, SURF, MP>(P, F, atm, in.pa, theta_a, x, Ts);

    ustar = F.init; tstar = F.init; qstar = F.init;
    ly = (F.formulation == COFLUX_FLUXES_COEFFICIENT_LARGE_YEAGER);
    U_ly = FT(0); rcdn_ly = FT(0);
    lnh10 = ly ? M<FT>::log(h / FT(10)) : FT(0);
    if (ly) {
      U_ly = M<FT>::max(MP::sqrt(du2dv2), F.ly_umin);
      FT cdn = ly_cdn<FT, MP>(U_ly);
      FT rcdn = MP::sqrt(cdn);
      FT chn = ((S.dtheta > FT(0)) ? FT(18e-3) : FT(32.7e-3)) * rcdn;
      FT cen = FT(34.6e-3) * rcdn;
      rcdn_ly = rcdn;
      ustar = rcdn * U_ly; tstar = MP::div(chn, rcdn) * S.dtheta; qstar = MP::div(cen, rcdn) * S.dq;
    }

    fixed = (F.stop_kind == COFLUX_STOP_FIXED_ITERATIONS);
    it = 0;
    go = fixed ? (F.maxit > 0) : true;
    su = ustar; st = tstar; sq = qstar; sT = Ts; sr = rcdn_ly;   // Brent snapshot, refreshed after 1, 2, 4, … passes
    snap_it = 0; window = 1; stop_at = -1;
    // the sea-ice parameter sets of omip_simulation.jl:52-69,91-113 (similarity theory, fixed roughness lengths, standard
    // log profile, SHEBA or Large–Yeager ψ) take the compact pass_ice(): ψ from the tables, ln(h/ℓ) hoisted
    ice_fast = COFLUX_PSI_TABLES_V1 && SURF == 1 && !ly && (COFLUX_ICE_COARE || F.form == COFLUX_PROFILE_LOGARITHMIC) &&
               F.mr.kind == COFLUX_ROUGHNESS_FIXED && F.tr.kind == COFLUX_ROUGHNESS_FIXED && F.qr.kind == COFLUX_ROUGHNESS_FIXED &&
               (F.stability == COFLUX_STABILITY_SHEBA_PAULSON || F.stability == COFLUX_STABILITY_LARGE_YEAGER) &&
               F.beta >= FT(0) && F.ugmin >= FT(0) && F.mr.fixed > FT(0) && F.tr.fixed > FT(0) && F.qr.fixed > FT(0);
    lnh_lu = lnh_lt = lnh_lq = FT(0);
    if (ice_fast) {
      lnh_lu = M<FT>::log(h / F.mr.fixed); lnh_lt = M<FT>::log(h / F.tr.fixed); lnh_lq = M<FT>::log(h / F.qr.fixed);
    }
  }

  // ψ of the ice-solve parameter sets at one argument: Paulson table (ζ < 0), SHEBA table or −5ζ (ζ ≥ 0); the
  // first-order term below the table (|x| < 2⁻³⁰: the quadratic term is < 1e-18), the formulas above it
  // |x| < 2⁻¹¹: degree-7 Taylor polynomial about 0 (truncation ≤ 9e-17 relative; coefficients from 60-digit mpmath
  // expansions of the Paulson (x < 0) and SHEBA (x > 0) functions).  The arguments ℓ/L★ of the logarithmic profile form
  // live here (ℓ = 5e-4 … 5e-5 m), three of the five ψ evaluations of a pass: no table row, no load.
  static __device__ __forceinline__ FT psi_ice_small(bool stable, FT x, int which) {
    FT c1, c2, c3, c4, c5, c6, c7;
    if (!stable) {
      if (which) { c1 = FT(-8); c2 = FT(-48); c3 = FT(-426.66666666666666667); c4 = FT(-4480); c5 = FT(-51609.6); c6 = FT(-630784); c7 = FT(-8032841.1428571428571); }
      else { c1 = FT(-4); c2 = FT(-20); c3 = FT(-160); c4 = FT(-1560); c5 = FT(-16972.8); c6 = FT(-198016); c7 = FT(-2424685.7142857142857); }
    } else {
      if (which) { c1 = FT(-5); c2 = FT(5); c3 = FT(-8.3333333333333333333); c4 = FT(16.25); c5 = FT(-34); c6 = FT(74.166666666666666667); c7 = FT(-166.42857142857142857); }
      else { c1 = FT(-5); c2 = FT(1.08974358974358974359); c3 = FT(-0.3736576813499890422967); c4 = FT(0.1384112454132177998056);
             c5 = FT(-0.04402388764903304932638); c6 = FT(0.003071835405143235340764); c7 = FT(0.01474035094893286164122); }
    }
    FT p = c7;
    p = p * x + c6; p = p * x + c5; p = p * x + c4; p = p * x + c3; p = p * x + c2; p = p * x + c1;
    return p * x;
  }
  __device__ __forceinline__ FT psi_ice(int stab, FT zz, int which) const {
    if (COFLUX_ICE_PSI_SERIES && M<FT>::abs(zz) < FT(4.8828125e-4) && (zz < FT(0) || stab != COFLUX_STABILITY_LARGE_YEAGER))
      return psi_ice_small(zz >= FT(0), zz, which);
    if (zz < FT(0)) {
      const FT mz = -zz;
      if (psi_tab_in_range(mz)) return psi_tab_eval(PsiTabs<FT>::paulson(), mz, which);
      if (mz < FT(1)) return (which ? FT(-8) : FT(-4)) * zz;
      return which ? psi_scalar(stab, zz) : psi_momentum(stab, zz);
    }
    if (stab == COFLUX_STABILITY_LARGE_YEAGER) return -FT(5) * zz;
    if (psi_tab_in_range(zz)) return psi_tab_eval(PsiTabs<FT>::sheba(), zz, which);
    if (zz < FT(1)) return FT(-5) * zz;
    return which ? psi_scalar(stab, zz) : psi_momentum(stab, zz);
  }
  // ψ_m and ψ_h at one argument (the same branches as psi_ice, taken once)
  __device__ __forceinline__ void psi_ice_pair(int stab, FT zz, FT& pm, FT& ps) const {
    if (zz < FT(0)) {
      if (psi_tab_in_range(-zz)) { psi_tab_eval_pair(PsiTabs<FT>::paulson(), -zz, pm, ps); return; }
    } else if (stab != COFLUX_STABILITY_LARGE_YEAGER) {
      if (psi_tab_in_range(zz)) { psi_tab_eval_pair(PsiTabs<FT>::sheba(), zz, pm, ps); return; }
    }
    pm = psi_ice(stab, zz, 0); ps = psi_ice(stab, zz, 1);
  }
  // ψ_m(za) and ψ_h(zb), za and zb of one sign (ℓ_u/L★ and ℓ_q/L★)
  __device__ __forceinline__ void psi_ice_two(int stab, FT za, FT zb, FT& pa, FT& pb) const {
    if (za < FT(0) && zb < FT(0)) {
      if (psi_tab_in_range(-za) && psi_tab_in_range(-zb)) { psi_tab_eval_two(PsiTabs<FT>::paulson(), -za, -zb, pa, pb); return; }
    } else if (za > FT(0) && zb > FT(0) && stab != COFLUX_STABILITY_LARGE_YEAGER) {
      if (psi_tab_in_range(za) && psi_tab_in_range(zb)) { psi_tab_eval_two(PsiTabs<FT>::sheba(), za, zb, pa, pb); return; }
    }
    pa = psi_ice(stab, za, 0); pb = psi_ice(stab, zb, 1);
  }
  // compact pass for the sea-ice parameter sets (see init): same formulas as pass(), one code path
  __device__ __forceinline__ void pass_ice(const DevParams<FT>& P, const FluxP<FT>& F) {
    const FT g = P.g, h = P.h, kappa = F.kappa;
    const FT u0 = ustar, t0 = tstar, q0 = qstar;
    if (!(u0 > FT(0))) { pass_generic(P, F); return; }   // u★ = 0 (after a zero-scales guard): the generic pass knows the
                                                         // limits — decided BEFORE the skin update, which it does itself
    if (F.itemp == COFLUX_TEMPERATURE_SKIN) skin_temperature(P, F, u0, t0, q0);
    const FT bstar = MP::div(g, S.T_v) * (t0 * (FT(1) + delta * S.q_vap) + delta * S.T_v * q0);
    const FT Jb = -u0 * bstar;
    FT UG = F.ugmin;
    if (Jb > FT(0)) UG = M<FT>::max(F.beta * LMath<FT>::cbrt(Jb * P.hbl), F.ugmin);
    const FT U = MP::sqrt(du2dv2 + UG * UG);
    if (U == FT(0)) {                                  // calm-cell guard, as in the generic pass
      ustar = tstar = qstar = FT(0);
      advance(F, u0, t0, q0);
      return;
    }
    const FT invL = MP::div(kappa * bstar, u0 * u0);  // 1/L★ (0 when b★ = 0)
    const FT zeta = h * invL;
    const int stab = F.stability;
    FT psi_hm, psi_hs;
    psi_ice_pair(stab, zeta, psi_hm, psi_hs);
    const bool logform = (F.form == COFLUX_PROFILE_LOGARITHMIC);     // the COARE form drops the ψ(ℓ/L★) terms
    FT prof_u = lnh_lu - psi_hm;
    FT pq = FT(0);
    if (logform) {                                   // ψ_m(ℓ_u/L★) and ψ_h(ℓ_q/L★) together (the second is only used when χ_u > 0)
      FT pu;
      psi_ice_two(stab, F.mr.fixed * invL, F.qr.fixed * invL, pu, pq);
      prof_u += pu;
    }
    if (!(prof_u > FT(0))) {
      ustar = tstar = qstar = FT(0);
    } else {
      FT prof_q = lnh_lq - psi_hs, prof_t = lnh_lt - psi_hs;
      if (logform) {
        prof_q += pq;
        prof_t += (F.tr.fixed == F.qr.fixed) ? pq : psi_ice(stab, F.tr.fixed * invL, 1);   // the same number, bit for bit
      }
      const FT chi_u = MP::div(kappa, prof_u);
      const FT chi_q = (prof_q > FT(0)) ? MP::div(kappa, prof_q) : FT(0);
      const FT chi_t = (prof_t > FT(0)) ? MP::div(kappa, prof_t) : FT(0);
      ustar = chi_u * U; tstar = chi_t * S.dtheta; qstar = chi_q * S.dq;
    }
    advance(F, u0, t0, q0);
  }
  __device__ __forceinline__ void pass(const DevParams<FT>& P, const FluxP<FT>& F) {
    if (ice_fast) pass_ice(P, F);
    else pass_generic(P, F);
  }
  // conductive flux balance through the slab (row a7): T_s relaxes towards the balance temperature, clamped
  __device__ __forceinline__ void skin_temperature(const DevParams<FT>& P, const FluxP<FT>& F, FT u0, FT t0, FT q0) {
    const ThermoC<FT>& c = P.th;
    // conductive flux balance through the slab (row a7)
    FT Tb = P.io.T0 - P.io.slope * in.S_ice + P.T_offset;
    FT Tm = P.io.T0 + P.T_offset;
    FT Ls = c.LH_s0 + (c.cp_v - c.cp_i) * (in.Ta - c.T_0);
    FT Qu = P.emis_i * P.sigma * Ts * Ts * Ts * Ts;
    FT Qd = -(FT(1) - in.albedo) * in.Qs - P.emis_i * in.Ql;
    FT Qc = -atm.rho * atm.cp_m * u0 * t0;
    FT Qv = -atm.rho * Ls * u0 * q0;
    FT Qa = Qv + Qu + Qc + Qd;
    FT Tstar = Tb - MP::div(Qa * in.h_ice, P.io.k_ice);      // MP::div: the lean division (bit-identical to IEEE, a third of the instructions)
    if (F.skin_update == COFLUX_SKIN_LINEARIZED_LONGWAVE) {     // emitted long wave implicit: Q_u ≈ σ ε T_s⁻³ · T_s⁺
      const FT alpha = P.sigma * P.emis_i * Ts * Ts * Ts / P.io.k_ice;
      Tstar = (Tb - (Qd + Qc + Qv) * in.h_ice / P.io.k_ice) / (FT(1) + alpha * in.h_ice);
    }
    if (Tstar != Tstar) Tstar = Ts;
    Tstar = M<FT>::max(FT(0), Tstar);
    FT Tnew = (in.h_ice >= P.io.h_c) ? Tstar : Tb;
    FT dT = Tnew - Ts;
    FT adT = M<FT>::min(F.skin_max_dT, M<FT>::abs(dT));
    FT sgn = (dT > FT(0)) ? FT(1) : ((dT < FT(0)) ? FT(-1) : FT(0));
    Ts = M<FT>::min(Ts + adT * sgn, Tm);
    S = surface_state<FT, SURF, MP>(P, F, atm, in.pa, theta_a, x, Ts, !ice_fast);
  }
  __device__ __forceinline__ void pass_generic(const DevParams<FT>& P, const FluxP<FT>& F) {
    const ThermoC<FT>& c = P.th;
    const FT g = P.g, h = P.h, kappa = F.kappa;
    const FT u0 = ustar, t0 = tstar, q0 = qstar;
    if (SURF == 1 && F.itemp == COFLUX_TEMPERATURE_SKIN) skin_temperature(P, F, u0, t0, q0);
    const FT bstar = g / S.T_v * (t0 * (FT(1) + delta * S.q_vap) + delta * S.T_v * q0);
    if (ly) {
      FT zeta = MP::div(kappa * bstar * h, u0 * u0);
      zeta = M<FT>::max(FT(-10), M<FT>::min(FT(10), zeta));
      FT psim, psih;                           // the Large–Yeager pair inline: one table row for both (ζ < 0), −5ζ (ζ ≥ 0)
      if (COFLUX_PSI_TABLES_V1 && zeta < FT(0) && psi_tab_in_range(-zeta)) psi_tab_eval_pair(PsiTabs<FT>::paulson(), -zeta, psim, psih);
      else if (zeta >= FT(0)) psim = psih = -FT(5) * zeta;
      else { psim = psi_momentum(COFLUX_STABILITY_LARGE_YEAGER, zeta); psih = psi_scalar(COFLUX_STABILITY_LARGE_YEAGER, zeta); }
      FT U10N = MP::div(U_ly, FT(1) + MP::div(rcdn_ly, kappa) * (lnh10 - psim));
      U10N = M<FT>::max(U10N, F.ly_umin);
      FT cdn = ly_cdn<FT, MP>(U10N);
      FT rcdn = MP::sqrt(cdn);
      FT cen = FT(34.6e-3) * rcdn;
      FT chn = ((zeta > FT(0)) ? FT(18e-3) : FT(32.7e-3)) * rcdn;
      FT xm = FT(1) + MP::div(rcdn, kappa) * (lnh10 - psim);
      FT cd = MP::div(cdn, xm * xm);
      FT rr = MP::sqrt(MP::div(cd, cdn));
      FT ch = MP::div(chn, FT(1) + MP::div(chn, kappa * rcdn) * (lnh10 - psih)) * rr;
      FT ce = MP::div(cen, FT(1) + MP::div(cen, kappa * rcdn) * (lnh10 - psih)) * rr;
      rcdn_ly = rcdn;
      FT rcd = MP::sqrt(cd);
      ustar = rcd * U_ly; tstar = MP::div(ch, rcd) * S.dtheta; qstar = MP::div(ce, rcd) * S.dq;
    } else {
      const FT Jb = -u0 * bstar;
      FT UG = F.beta * M<FT>::cbrt(Jb * P.hbl);
      UG = M<FT>::max(UG, F.ugmin);
      const FT U = M<FT>::sqrt(du2dv2 + UG * UG);
      if (U == FT(0)) {  // documented calm-cell guard
        ustar = tstar = qstar = FT(0);
      } else {
        const FT lu = momentum_roughness(F.mr, u0, U, S.nu_m);
        const FT lq = scalar_roughness(F.qr, lu, u0, S.nu_q);
        const FT Lstar = (bstar == FT(0)) ? M<FT>::inf() : u0 * u0 / (kappa * bstar);
        const FT zeta = h / Lstar;
        const FT psi_hm = psi_momentum(F.stability, zeta);
        const FT psi_hs = psi_scalar(F.stability, zeta);
        const FT prof_u = similarity_profile<FT, false>(F.form, F.stability, h, lu, Lstar, psi_hm);
        if (!(prof_u > FT(0))) {  // documented guard (COARE form only): reset to zero scales
          ustar = tstar = qstar = FT(0);
        } else {
          const FT prof_q = similarity_profile<FT, true>(F.form, F.stability, h, lq, Lstar, psi_hs);
          const FT chi_u = kappa / prof_u;
          const FT chi_q = (prof_q > FT(0)) ? kappa / prof_q : FT(0);
          FT chi_t = chi_q;
          if (!F.same_scalar) {
            const FT lt = scalar_roughness(F.tr, lu, u0, S.nu_t);
            const FT prof_t = similarity_profile<FT, true>(F.form, F.stability, h, lt, Lstar, psi_hs);
            chi_t = (prof_t > FT(0)) ? kappa / prof_t : FT(0);
          }
          ustar = chi_u * U; tstar = chi_t * S.dtheta; qstar = chi_q * S.dq;
        }
      }
    }
    advance(F, u0, t0, q0);
  }
  // bookkeeping after a pass: iteration count, stop rule, Brent cycle detection
  __device__ __forceinline__ void advance(const FluxP<FT>& F, FT u0, FT t0, FT q0) {
    ++it;
    if (fixed) {
      go = it < F.maxit;
    } else if (stop_at >= 0) {                  // finishing a detected limit cycle
      go = it < stop_at;
      if (!go) it = F.maxit;
    } else {
      FT drift = M<FT>::abs(ustar - u0) + M<FT>::abs(tstar - t0) + M<FT>::abs(qstar - q0);
      go = !((drift < F.tol) || (it >= F.maxit));
      // Brent cycle detection.  The pass is a pure function of (u★, θ★, q★, T_s, √Cd_N); with a skin temperature
      // (ΔT clamp, melting cap) one cell in nine never meets the stop rule but flips between two states for ever.
      // The reference iterates such a cell to maxiter; the state it ends in is the one (maxiter − it) mod λ passes
      // further along the orbit, λ the period — run just those.  Bit-identical to iterating on.
      if (go) {
        if (ustar == su && tstar == st && qstar == sq && Ts == sT && rcdn_ly == sr) {
          const int lambda = it - snap_it;
          stop_at = it + (F.maxit - it) % lambda;
          go = it < stop_at;
          if (!go) it = F.maxit;
        } else if (it - snap_it == window) {
          su = ustar; st = tstar; sq = qstar; sT = Ts; sr = rcdn_ly; snap_it = it; window *= 2;
        }
      }
    }
  }

  __device__ __forceinline__ void finish(CellOut<FT>& out) const {
    out.ustar = ustar; out.tstar = tstar; out.qstar = qstar; out.Ts = Ts;
    out.rho_a = atm.rho; out.cp_a = atm.cp_m; out.du = du; out.dv = dv; out.it = it;
  }
};

template <typename FT, int SURF>
__device__ void solve_cell(const DevParams<FT>& P, const FluxP<FT>& F, const CellIn<FT>& in, CellOut<FT>& out) {
  CellSolver<FT, SURF> s;
  s.init(P, F, in);
  while (s.go) s.pass(P, F);
  s.finish(out);
}

}  // namespace coflux

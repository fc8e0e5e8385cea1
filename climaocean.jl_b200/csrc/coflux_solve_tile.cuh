// coflux_solve_tile.cuh — the production atmosphere–ocean flux kernel (similarity theory, bulk
// interface temperature).
//
// The one-cell-per-thread kernel (coflux_kernels.cuh::flux_kernel, still used for the Large–Yeager solve and for
// parameter sets this kernel is not eligible for) showed what limits a straightforward mapping: 21 of 32 lanes active on
// average (cells of one warp need 9…27 passes and take different ψ branches), 128 registers → 16 warps/SM.  This
// kernel splits the work of a tile of TILE cells into three convergent phases separated by __syncthreads():
//
//   A  per cell: coalesced loads, atmosphere interpolation, exchange-state stores, both thermodynamic states
//      (COFLUX_TILE_PRE lock-step similarity passes could follow here; measured best: none).  The few scalars the
//      iteration needs go to shared memory as a "task"; tasks are queued sorted by stability class.
//   B  lanes pop tasks from the shared queue and iterate; a lane whose cell has converged writes
//      its result back and immediately pops the next task ("lane refill"), so a warp has no idle
//      lanes until the tile's queue drains, and — because the queue is class-sorted — the lanes of a warp take the
//      same ψ branch.
//   C  per cell: fluxes, net-flux assembly, coalesced stores.
//
// Two forms of the pass live here.  iterate_fast (any precision, any eligible parameter set; the exact fall-back of
// the lean pass) is algebraically the reference iteration (SURVEY Appendix A4–A6) with rounding-level reformulations
// (each ≤ a few ulp, covered by the 1e-12 parity tests):
//   * 1/L★ = κ b★/u★² once, then ζ = h/L★ and ℓ/L★ are products (three divisions fewer);
//   * g/T_v, 1+δq_v, δT_v hoisted out of the loop (exactly the same values);
//   * the gustiness cube root is skipped when the buoyancy flux is ≤ 0 (the floor wins anyway);
//   * Reynolds-scaling scalar roughness ℓ = A·R★^(−b): ln(h/ℓ) = ln(h/A) + b·ln R★ — one log instead
//     of pow + log; ℓ itself (needed only inside ψ(ℓ/L)) from one exp;
//   * the Edson ψ_u/ψ_θ pair at the same ζ shares √(1−15ζ), ζ²/(1+ζ²), exp(−0.35ζ); x^1.5 = x·√x;
//   * ψ(ℓ/L) at |ℓ/L| ≤ 2⁻⁹ by the Taylor series of the same function (tools/gen_psi_taylor.py; |error| < 3e-18);
//   * the unstable Edson ψ_u, ψ_θ on −ζ ∈ [2⁻⁹, 2¹³) from the table of degree-7 piecewise polynomials
//     (16 per binade, tools/gen_psi_table.py; error ≤ 3e-16·max(1,|ψ|), i.e. rounding level) instead of
//     4 log + 3 atan + 2 cbrt + 2 sqrt + 5 divisions; outside that range the exact formulas are used;
//   * a limit cycle of the iterate (common in Float32, where Σ|Δ| < 1e-8 is below the resolution:
//     8 % of the cells never "converge") is detected exactly (Brent) and the state the reference
//     reaches at maxiter is produced after < period extra passes — bit-identical to iterating on.
// iterate_lean (SPEC 1 / SPEC 2, both precisions) is the pass the hot loop runs; see its own header below.
#pragma once
#include "coflux_kernels.cuh"
#include "coflux_psi_table.h"
#include <type_traits>

namespace coflux {

template <typename FT> __device__ __forceinline__ FT fma_(FT a, FT b, FT c);
template <> __device__ __forceinline__ double fma_<double>(double a, double b, double c) { return ::fma(a, b, c); }
template <> __device__ __forceinline__ float fma_<float>(float a, float b, float c) { return ::fmaf(a, b, c); }

// Taylor series about 0 of the Edson et al. (2013) functions on each branch (coefficients from
// tools/gen_psi_taylor.py; literals so that they fold into immediates in either precision)
template <typename FT> __device__ __forceinline__ FT psi_series_m_unst(FT x) {
  FT acc = FT(-5953014842481.911);
  acc = fma_<FT>(acc, x, FT(-456221314784.8534));
  acc = fma_<FT>(acc, x, FT(-35388422828.92556));
  acc = fma_<FT>(acc, x, FT(-2784789427.619603));
  acc = fma_<FT>(acc, x, FT(-222988725.74826965));
  acc = fma_<FT>(acc, x, FT(-18243598.617862545));
  acc = fma_<FT>(acc, x, FT(-1533783.4280264128));
  acc = fma_<FT>(acc, x, FT(-133623.07255925206));
  acc = fma_<FT>(acc, x, FT(-12220.416809168435));
  acc = fma_<FT>(acc, x, FT(-1198.931685655382));
  acc = fma_<FT>(acc, x, FT(-131.46927083333333));
  acc = fma_<FT>(acc, x, FT(-17.578125));
  acc = fma_<FT>(acc, x, FT(-3.75));
  acc = fma_<FT>(acc, x, FT(0.0));
  return acc;
}
template <typename FT> __device__ __forceinline__ FT psi_series_m_stab(FT x) {
  FT acc = FT(-3.2826171875e-06);
  acc = fma_<FT>(acc, x, FT(6.018131510416667e-05));
  acc = fma_<FT>(acc, x, FT(-0.000937890625));
  acc = fma_<FT>(acc, x, FT(0.01205859375));
  acc = fma_<FT>(acc, x, FT(-0.1225));
  acc = fma_<FT>(acc, x, FT(0.91875));
  acc = fma_<FT>(acc, x, FT(-5.2));
  acc = fma_<FT>(acc, x, FT(0.0));
  return acc;
}
template <typename FT> __device__ __forceinline__ FT psi_series_s_unst(FT x) {
  FT acc = FT(-4.460966866604527e+17);
  acc = fma_<FT>(acc, x, FT(-1.5085597225417256e+16));
  acc = fma_<FT>(acc, x, FT(-522821942131960.25));
  acc = fma_<FT>(acc, x, FT(-18867805615839.246));
  acc = fma_<FT>(acc, x, FT(-728767553788.2886));
  acc = fma_<FT>(acc, x, FT(-31347223918.678947));
  acc = fma_<FT>(acc, x, FT(-1562843283.4957829));
  acc = fma_<FT>(acc, x, FT(-91771923.36737002));
  acc = fma_<FT>(acc, x, FT(-6233166.161298048));
  acc = fma_<FT>(acc, x, FT(-473686.6082199878));
  acc = fma_<FT>(acc, x, FT(-39314.5732184928));
  acc = fma_<FT>(acc, x, FT(-3548.0861371527776));
  acc = fma_<FT>(acc, x, FT(-355.4458333333333));
  acc = fma_<FT>(acc, x, FT(-42.1875));
  acc = fma_<FT>(acc, x, FT(-7.5));
  acc = fma_<FT>(acc, x, FT(0.0));
  return acc;
}
template <typename FT> __device__ __forceinline__ FT psi_series_s_stab(FT x) {
  FT acc = FT(0.00025428425045974796);
  acc = fma_<FT>(acc, x, FT(-0.0005466523981695816));
  acc = fma_<FT>(acc, x, FT(0.0007096960570987654));
  acc = fma_<FT>(acc, x, FT(0.006086738425925926));
  acc = fma_<FT>(acc, x, FT(-0.09034314814814814));
  acc = fma_<FT>(acc, x, FT(0.6497666666666667));
  acc = fma_<FT>(acc, x, FT(-4.998666666666667));
  acc = fma_<FT>(acc, x, FT(-0.005));
  return acc;
}
#define COFLUX_PSI_SMALL 0.001953125 /* 2^-9 */

// Edson ψ_u(z), ψ_θ(z) at the same argument
template <typename FT> __device__ __noinline__ void psi_edson_pair(FT z, FT& pm, FT& ps) {
  if (z >= FT(0)) {
    const FT dz = M<FT>::min(FT(50), FT(0.35) * z);
    const FT e = M<FT>::exp(-dz);
    pm = -FT(0.7) * z - FT(0.75) * (z - FT(5) / FT(0.35)) * e - FT(0.75) * FT(5) / FT(0.35);
    const FT w = FT(1) + FT(2) / FT(3) * z;
    ps = -(w * M<FT>::sqrt(w)) - FT(2) / FT(3) * (z - FT(14.28)) * e - FT(8.525);
  } else {
    const FT s = M<FT>::sqrt(FT(1) - FT(15) * z);      // (1−15ζ)^{1/2}: x_θ, and x_u²
    const FT xu = M<FT>::sqrt(s);
    const FT lg = M<FT>::log((FT(1) + s) / FT(2));
    const FT pku = FT(2) * M<FT>::log((FT(1) + xu) / FT(2)) + lg - FT(2) * M<FT>::atan(xu) + M<FT>::pi() / FT(2);
    const FT pks = FT(2) * lg;
    const FT pcu = psi_conv_cbrt(M<FT>::cbrt(FT(1) - FT(10.15) * z));
    const FT pcs = psi_conv_cbrt(M<FT>::cbrt(FT(1) - FT(34.15) * z));
    const FT f = z * z / (FT(1) + z * z);
    pm = (FT(1) - f) * pku + f * pcu;
    ps = (FT(1) - f) * pks + f * pcs;
  }
}
// ---------------------------------------------------------------------------------------------
// Table evaluation of the unstable Edson functions (coflux_psi_table.h).  z = −ζ ∈ [2^KMIN, 2^KMAX).
// The interval index comes straight from the floating-point bits (exponent + top 4 mantissa bits),
// the local coordinate t ∈ [−1,1) from the remaining mantissa bits — all exact operations.
// ---------------------------------------------------------------------------------------------
template <typename FT> struct PsiTable;
template <> struct PsiTable<double> {
  static __device__ __forceinline__ const double* row(double z, double& t) {
    const long long bits = __double_as_longlong(z);
    const int k = (int)((bits >> 52) & 0x7ff) - 1023;
    const int j = (int)((bits >> 48) & 0xf);
    const double one_plus_u = __longlong_as_double(((bits & 0x0000ffffffffffffLL) << 4) | 0x3ff0000000000000LL);
    t = 2.0 * one_plus_u - 3.0;
    return &COFLUX_PSI_TABLE_F64[(k - COFLUX_PSI_KMIN) * COFLUX_PSI_NS + j][0][0];
  }
  static __device__ __forceinline__ bool in_range(double z) { return z >= 0.001953125 && z < (double)(1 << COFLUX_PSI_KMAX); }
};
template <> struct PsiTable<float> {
  static __device__ __forceinline__ const float* row(float z, float& t) {
    const int bits = __float_as_int(z);
    const int k = ((bits >> 23) & 0xff) - 127;
    const int j = (bits >> 19) & 0xf;
    const float one_plus_u = __int_as_float(((bits & 0x0007ffff) << 4) | 0x3f800000);
    t = 2.0f * one_plus_u - 3.0f;
    return &COFLUX_PSI_TABLE_F32[(k - COFLUX_PSI_KMIN) * COFLUX_PSI_NS + j][0][0];
  }
  static __device__ __forceinline__ bool in_range(float z) { return z >= 0.001953125f && z < (float)(1 << COFLUX_PSI_KMAX); }
};
static_assert(COFLUX_PSI_NS == 16 && COFLUX_PSI_DEG == 7 && COFLUX_PSI_KMIN <= -9 && COFLUX_PSI_KMAX >= 7 && COFLUX_PSI_KMAX < 31, "table layout changed");
template <typename FT> __device__ __forceinline__ FT poly8(const FT* c, FT t) {
  FT acc = __ldg(c + 7);
#pragma unroll
  for (int k = 6; k >= 0; --k) acc = fma_<FT>(acc, t, __ldg(c + k));
  return acc;
}
// ψ_u(ζ) and ψ_θ(ζ), Edson et al. (2013): series near 0, table on the bulk of the unstable range,
// exact formulas elsewhere (stable side, and −ζ ≥ 2^KMAX which only occurs in the start-up transient of calm cells)
template <typename FT> __device__ __forceinline__ void psi_edson_pair_fast(FT z, FT& pm, FT& ps) {
  if (z < FT(0)) {
    const FT mz = -z;
    if (__builtin_expect(PsiTable<FT>::in_range(mz), 1)) {
      FT t;
      const FT* c = PsiTable<FT>::row(mz, t);
      pm = poly8<FT>(c, t); ps = poly8<FT>(c + 8, t);
      return;
    }
    if (mz <= FT(COFLUX_PSI_SMALL)) { pm = psi_series_m_unst<FT>(z); ps = psi_series_s_unst<FT>(z); return; }
  }
  psi_edson_pair<FT>(z, pm, ps);
}
template <typename FT> __device__ __forceinline__ FT psi_edson_u_fast(FT x) {
  if (M<FT>::abs(x) <= FT(COFLUX_PSI_SMALL)) return (x >= FT(0)) ? psi_series_m_stab<FT>(x) : psi_series_m_unst<FT>(x);
  if (x < FT(0) && PsiTable<FT>::in_range(-x)) { FT t; const FT* c = PsiTable<FT>::row(-x, t); return poly8<FT>(c, t); }
  return psi_momentum((int)COFLUX_STABILITY_EDSON, x);
}
template <typename FT> __device__ __forceinline__ FT psi_edson_t_fast(FT x) {
  if (M<FT>::abs(x) <= FT(COFLUX_PSI_SMALL)) return (x >= FT(0)) ? psi_series_s_stab<FT>(x) : psi_series_s_unst<FT>(x);
  if (x < FT(0) && PsiTable<FT>::in_range(-x)) { FT t; const FT* c = PsiTable<FT>::row(-x, t); return poly8<FT>(c + 8, t); }
  return psi_scalar((int)COFLUX_STABILITY_EDSON, x);
}

template <typename FT, int SPEC> __device__ __forceinline__ FT psi_small_momentum(bool edson, int stab, FT x) {
  if (SPEC) return psi_edson_u_fast<FT>(x);
  if (edson && M<FT>::abs(x) <= FT(COFLUX_PSI_SMALL)) return (x >= FT(0)) ? psi_series_m_stab<FT>(x) : psi_series_m_unst<FT>(x);
  return psi_momentum(SPEC ? (int)COFLUX_STABILITY_EDSON : stab, x);
}
template <typename FT, int SPEC> __device__ __forceinline__ FT psi_small_scalar(bool edson, int stab, FT x) {
  if (SPEC) return psi_edson_t_fast<FT>(x);
  if (edson && M<FT>::abs(x) <= FT(COFLUX_PSI_SMALL)) return (x >= FT(0)) ? psi_series_s_stab<FT>(x) : psi_series_s_unst<FT>(x);
  return psi_scalar(SPEC ? (int)COFLUX_STABILITY_EDSON : stab, x);
}

// ln(h/ℓ) and ℓ of one scalar roughness length.  Returns false when ℓ = 0 (profile = +∞ ⇒ χ = 0).
// SPEC != 0: the parameterisation is known at compile time to be Reynolds scaling with A, b, ℓmax > 0.
template <typename FT, int SPEC>
__device__ __forceinline__ bool scalar_log_roughness(const ScaRough<FT>& r, bool fast, FT lnhA, FT lnhl, FT lrclip, bool need_l, FT lnh,
                                                     FT lu, FT u0, FT nu, FT& ln_h_l, FT& l) {
  if (!SPEC && r.kind == COFLUX_ROUGHNESS_FIXED) { ln_h_l = lnhl; l = r.fixed; return true; }
  const FT Rstar = lu * u0 / nu;
  if (Rstar == FT(0)) { l = FT(0); ln_h_l = M<FT>::inf(); return false; }
  if ((SPEC || fast) && Rstar > FT(0)) {
    const FT lr = M<FT>::log(Rstar);
    if (lr > lrclip) {                       // A·R★^(−b) < ℓ_max
      ln_h_l = fma_<FT>(r.b, lr, lnhA);
      l = need_l ? r.A * M<FT>::exp(-r.b * lr) : FT(0);
    } else {
      ln_h_l = lnhl; l = r.lmax;
    }
    return true;
  }
  if (SPEC) { l = FT(0); ln_h_l = M<FT>::inf(); return false; }   // R★ < 0 or NaN: unreachable for sane iterates
  l = M<FT>::min(r.A / M<FT>::pow(Rstar, r.b), r.lmax);
  ln_h_l = lnh - M<FT>::log(l);
  return true;
}

// One fixed-point pass (A4), fast formulation.  Inputs: invariants of the cell; in/out: the scales.
// SPEC selects a compile-time specialisation so that the hot loop carries no dead generic code:
//   0  everything decided at run time (any parameter set the tile kernel is eligible for)
//   1  `:default`   — Edson ψ, standard log profile, constant Charnock, Reynolds-scaling scalars (θ ≡ q)
//   2  `:corrected` — Edson ψ, COARE log profile, wind-dependent Charnock, Reynolds-scaling scalars (θ ≡ q)
template <typename FT, int SPEC>
__device__ __forceinline__ void iterate_fast(const DevParams<FT>& P, const FluxP<FT>& F, const FastConsts<FT>& K, FT U2, FT dth, FT dq,
                                             FT gTv, FT a1, FT a2, FT nu, FT& us, FT& ts, FT& qs) {
  const bool edson = SPEC ? true : (K.edson != 0);
  const bool logform = (SPEC == 1) ? true : ((SPEC == 2) ? false : (F.form == COFLUX_PROFILE_LOGARITHMIC));
  const bool gust_skip = SPEC ? true : (K.gust_skip != 0);
  const bool same_scalar = SPEC ? true : (F.same_scalar != 0);
  const FT u0 = us, t0 = ts, q0 = qs;
  const FT kappa = F.kappa, h = P.h;
  const FT bstar = gTv * (t0 * a1 + a2 * q0);
  const FT Jb = -u0 * bstar;
  FT UG = F.ugmin;
  if (!gust_skip || Jb > FT(0)) UG = M<FT>::max(F.beta * M<FT>::cbrt(Jb * P.hbl), F.ugmin);
  const FT U = M<FT>::sqrt(U2 + UG * UG);
  if (U == FT(0)) { us = ts = qs = FT(0); return; }
  FT lu;
  if (SPEC) {
    FT alpha = F.mr.alpha;
    if (SPEC == 2) alpha = M<FT>::max(F.mr.a1 * M<FT>::min(U, F.mr.umax) + F.mr.a2, F.mr.amin);
    const FT lR = (u0 == FT(0)) ? F.mr.lmax : F.mr.beta_s * nu / u0;
    lu = M<FT>::min(alpha * u0 * u0 / F.mr.g + lR, F.mr.lmax);
  } else {
    lu = momentum_roughness(F.mr, u0, U, nu);
  }
  const FT invL = (bstar == FT(0)) ? FT(0) : kappa * bstar / (u0 * u0);
  const FT zeta = h * invL;
  FT psi_hm, psi_hs;
  if (SPEC) psi_edson_pair_fast<FT>(zeta, psi_hm, psi_hs);
  else if (edson) psi_edson_pair<FT>(zeta, psi_hm, psi_hs);
  else { psi_hm = psi_momentum(F.stability, zeta); psi_hs = psi_scalar(F.stability, zeta); }
  FT prof_u = (K.lnh - M<FT>::log(lu)) - psi_hm;
  if (logform) prof_u += psi_small_momentum<FT, SPEC>(edson, F.stability, lu * invL);
  if (!(prof_u > FT(0))) { us = ts = qs = FT(0); return; }
  FT lnq, lq;
  FT chi_q = FT(0);
  if (scalar_log_roughness<FT, SPEC>(F.qr, K.fast_q != 0, K.lnhA_q, K.lnhl_q, K.lrclip_q, logform, K.lnh, lu, u0, nu, lnq, lq)) {
    FT prof_q = lnq - psi_hs;
    if (logform) prof_q += psi_small_scalar<FT, SPEC>(edson, F.stability, lq * invL);
    chi_q = (prof_q > FT(0)) ? kappa / prof_q : FT(0);
  }
  FT chi_t = chi_q;
  if (!same_scalar) {
    FT lnt, lt;
    chi_t = FT(0);
    if (scalar_log_roughness<FT, 0>(F.tr, K.fast_t != 0, K.lnhA_t, K.lnhl_t, K.lrclip_t, logform, K.lnh, lu, u0, nu, lnt, lt)) {
      FT prof_t = lnt - psi_hs;
      if (logform) prof_t += psi_small_scalar<FT, 0>(edson, F.stability, lt * invL);
      chi_t = (prof_t > FT(0)) ? kappa / prof_t : FT(0);
    }
  }
  const FT chi_u = kappa / prof_u;
  us = chi_u * U; ts = chi_t * dth; qs = chi_q * dq;
}

template <typename FT> __device__ __forceinline__ bool keep_going(const FluxP<FT>& F, int it, FT us, FT ts, FT qs, FT u0, FT t0, FT q0) {
  if (F.stop_kind == COFLUX_STOP_FIXED_ITERATIONS) return it < F.maxit;
  const FT drift = M<FT>::abs(us - u0) + M<FT>::abs(ts - t0) + M<FT>::abs(qs - q0);
  return !((drift < F.tol) || (it >= F.maxit));
}

// ---------------------------------------------------------------------------------------------
// Lean Float64 pass (SPEC 1 / SPEC 2 only).  Same reference iteration, organised so that the loop is
// straight-line code on registers:
//   * every elementary function is the branch-free FMA sequence of coflux_fastmath.cuh (SFU seed + Newton,
//     table log/exp); the three divisions by u★ share one reciprocal;
//   * the sign of b★ selects ONE of two blocks (unstable: gustiness cube root, ψ pair from the table, 5-term
//     series for ψ(ℓ/L); stable: closed forms with the lean exp/√, 3-term series) — the queue is sorted by
//     that sign, so warps do not diverge;
//   * anything outside the domain of the short path (u★ ≤ 2⁻¹⁰⁰, a calm cell, −ζ outside [2⁻²⁰, 2¹³) — the table
//     itself reaches down to 2⁻³⁰ —,
//     an out-of-range buoyancy-flux argument) sets one flag, and the caller redoes the whole pass
//     with the exact code behind a single by-value call (lean_cold_pass) — rare after the start-up transient;
//   * all literals come from constant memory (LeanLit), not 64-bit immediates.
// ≈ 165 FP64 instructions per pass instead of ≈ 540.  Every step differs from iterate_fast by rounding-level
// amounts only; tests/test_gpu_parity.py holds the kernel to the same 1e-12 bar, and COFLUX_LEAN=0 builds
// the previous loop for A/B runs.
// ---------------------------------------------------------------------------------------------
#ifndef COFLUX_LEAN
#define COFLUX_LEAN 1
#endif
#ifndef COFLUX_LEAN_F32
#define COFLUX_LEAN_F32 1
#endif
// Shared-memory tables of the lean pass: log / exp tables (Float64 only) and the HOT part of the unstable Edson ψ table.
//
// Why the ψ table is staged in shared memory (round 2, profiles/README.md): every lane looks up its own row (128 B in
// Float64), so one LDG.128 of a warp touches up to 32 different cache lines = 32 L1 wavefronts, 8 such loads per pass.
// ncu showed the L1 data pipe 84 % busy (l1tex__data_pipe_lsu_wavefronts) — THAT, not FP64 latency, bounded the kernel.
// Shared memory serves 8 different rows per wavefront when their 16-byte pieces fall into different bank groups, so the
// rows are stored XOR-swizzled: piece p of row r lives at piece position p ^ (r & 7).  Lanes hold unrelated rows, hence
// a lane's bank group is uniformly random: ≈ 7 wavefronts per LDS.128 instead of 32.  Values, evaluation order and
// therefore results are bit-identical to the global-memory path, which still serves the rows outside the hot range.
// Hot range: −ζ ∈ [2^KLO, 2^KHI) = [2⁻⁷, 2⁶) holds 96.6 % of the unstable passes on the SURVEY §8d inputs (26 KB).
#ifndef COFLUX_PSI_SM_KLO
#define COFLUX_PSI_SM_KLO (-7)
#endif
#ifndef COFLUX_PSI_SM_KHI
#define COFLUX_PSI_SM_KHI 6
#endif
#define COFLUX_PSI_SM_ROW0 ((COFLUX_PSI_SM_KLO - COFLUX_PSI_KMIN) * COFLUX_PSI_NS)
#define COFLUX_PSI_SM_ROWS ((COFLUX_PSI_SM_KHI - COFLUX_PSI_SM_KLO) * COFLUX_PSI_NS)
static_assert(COFLUX_PSI_SM_KLO >= COFLUX_PSI_KMIN && COFLUX_PSI_SM_KHI <= COFLUX_PSI_KMAX && COFLUX_PSI_SM_KLO < COFLUX_PSI_SM_KHI, "hot range outside the table");
struct LeanTabs { const double* lg; const double* ex; unsigned psi; };   // psi: shared-window byte address of the swizzled rows
template <typename FT> struct LeanCell { FT U2, Ustab, dth, dq, cb1, cb2, bnu, inv_nu; };   // b★ = cb1·θ★ + cb2·q★
template <typename FT> struct D3 { FT u, t, q; };

// Literals of the hot loop.  In Float64 those with ≤ 21 significant bits are written in place: an FP64 instruction
// takes them as a 32-bit immediate (the high word).  High-order series coefficients only need a few correct bits
// (they multiply |x|^k ≤ 2^(−13k)), so they are rounded to that form; the rest sit in constant memory.
template <typename FT> struct LeanLit {
  FT su3, ms2, ms1, ss2, ss1, ss0;
  FT c035, c5_035, m07, m075c, two3, m23, c1428, m8525;
};
#define COFLUX_LEAN_LITERALS(T) { T(-355.4458333333333), T(0.91875), T(-5.2), T(0.6497666666666667), T(-4.998666666666667), T(-0.005), \
  T(0.35), T(5.0 / 0.35), T(-0.7), T(-0.75 * 5.0 / 0.35), T(2.0 / 3.0), T(-(2.0 / 3.0)), T(14.28), T(-8.525) }
__constant__ LeanLit<double> LL64 = COFLUX_LEAN_LITERALS(double);
__constant__ LeanLit<float> LL32 = COFLUX_LEAN_LITERALS(float);
template <typename FT> __device__ __forceinline__ const LeanLit<FT>& lean_lit();
template <> __device__ __forceinline__ const LeanLit<double>& lean_lit<double>() { return LL64; }
template <> __device__ __forceinline__ const LeanLit<float>& lean_lit<float>() { return LL32; }
#define COFLUX_NROWS ((COFLUX_PSI_KMAX - COFLUX_PSI_KMIN) * COFLUX_PSI_NS)

// Row index of the ψ table for z (clamped into the table, so that the loads are safe for ANY z: the caller may then
// issue them before it knows whether z is in range, and the scheduler can hoist them above the cube root) and
// the local coordinate t ∈ [−1, 1).  A row holds the 8 coefficients of ψ_u followed by the 8 of ψ_θ.
__device__ __forceinline__ int psi_table_row(double z, double& t) {
  const long long bits = __double_as_longlong(z);
  const int hi = (int)(bits >> 32);
  int row = (hi >> 16) - ((1023 + COFLUX_PSI_KMIN) << 4);              // (exponent − KMIN)·16 + top 4 mantissa bits
  row = max(0, min(row, COFLUX_NROWS - 1));
  const double one_plus_u = __longlong_as_double(((bits & 0x0000ffffffffffffLL) << 4) | 0x3ff0000000000000LL);
  t = fm::fma_(2.0, one_plus_u, -3.0);
  return row;
}
__device__ __forceinline__ int psi_table_row(float z, float& t) {
  const int bits = __float_as_int(z);
  int row = (bits >> 19) - ((127 + COFLUX_PSI_KMIN) << 4);
  row = max(0, min(row, COFLUX_NROWS - 1));
  const float one_plus_u = __int_as_float(((bits & 0x0007ffff) << 4) | 0x3f800000);
  t = fm::fma_(2.0f, one_plus_u, -3.0f);
  return row;
}
template <typename FT> struct Coef8 { FT c[8]; };
__device__ __forceinline__ Coef8<double> ld_coef8(const double* p) {       // 2 × LDG.E.256 (sm_100): half the L1 requests of 4 × LDG.128
  Coef8<double> k;
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(k.c[0]), "=d"(k.c[1]), "=d"(k.c[2]), "=d"(k.c[3]) : "l"(p));
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(k.c[4]), "=d"(k.c[5]), "=d"(k.c[6]), "=d"(k.c[7]) : "l"(p + 4));
  return k;
}
__device__ __forceinline__ Coef8<float> ld_coef8(const float* p) {         // 2 × LDG.128
  const float4* q = reinterpret_cast<const float4*>(p);
  const float4 a = __ldg(q), b = __ldg(q + 1);
  return Coef8<float>{{a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w}};
}
__device__ __forceinline__ double2 lds_f64x2(unsigned addr) {
  double2 v; asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr)); return v;
}
__device__ __forceinline__ float4 lds_f32x4(unsigned addr) {
  float4 v; asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr)); return v;
}
// Swizzled shared-memory address of piece 0 of local row rl; piece p is at (that address) ^ (p << 4).
//   Float64: 128-byte rows of 8 pieces (ψ_u: 0–3, ψ_θ: 4–7), swizzle rl & 7.
//   Float32:  64-byte rows of 4 pieces (ψ_u: 0–1, ψ_θ: 2–3), swizzle (rl >> 1) & 3 (bit 2 of the bank group is rl & 1).
__device__ __forceinline__ unsigned psi_sm_addr(unsigned base, unsigned rl, double) { return base + (rl << 7) + ((rl & 7u) << 4); }
__device__ __forceinline__ unsigned psi_sm_addr(unsigned base, unsigned rl, float) { return base + (rl << 6) + (((rl >> 1) & 3u) << 4); }
// coefficients of ψ_u (which = 0) or ψ_θ (which = 1) of table row `row`
__device__ __forceinline__ Coef8<double> ld_psi(const LeanTabs& tb, int row, int which, double tag) {
  const unsigned rl = (unsigned)(row - COFLUX_PSI_SM_ROW0);
  if (__builtin_expect(rl < (unsigned)COFLUX_PSI_SM_ROWS, 1)) {
    const unsigned a0 = psi_sm_addr(tb.psi, rl, tag) ^ (which ? 64u : 0u);
    const double2 a = lds_f64x2(a0), b = lds_f64x2(a0 ^ 16u), c = lds_f64x2(a0 ^ 32u), d = lds_f64x2(a0 ^ 48u);
    return Coef8<double>{{a.x, a.y, b.x, b.y, c.x, c.y, d.x, d.y}};
  }
  return ld_coef8(&COFLUX_PSI_TABLE_F64[row][which][0]);
}
__device__ __forceinline__ Coef8<float> ld_psi(const LeanTabs& tb, int row, int which, float tag) {
  const unsigned rl = (unsigned)(row - COFLUX_PSI_SM_ROW0);
  if (__builtin_expect(rl < (unsigned)COFLUX_PSI_SM_ROWS, 1)) {
    const unsigned a0 = psi_sm_addr(tb.psi, rl, tag) ^ (which ? 32u : 0u);
    const float4 a = lds_f32x4(a0), b = lds_f32x4(a0 ^ 16u);
    return Coef8<float>{{a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w}};
  }
  return ld_coef8(&COFLUX_PSI_TABLE_F32[row][which][0]);
}
// both functions of one row (the ψ(h/L★) pair): one range test for the 8 (4) loads
__device__ __forceinline__ void ld_psi_pair(const LeanTabs& tb, int row, Coef8<double>& km, Coef8<double>& ks, double tag) {
  const unsigned rl = (unsigned)(row - COFLUX_PSI_SM_ROW0);
  if (__builtin_expect(rl < (unsigned)COFLUX_PSI_SM_ROWS, 1)) {
    const unsigned a0 = psi_sm_addr(tb.psi, rl, tag);
    const double2 a = lds_f64x2(a0), b = lds_f64x2(a0 ^ 16u), c = lds_f64x2(a0 ^ 32u), d = lds_f64x2(a0 ^ 48u);
    const double2 e = lds_f64x2(a0 ^ 64u), f = lds_f64x2(a0 ^ 80u), g = lds_f64x2(a0 ^ 96u), h = lds_f64x2(a0 ^ 112u);
    km = Coef8<double>{{a.x, a.y, b.x, b.y, c.x, c.y, d.x, d.y}};
    ks = Coef8<double>{{e.x, e.y, f.x, f.y, g.x, g.y, h.x, h.y}};
  } else {
    km = ld_coef8(&COFLUX_PSI_TABLE_F64[row][0][0]); ks = ld_coef8(&COFLUX_PSI_TABLE_F64[row][1][0]);
  }
}
__device__ __forceinline__ void ld_psi_pair(const LeanTabs& tb, int row, Coef8<float>& km, Coef8<float>& ks, float tag) {
  const unsigned rl = (unsigned)(row - COFLUX_PSI_SM_ROW0);
  if (__builtin_expect(rl < (unsigned)COFLUX_PSI_SM_ROWS, 1)) {
    const unsigned a0 = psi_sm_addr(tb.psi, rl, tag);
    const float4 a = lds_f32x4(a0), b = lds_f32x4(a0 ^ 16u), c = lds_f32x4(a0 ^ 32u), d = lds_f32x4(a0 ^ 48u);
    km = Coef8<float>{{a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w}};
    ks = Coef8<float>{{c.x, c.y, c.z, c.w, d.x, d.y, d.z, d.w}};
  } else {
    km = ld_coef8(&COFLUX_PSI_TABLE_F32[row][0][0]); ks = ld_coef8(&COFLUX_PSI_TABLE_F32[row][1][0]);
  }
}
template <typename FT> __device__ __forceinline__ FT poly8v(const Coef8<FT>& k, FT t) {
  FT a = fm::fma_(k.c[7], t, k.c[6]);
  a = fm::fma_(a, t, k.c[5]); a = fm::fma_(a, t, k.c[4]);
  a = fm::fma_(a, t, k.c[3]); a = fm::fma_(a, t, k.c[2]);
  a = fm::fma_(a, t, k.c[1]);
  return fm::fma_(a, t, k.c[0]);
}
// stable closed forms (Edson et al. 2013), any z ≥ 0;  e = exp(−min(50, 0.35 z))
template <typename FT> __device__ __forceinline__ FT psi_stable_m(FT z, FT e) {
  const LeanLit<FT>& LL = lean_lit<FT>();
  return fm::fma_(FT(-0.75) * (z - LL.c5_035), e, fm::fma_(LL.m07, z, LL.m075c));
}
template <typename FT> __device__ __forceinline__ FT psi_stable_s(FT z, FT e) {
  const LeanLit<FT>& LL = lean_lit<FT>();
  const FT w = fm::fma_(LL.two3, z, FT(1));
  return fm::fma_(LL.m23 * (z - LL.c1428), e, fm::fma_(-w, fm::sqrt(w), LL.m8525));
}

// min / max by compare-select (3 instructions; fmin/fmax cost 5 with their NaN handling — the second operand
// is a finite parameter everywhere below, so a NaN first operand still yields the parameter)
template <typename FT> __device__ __forceinline__ FT dmin_(FT a, FT b) { return (a < b) ? a : b; }
template <typename FT> __device__ __forceinline__ FT dmax_(FT a, FT b) { return (a > b) ? a : b; }

// One pass.  Returns false when the pass left the short path (the scales are then unchanged and the
// caller must run lean_cold_pass).  Float32 runs the same pass with the CUDA single-precision functions
// (fm:: overloads), the Float32 ψ table and the same series.
// FIRST: the pass that starts from the initial guess u★ = θ★ = q★ = init > 0.  Its buoyancy scale is positive whatever
// the cell (b★ = init·(cb1 + cb2), both coefficients positive), so only the stable block is compiled; phase A runs it
// for every cell in lock step.  chi_out (optional) receives χ_q: θ★ = χ_q·Δθ and q★ = χ_q·Δq are then reproducible
// from ONE stored number.
template <typename FT, int SPEC, bool FIRST = false>
__device__ __forceinline__ bool iterate_lean(const DevParams<FT>& P, const FluxP<FT>& F, const FastConsts<FT>& K,
                                             const LeanTabs& tb, const LeanCell<FT>& c, FT& us, FT& ts, FT& qs, FT* chi_out = nullptr) {
  const LeanLit<FT>& LL = lean_lit<FT>();
  constexpr FT TINY = FT(0.0001220703125);                   // 2⁻¹³: |ℓ/L★| below which the short series are exact to rounding
  constexpr FT Z_LO = FT(9.5367431640625e-07), Z_HI = FT(8192.0);   // table domain of −ζ: [2⁻²⁰, 2¹³)
  constexpr FT SMALL = FT(7.888609052210118e-31), BIG = FT(1.2676506002282294e30);   // 2⁻¹⁰⁰, 2¹⁰⁰
  const FT u0 = us, t0 = ts, q0 = qs;
  const FT bstar = fm::fma_(c.cb1, t0, c.cb2 * q0);
  const bool unstable = FIRST ? false : (bstar < FT(0));
  bool ok = (u0 > SMALL) && (c.Ustab > FT(0));
  const FT r = fm::rcp(u0);
  const FT invL = (F.kappa * bstar) * (r * r);
  const FT zeta = P.h * invL;
  // SPEC 1 (constant Charnock): the roughness lengths do not depend on this pass's wind speed — evaluate them
  // first, so that ONE block per stability class holds everything that depends on the sign of ζ
  FT lu = 0, ll = 0, lnq = K.lnhl_q, lq = F.qr.lmax, xm = 0, xs = 0;
  auto roughness = [&](FT alpha_g) {
    lu = dmin_<FT>(fm::fma_(alpha_g * u0, u0, c.bnu * r), F.mr.lmax);
    ll = fm::log(lu, tb.lg);
    const FT lr = fm::log((lu * u0) * c.inv_nu, tb.lg);      // ln R★
    if (lr > K.lrclip_q) {                                   // A·R★^(−b) < ℓ_max
      lnq = fm::fma_(F.qr.b, lr, K.lnhA_q);
      if (SPEC == 1) lq = F.qr.A * fm::exp(-F.qr.b * lr, tb.ex);
    }
    xm = lu * invL; xs = lq * invL;                          // ℓ/L★ (same sign as ζ)
  };
  if (SPEC == 1) roughness(K.alpha_g);
  FT U = c.Ustab;                                            // √(Δu² + U_G,min²): no gustiness when Jᵇ ≤ 0
  FT psi_hm, psi_hs, sm_ = FT(0), ss_ = FT(0);
  if (unstable) {
    FT t;
    Coef8<FT> km, ks;
    ld_psi_pair(tb, psi_table_row(-zeta, t), km, ks, FT(0)); // loads first: the cube root below hides their latency
    const FT w = (-u0 * bstar) * P.hbl;                      // Jᵇ·h_bl > 0
    ok = ok && (w > SMALL) && (w < BIG) && (-zeta >= Z_LO) && (-zeta < Z_HI);
    const FT UG = dmax_<FT>(F.beta * fm::cbrt(w), F.ugmin);
    U = fm::sqrt(fm::fma_(UG, UG, c.U2));
    if (SPEC == 1) {                                         // ψ(ℓ/L★): short series near 0, the same table beyond
      ok = ok && (-xm < Z_HI) && (-xs < Z_HI);
      if (-xm <= TINY) {
        FT a = fm::fma_(FT(-12220.4140625), xm, FT(-1198.931640625));
        a = fm::fma_(a, xm, FT(-131.46923828125)); a = fm::fma_(a, xm, FT(-17.578125)); a = fm::fma_(a, xm, FT(-3.75));
        sm_ = a * xm;
      } else {
        FT tt;
        const int rr = psi_table_row(-xm, tt);
        sm_ = poly8v<FT>(ld_psi(tb, rr, 0, FT(0)), tt);
      }
      if (-xs <= TINY) {
        FT b = fm::fma_(FT(-39314.5625), xs, FT(-3548.0859375));
        b = fm::fma_(b, xs, LL.su3); b = fm::fma_(b, xs, FT(-42.1875)); b = fm::fma_(b, xs, FT(-7.5));
        ss_ = b * xs;
      } else {
        FT tt;
        const int rr = psi_table_row(-xs, tt);
        ss_ = poly8v<FT>(ld_psi(tb, rr, 1, FT(0)), tt);
      }
    }
    psi_hm = poly8v<FT>(km, t);
    psi_hs = poly8v<FT>(ks, t);
  } else {
    const FT e = fm::exp(-dmin_<FT>(LL.c035 * zeta, FT(50)), tb.ex);
    psi_hm = psi_stable_m<FT>(zeta, e);
    psi_hs = psi_stable_s<FT>(zeta, e);
    if (SPEC == 1) {
      if (xm <= TINY) sm_ = xm * fm::fma_(fm::fma_(FT(-0.12249755859375), xm, LL.ms2), xm, LL.ms1);
      else sm_ = psi_stable_m<FT>(xm, fm::exp(-dmin_<FT>(LL.c035 * xm, FT(50)), tb.ex));
      if (xs <= TINY) ss_ = fm::fma_(fm::fma_(fm::fma_(FT(-0.09034299850463867), xs, LL.ss2), xs, LL.ss1), xs, LL.ss0);
      else ss_ = psi_stable_s<FT>(xs, fm::exp(-dmin_<FT>(LL.c035 * xs, FT(50)), tb.ex));
    }
  }
  if (SPEC == 2) roughness(dmax_<FT>(fm::fma_(F.mr.a1, dmin_<FT>(U, F.mr.umax), F.mr.a2), F.mr.amin) * K.inv_g);
  if (!__builtin_expect(ok, 1)) return false;
  const FT prof_u = ((K.lnh - ll) - psi_hm) + sm_;
  const FT prof_q = (lnq - psi_hs) + ss_;
  if (!(prof_u > FT(0))) { us = ts = qs = FT(0); if (chi_out) *chi_out = FT(0); return true; }
  const FT chi_u = F.kappa * fm::rcp(prof_u);
  const FT chi_q = (prof_q > FT(0)) ? F.kappa * fm::rcp(prof_q) : FT(0);
  us = chi_u * U; ts = chi_q * c.dth; qs = chi_q * c.dq;
  if (chi_out) *chi_out = chi_q;
  return true;
}
// the exact pass behind one by-value call, so that the lean loop stays small and its state stays in registers
template <typename FT, int SPEC>
__device__ __noinline__ D3<FT> lean_cold_pass(const DevParams<FT>* P, FT U2, FT dth, FT dq, FT cb1, FT cb2, FT nu, FT us, FT ts, FT qs) {
  iterate_fast<FT, SPEC>(*P, P->ao, P->K, U2, dth, dq, FT(1), cb1, cb2, nu, us, ts, qs);
  return D3<FT>{us, ts, qs};
}
template <typename FT> __device__ __forceinline__ bool same_bits(FT a, FT b);
template <> __device__ __forceinline__ bool same_bits<double>(double a, double b) { return __double_as_longlong(a) == __double_as_longlong(b); }
template <> __device__ __forceinline__ bool same_bits<float>(float a, float b) { return __float_as_int(a) == __float_as_int(b); }

// Launch shape (round 2).  One copy of the hot ψ rows per CTA (26 KB in Float64) wants few, large CTAs; the phases of
// different CTAs overlap on an SM (A is latency bound on gathers, B on the FP64 / issue pipes), which wants several.
//   Float64 `:default`   320 threads × 2 CTAs/SM, ≤ 1280 cells per CTA (58 B/cell + 29 KB of tables = 103 KB), 96 registers:
//                        no spills, 4 cells per lane (3.20 ms at 1/12°; 256 × 3 × 768 at 80 registers with spills: 3.32 ms)
//   Float64 `:corrected` 384 threads × 2 CTAs/SM, 1152 cells (74 B/cell: ν and 1/ν vary), 80 registers
//   Float32              384 threads × 3 CTAs/SM, 1536 cells (30 / 38 B/cell + 13 KB table), 56 registers
// A/B on B200 at 1/12° (profiles/README.md): Float64 at 64 / 72 registers (32 / 28 warps per SM) is SLOWER (3.98 / 3.56 ms
// against 3.32 ms: the spills cost more than the extra warps hide); Float32 256 × 4 × 1024 2.40 ms, 384 × 3 × 1536 2.34 ms.
// Round-1 shape for reference (table in global memory): 128 threads × 6 CTAs, 384 cells (profiles/README.md).
#ifndef COFLUX_TILE_CARRY2
#define COFLUX_TILE_CARRY2 0     /* 1: Δu, Δv, ρ_a, c_p,m of every cell stay in shared memory between phase A and phase C (32 B per cell more)
                                   instead of being re-read / parked in the ρτ output arrays (A/B knob) */
#endif
#ifndef COFLUX_TILE_PRE1
#define COFLUX_TILE_PRE1 1     /* first pass in lock step in phase A; 0: every pass in phase B and 16 B per queued cell less (A/B knob) */
#endif
#ifndef COFLUX_TILE_C2CONST
#define COFLUX_TILE_C2CONST 0  /* lean pass: the q★ coefficient of b★ is g·δ for every cell (it is (g/T_v)·(δ·T_v) up to rounding): 8 B per
                                  queued cell less (A/B knob) */
#endif
#ifndef COFLUX_TILE_NT64
#define COFLUX_TILE_NT64 320
#endif
#ifndef COFLUX_TILE_CELLS64
#define COFLUX_TILE_CELLS64 1280
#endif
#ifndef COFLUX_TILE_MIN_BLOCKS64
#define COFLUX_TILE_MIN_BLOCKS64 2
#endif
#ifndef COFLUX_TILE_NT64_S2
#define COFLUX_TILE_NT64_S2 384
#endif
#ifndef COFLUX_TILE_CELLS64_S2
#define COFLUX_TILE_CELLS64_S2 1152
#endif
#ifndef COFLUX_TILE_MIN_BLOCKS64_S2
#define COFLUX_TILE_MIN_BLOCKS64_S2 2
#endif
#ifndef COFLUX_TILE_NT32
#define COFLUX_TILE_NT32 384
#endif
#ifndef COFLUX_TILE_CELLS32
#define COFLUX_TILE_CELLS32 1536
#endif
#ifndef COFLUX_TILE_MIN_BLOCKS32
#define COFLUX_TILE_MIN_BLOCKS32 3
#endif

// shared-memory layout of one tile (SoA: consecutive lanes touch consecutive words — no bank conflicts).
// A cell's slots are reused over its life:
//   queued      U2, dth, dq, c1, c2 (+ nu, inu): invariants of the iteration (lean pass: c1, c2 = the coefficients of
//               b★ = c1·θ★ + c2·q★; generic pass: T_v, q_v);  us1, chi1: u★ and χ_q after the first pass, which phase A
//               runs in lock step (us1 < 0: not run, the cell starts from the initial guess)
//   in flight   the lane holds the invariants in registers; U2/dth/dq hold the Brent snapshot, c1 the packed Brent state
//   finished    U2, dth, dq ← u★, θ★, q★;  c1 ← iteration count (as an integer)
// ρ_a and c_p,m — needed again by phase C — are parked in the ρτx / ρτy OUTPUT arrays by phase A (same thread, same
// element; the lines are still in L2 when phase C overwrites them with the stresses), not in shared memory: 16 B per
// cell less, i.e. more cells per lane for the same footprint.
template <typename FT, int TILE, bool VARNU, bool LEAN> struct TileSmem {
  FT U2[TILE], dth[TILE], dq[TILE], c1[TILE], c2[(LEAN && COFLUX_TILE_C2CONST) ? 1 : TILE];
  FT nu[VARNU ? TILE : 1];                                // air viscosity at T_s (only when it varies)
  FT inu[(LEAN && VARNU) ? TILE : 1];                     // 1/ν (lean pass)
  FT us1[(LEAN && COFLUX_TILE_PRE1) ? TILE : 1], chi1[(LEAN && COFLUX_TILE_PRE1) ? TILE : 1];   // state after the lock-step first pass (lean pass)
  FT du[COFLUX_TILE_CARRY2 ? TILE : 1], dv[COFLUX_TILE_CARRY2 ? TILE : 1], rho[COFLUX_TILE_CARRY2 ? TILE : 1], cp[COFLUX_TILE_CARRY2 ? TILE : 1];
  unsigned short queue[TILE];
  int n_front, n_back, head[2];     // head[0]: next unstable cell (front of the queue), head[1]: next stable cell (back)
};
template <typename FT> struct SlotInt;      // integer view of a slot
template <> struct SlotInt<double> {
  static __device__ __forceinline__ int get(const double& s) { return (int)__double_as_longlong(s); }
  static __device__ __forceinline__ void set(double& s, int v) { s = __longlong_as_double((long long)v); }
};
template <> struct SlotInt<float> {
  static __device__ __forceinline__ int get(const float& s) { return __float_as_int(s); }
  static __device__ __forceinline__ void set(float& s, int v) { s = __int_as_float(v); }
};
template <typename FT, int SPEC> struct TileTraits {
  static constexpr bool F64 = (sizeof(FT) == 8);
  static constexpr bool VARNU = (SPEC != 1);   // `:default` uses a constant air viscosity
  static constexpr bool LEAN = (COFLUX_LEAN != 0) && (SPEC != 0) && (F64 || (COFLUX_LEAN_F32 != 0));
  static constexpr bool TABS = LEAN && F64;    // log / exp tables in shared memory
  static constexpr int NT = F64 ? ((SPEC == 1) ? COFLUX_TILE_NT64 : COFLUX_TILE_NT64_S2) : COFLUX_TILE_NT32;
  static constexpr int TILE = F64 ? ((SPEC == 1) ? COFLUX_TILE_CELLS64 : COFLUX_TILE_CELLS64_S2) : COFLUX_TILE_CELLS32;
  static constexpr int MIN_BLOCKS = F64 ? ((SPEC == 1) ? COFLUX_TILE_MIN_BLOCKS64 : COFLUX_TILE_MIN_BLOCKS64_S2) : COFLUX_TILE_MIN_BLOCKS32;
  static constexpr int PSI_BYTES = LEAN ? COFLUX_PSI_SM_ROWS * 16 * (int)sizeof(FT) : 16;
  static_assert(TILE <= 65535, "queue entries are 16-bit");
};

// COFLUX_TILE_MAXNREG (A/B knob): state the register budget directly — ptxas derives it from __launch_bounds__ with the CTA size
// rounded up to a multiple of 64 threads, which wastes the headroom of 224- or 352-thread CTAs
#ifdef COFLUX_TILE_MAXNREG
#define COFLUX_TILE_BOUNDS(FT, SPEC) __maxnreg__(COFLUX_TILE_MAXNREG)
#else
#define COFLUX_TILE_BOUNDS(FT, SPEC) __launch_bounds__(TileTraits<FT, SPEC>::NT, TileTraits<FT, SPEC>::MIN_BLOCKS)
#endif
template <typename FT, bool INTERP, bool ASSEMBLE, int TILE, int SPEC>
__global__ void COFLUX_TILE_BOUNDS(FT, SPEC) flux_tile_kernel(const __grid_constant__ FluxArgs<FT> a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  using TT = TileTraits<FT, SPEC>;
  constexpr bool VARNU = TT::VARNU;
  constexpr bool LEAN = TT::LEAN;
  constexpr bool TABS = TT::TABS;
  constexpr int NT = TT::NT;
  using MP = std::conditional_t<TABS, MLeanD, M<FT>>;      // math policy of the phase-A thermodynamics
  using SI = SlotInt<FT>;
  TileSmem<FT, TILE, VARNU, LEAN>& sm = *reinterpret_cast<TileSmem<FT, TILE, VARNU, LEAN>*>(smem_raw);
  // log / exp tables of the lean Float64 functions and the hot ψ rows: STATIC shared arrays, so that their addresses are
  // compile-time shared-window offsets (through the dynamic block every lookup paid a generic→shared address conversion)
  __shared__ __align__(16) double s_lgt[TABS ? 256 : 2];
  __shared__ double s_ext[TABS ? 64 : 2];
  __shared__ __align__(128) unsigned char s_psi[TT::PSI_BYTES];
  const DevParams<FT>& P = a.P;
  const FluxP<FT>& F = P.ao;
  const ThermoC<FT>& c = P.th;
  const int tid = threadIdx.x;
  // balanced tiling (FluxArgs::tile_cells): this CTA's cells are [tile0, tile0 + tile_n)
  // balanced tiling (FluxArgs::tile_cells): this CTA's cells are [tile0, tile0 + tile_n)
  const int tile_n = (a.tile_cells > 0 && a.tile_cells < TILE) ? a.tile_cells : TILE;
  const long long tile0 = a.cell0 + (long long)blockIdx.x * tile_n;
  if (tid == 0) { sm.n_front = 0; sm.n_back = 0; sm.head[0] = 0; sm.head[1] = 0; }
  if (TABS) {
    for (int k = tid; k < 256; k += NT) s_lgt[k] = (&COFLUX_LOG_TABLE[0][0])[k];
    if (tid < 64) s_ext[tid] = COFLUX_EXP_TABLE[tid];
  }
  if constexpr (LEAN) {
    if constexpr (sizeof(FT) == 8) {     // 8 pieces of 16 B per row, piece p at position p ^ (r & 7)
      const double2* src = reinterpret_cast<const double2*>(&COFLUX_PSI_TABLE_F64[COFLUX_PSI_SM_ROW0][0][0]);
      double2* dst = reinterpret_cast<double2*>(s_psi);
      for (int k = tid; k < COFLUX_PSI_SM_ROWS * 8; k += NT) { const int r = k >> 3, p = k & 7; dst[(r << 3) + (p ^ (r & 7))] = __ldg(src + k); }
    } else {                             // 4 pieces per row, piece p at position p ^ ((r >> 1) & 3)
      const float4* src = reinterpret_cast<const float4*>(&COFLUX_PSI_TABLE_F32[COFLUX_PSI_SM_ROW0][0][0]);
      float4* dst = reinterpret_cast<float4*>(s_psi);
      for (int k = tid; k < COFLUX_PSI_SM_ROWS * 4; k += NT) { const int r = k >> 2, p = k & 3; dst[(r << 2) + (p ^ ((r >> 1) & 3))] = __ldg(src + k); }
    }
  }
  __syncthreads();
  const LeanTabs tb{s_lgt, s_ext, (unsigned)__cvta_generic_to_shared(s_psi)};
  const FastConsts<FT>& K = P.K;
  const FT delta = c.eps - FT(1);
  const bool fixed = (F.stop_kind == COFLUX_STOP_FIXED_ITERATIONS);

  // (ii, jj) of the tile's first cell: ONE 64-bit division per thread; a cell's own row / column follow from a 32-bit one
  const int tile_jj0 = (int)(tile0 / a.nxr);
  const int tile_ii0 = (int)(tile0 - (long long)tile_jj0 * a.nxr);
  const unsigned nxr_u = (unsigned)a.nxr;

  // ------------------------------------------------------------------ phase A
  for (int cidx = tid; cidx < tile_n; cidx += NT) {
    const long long idx = tile0 + cidx;
    if (idx >= a.ncell) continue;
    const unsigned t = (unsigned)(tile_ii0 + cidx), dj = t / nxr_u;
    const int jj = tile_jj0 + (int)dj, ii = (int)(t - dj * nxr_u);
    const int i = ii - a.ring, j = jj - a.ring;
    FT ua, va, Ta, pa, qa;
    const int off = j * a.usj + i;          // element offset of the cell in every 2-D surface array (uniform layout)
    auto LD = [&](const DArr& d, int o) -> FT { return __ldg(reinterpret_cast<const FT*>(d.p) + o); };
    auto ST = [&](const DArr& d, FT v) { if (d.p) reinterpret_cast<FT*>(d.p)[off] = v; };
    if (INTERP) {
      const int foff = j * a.fsj + i;
      const FT fi = LD(a.fi, foff), fj = LD(a.fj, foff);
      const int i0 = (int)M<FT>::trunc(fi), j0 = (int)M<FT>::trunc(fj);
      const int i1 = i0 + ((fi > FT(0)) - (fi < FT(0))), j1 = j0 + ((fj > FT(0)) - (fj < FT(0)));
      const FT xi = fi - M<FT>::floor(fi), eta = fj - M<FT>::floor(fj);
      const FT w00 = (FT(1) - xi) * (FT(1) - eta), w01 = (FT(1) - xi) * eta, w10 = xi * (FT(1) - eta), w11 = xi * eta;
      const int o00 = j0 * a.ssj + i0, o01 = j1 * a.ssj + i0, o10 = j0 * a.ssj + i1, o11 = j1 * a.ssj + i1;   // one set for all series
      ua = interp_series_u<FT>(a.su, o00, o01, o10, o11, w00, w01, w10, w11, a.nfrac);
      va = interp_series_u<FT>(a.sv, o00, o01, o10, o11, w00, w01, w10, w11, a.nfrac);
      Ta = interp_series_u<FT>(a.sT, o00, o01, o10, o11, w00, w01, w10, w11, a.nfrac);
      qa = interp_series_u<FT>(a.sq, o00, o01, o10, o11, w00, w01, w10, w11, a.nfrac);
      pa = interp_series_u<FT>(a.sp, o00, o01, o10, o11, w00, w01, w10, w11, a.nfrac);
      const FT Qs = interp_series_u<FT>(a.sQs, o00, o01, o10, o11, w00, w01, w10, w11, a.nfrac);
      const FT Ql = interp_series_u<FT>(a.sQl, o00, o01, o10, o11, w00, w01, w10, w11, a.nfrac);
      FT Mp = FT(0);
      if (a.srain.p1) Mp += interp_series_u<FT>(a.srain, o00, o01, o10, o11, w00, w01, w10, w11, a.nfrac);
      if (a.ssnow.p1) Mp += interp_series_u<FT>(a.ssnow, o00, o01, o10, o11, w00, w01, w10, w11, a.nfrac);
      if (a.lfi.p) Mp += land_freshwater<FT>(a, i, j);
      if (a.cs.p && a.sn.p) {
        const FT cs = LD(a.cs, foff), sn = LD(a.sn, foff);
        const FT ur = ua * cs + va * sn, vr = -ua * sn + va * cs;
        ua = ur; va = vr;
      }
      ST(a.xu, ua); ST(a.xv, va); ST(a.xT, Ta); ST(a.xp, pa);
      ST(a.xq, qa); ST(a.xQs, Qs); ST(a.xQl, Ql); ST(a.xMp, Mp);
    } else {
      ua = LD(a.xu, off); va = LD(a.xv, off); Ta = LD(a.xT, off); pa = LD(a.xp, off);
      qa = LD(a.xq, off);
    }
    bool queued = false, finished_in_a = false;
    const bool wet = !a.mask.p || __ldg(reinterpret_cast<const uint8_t*>(a.mask.p) + off) != 0;
    if (wet) {
      const FT uo = (LD(a.ou, off) + LD(a.ou, off + 1)) * FT(0.5);
      const FT vo = (LD(a.ov, off) + LD(a.ov, off + a.usj)) * FT(0.5);
      const FT Ts = LD(a.oT, off) + P.T_offset;
      const FT So = LD(a.oS, off);
      FT du, dv;
      if (F.velocity == COFLUX_VELOCITY_RELATIVE) { du = ua - uo; dv = va - vo; } else { du = ua; dv = va; }
      const FT U2 = du * du + dv * dv;
      const Thermo<FT> atm = phase_equil_pTq<FT, MP>(c, pa, Ta, qa);
      const FT s = MP::div(So, FT(1000));
      const FT x = MP::div(FT(1) - s, FT(1) - s + P.wmf_alpha * s);
      const FT theta_a = Ta + MP::div(P.g * P.h, atm.cp_m);
      const SurfaceState<FT> S = surface_state<FT, 0, MP>(P, F, atm, pa, theta_a, x, Ts);
#if COFLUX_TILE_CARRY2
      sm.du[cidx] = du; sm.dv[cidx] = dv; sm.rho[cidx] = atm.rho; sm.cp[cidx] = atm.cp_m;
#else
      ST(a.rtx, atm.rho); ST(a.rty, atm.cp_m);     // parked for phase C (see TileSmem)
#endif
      if (fixed ? (F.maxit > 0) : true) {
        queued = true;
        FT c1, c2;
        if constexpr (LEAN) {    // b★ = c1·θ★ + c2·q★
          const FT gTv = P.g * fm::rcp(S.T_v);
          c1 = gTv * (FT(1) + delta * S.q_vap); c2 = COFLUX_TILE_C2CONST ? P.g * delta : gTv * (delta * S.T_v);
          const FT inv_nu = VARNU ? fm::rcp(S.nu_m) : K.inv_nu;
          if (VARNU) sm.inu[cidx] = inv_nu;
          // first pass, in lock step (every lane busy, one code path)
          FT us = F.init, ts = F.init, qs = F.init, chi = FT(0);
          bool pre = COFLUX_TILE_PRE1 && F.init > FT(0) && (c1 + c2) > FT(0);
          if (pre) {
            LeanCell<FT> lc;
            lc.U2 = U2; lc.dth = S.dtheta; lc.dq = S.dq; lc.cb1 = c1; lc.cb2 = c2;
            { const FT v = fm::fma_(F.ugmin, F.ugmin, U2); lc.Ustab = (v > FT(0)) ? fm::sqrt(v) : FT(0); }
            lc.bnu = VARNU ? F.mr.beta_s * S.nu_m : K.bnu; lc.inv_nu = inv_nu;
            pre = iterate_lean<FT, SPEC, true>(P, F, K, tb, lc, us, ts, qs, &chi);
          }
          if (pre && !keep_going<FT>(F, 1, us, ts, qs, F.init, F.init, F.init)) {     // done after one pass
            queued = false;
            sm.U2[cidx] = us; sm.dth[cidx] = ts; sm.dq[cidx] = qs; SI::set(sm.c1[cidx], 1);
          }
          if constexpr (COFLUX_TILE_PRE1 != 0) { sm.us1[cidx] = pre ? us : FT(-1); sm.chi1[cidx] = chi; }
        } else { c1 = S.T_v; c2 = S.q_vap; }
        if (queued) {
          sm.U2[cidx] = U2; sm.dth[cidx] = S.dtheta; sm.dq[cidx] = S.dq; sm.c1[cidx] = c1;
          if constexpr (!(LEAN && COFLUX_TILE_C2CONST)) sm.c2[cidx] = c2;
          if (VARNU) sm.nu[cidx] = S.nu_m;
          // stability class of every later pass: sign of the buoyancy scale ∝ Δθ·a1 + a2·Δq (χ_θ = χ_q > 0)
          const bool unstable = LEAN ? ((S.dtheta * c1 + c2 * S.dq) < FT(0)) : ((S.dtheta * (FT(1) + delta * S.q_vap) + (delta * S.T_v) * S.dq) < FT(0));
          if (unstable) sm.queue[atomicAdd(&sm.n_front, 1)] = (unsigned short)cidx;
          else sm.queue[TILE - 1 - atomicAdd(&sm.n_back, 1)] = (unsigned short)cidx;
        }
        finished_in_a = !queued;
      }
    }
    if (!queued && !finished_in_a) {      // finished on the spot: land, or a zero-pass solve
      const FT r0 = wet ? F.init : FT(0);
      sm.U2[cidx] = r0; sm.dth[cidx] = r0; sm.dq[cidx] = r0; SI::set(sm.c1[cidx], 0);
    }
  }
  __syncthreads();

  const int n_front = sm.n_front, n_back = sm.n_back;

  // ------------------------------------------------------------------ phase B: lane refill, class-pure warps
  // The two stability classes run different code (unstable: cube root + ψ table; stable: closed forms).  With ONE queue
  // (unstable cells first) the lanes of every warp drift into the stable cells one by one, and for the last third of
  // the phase every warp executes BOTH blocks per pass (ncu, round 2: the shared tail of the pass ran 1.49× per loop
  // trip).  So each warp serves ONE class: warps [0, W_u) pop unstable cells from the front of the queue, the others
  // stable cells from its back, W_u proportional to the classes' expected work (stable cells need ≈ 15 % more passes).
  // A warp whose class has run dry AND whose lanes have all finished moves over to help with the other class.
  constexpr int NW = NT / 32;
  int cls;                                       // 0 unstable, 1 stable
  {
    const float wu = 13.9f * (float)n_front, ws = 15.9f * (float)n_back;
    int W_u = (wu + ws > 0.f) ? (int)((float)NW * wu / (wu + ws) + 0.5f) : NW;
    if (n_front > 0 && W_u < 1) W_u = 1;
    if (n_back > 0 && W_u > NW - 1) W_u = NW - 1;
    if (n_front == 0) W_u = 0;
    cls = ((tid >> 5) < W_u) ? 0 : 1;
  }
  bool moved = false;
  // Brent cycle detection: (su, st, sq) is a snapshot of the iterate taken at pass `snap_it`; it is
  // refreshed after 1, 2, 4, 8 … passes.  When the iterate returns EXACTLY to the snapshot the orbit
  // is periodic with period λ = it − snap_it; the reference keeps iterating until maxiter, i.e. it
  // ends (maxiter − it) mod λ passes further along the same orbit — run just those and stop.
  if constexpr (LEAN) {
    constexpr int BRENT_FROM = (sizeof(FT) == 8) ? 24 : 6;   // cycle detection starts here (Float64 converges in < 30 passes;
                                                             // in Float32 8 % of the cells only ever reach a limit cycle)
    int slot = -1, it = 0;
    LeanCell<FT> lc{};
    FT nu = F.mr.visc.nu, us = 0, ts = 0, qs = 0;
    if (!VARNU) { lc.bnu = K.bnu; lc.inv_nu = K.inv_nu; }
    auto pop = [&]() {
      const int pos = atomicAdd(&sm.head[cls], 1);
      slot = -1;
      if (pos < (cls ? n_back : n_front)) {
        slot = cls ? sm.queue[TILE - 1 - pos] : sm.queue[pos];
        lc.U2 = sm.U2[slot]; lc.dth = sm.dth[slot]; lc.dq = sm.dq[slot]; lc.cb1 = sm.c1[slot];
        if constexpr (COFLUX_TILE_C2CONST != 0) lc.cb2 = P.g * delta; else lc.cb2 = sm.c2[slot];
        { const FT v = fm::fma_(F.ugmin, F.ugmin, lc.U2); lc.Ustab = (v > FT(0)) ? fm::sqrt(v) : FT(0); }
        if (VARNU) { nu = sm.nu[slot]; lc.inv_nu = sm.inu[slot]; lc.bnu = F.mr.beta_s * nu; }
        if constexpr (COFLUX_TILE_PRE1 != 0) {
          us = sm.us1[slot];
          if (us >= FT(0)) { const FT chi = sm.chi1[slot]; ts = chi * lc.dth; qs = chi * lc.dq; it = 1; }
          else { us = ts = qs = F.init; it = 0; }
        } else { us = ts = qs = F.init; it = 0; }
      }
    };
    // rare tail of a cell's iteration: the Brent bookkeeping lives in the cell's own shared-memory slots
    // (U2/dth/dq: snapshot; c1: snap_it | window << 8 | (stop_at + 1) << 16), not in registers
    auto brent = [&](bool go) -> bool {
      if (it == BRENT_FROM) {
        sm.U2[slot] = us; sm.dth[slot] = ts; sm.dq[slot] = qs;
        SI::set(sm.c1[slot], it | (1 << 8));
        return go;
      }
      const int packed = SI::get(sm.c1[slot]);
      const int snap_it = packed & 0xff, window = (packed >> 8) & 0xff, stop_at = (packed >> 16) - 1;
      if (stop_at >= 0) {                         // finishing a detected cycle
        if (it < stop_at) return true;
        it = F.maxit;
        return false;
      }
      if (!go) return false;
      if (same_bits<FT>(us, sm.U2[slot]) && same_bits<FT>(ts, sm.dth[slot]) && same_bits<FT>(qs, sm.dq[slot])) {
        const int lambda = it - snap_it;
        const int stop = it + (F.maxit - it) % lambda;
        SI::set(sm.c1[slot], packed | ((stop + 1) << 16));
        if (it < stop) return true;
        it = F.maxit;
        return false;
      }
      if (it - snap_it == window) {
        sm.U2[slot] = us; sm.dth[slot] = ts; sm.dq[slot] = qs;
        SI::set(sm.c1[slot], it | ((window * 2) << 8));
      }
      return true;
    };
    pop();
    for (;;) {
      if (!__any_sync(0xffffffffu, slot >= 0)) {       // the whole warp is idle: help with the other class, once
        if (moved) break;
        moved = true; cls ^= 1;
        pop();
        continue;
      }
      if (slot >= 0) {
        const FT u0 = us, t0 = ts, q0 = qs;
        if (__builtin_expect(!iterate_lean<FT, SPEC>(P, F, K, tb, lc, us, ts, qs), 0)) {
          const D3<FT> r = lean_cold_pass<FT, SPEC>(&P, lc.U2, lc.dth, lc.dq, lc.cb1, lc.cb2, nu, u0, t0, q0);
          us = r.u; ts = r.t; qs = r.q;
        }
        ++it;
        bool go = keep_going<FT>(F, it, us, ts, qs, u0, t0, q0);
        if (__builtin_expect(it >= BRENT_FROM && !fixed && F.maxit < 250, 0)) go = brent(go);   // (8-bit fields)
        if (!go) {
          sm.U2[slot] = us; sm.dth[slot] = ts; sm.dq[slot] = qs; SI::set(sm.c1[slot], it);
          pop();
        }
      }
    }
  } else {
    int slot = -1, it = 0;
    FT U2 = 0, dth = 0, dq = 0, gTv = 0, a1 = 0, a2 = 0, nu = 0, us = 0, ts = 0, qs = 0;
    FT su = 0, st = 0, sq = 0;
    int snap_it = 0, window = 1, stop_at = 0;
    auto pop = [&]() {
      const int pos = atomicAdd(&sm.head[cls], 1);
      slot = -1;
      if (pos < (cls ? n_back : n_front)) {
        slot = cls ? sm.queue[TILE - 1 - pos] : sm.queue[pos];
        U2 = sm.U2[slot]; dth = sm.dth[slot]; dq = sm.dq[slot];
        { const FT Tv = sm.c1[slot], qv = sm.c2[slot]; gTv = P.g / Tv; a1 = FT(1) + delta * qv; a2 = delta * Tv; }
        nu = VARNU ? sm.nu[slot] : F.mr.visc.nu;
        us = ts = qs = F.init; it = 0;
        su = us; st = ts; sq = qs; snap_it = it; window = 1; stop_at = -1;
      }
    };
    pop();
    for (;;) {
      if (!__any_sync(0xffffffffu, slot >= 0)) {       // the whole warp is idle: help with the other class, once
        if (moved) break;
        moved = true; cls ^= 1;
        pop();
        continue;
      }
      if (slot >= 0) {
        const FT u0 = us, t0 = ts, q0 = qs;
        iterate_fast<FT, SPEC>(P, F, K, U2, dth, dq, gTv, a1, a2, nu, us, ts, qs);
        ++it;
        bool go;
        if (stop_at >= 0) {                       // finishing a detected cycle
          go = it < stop_at;
          if (!go) it = F.maxit;
        } else {
          go = keep_going<FT>(F, it, us, ts, qs, u0, t0, q0);
          if (go && !fixed) {
            if (us == su && ts == st && qs == sq) {
              const int lambda = it - snap_it;
              stop_at = it + (F.maxit - it) % lambda;
              go = it < stop_at;
              if (!go) it = F.maxit;
            } else if (it - snap_it == window) {
              su = us; st = ts; sq = qs; snap_it = it; window *= 2;
            }
          }
        }
        if (!go) {
          sm.U2[slot] = us; sm.dth[slot] = ts; sm.dq[slot] = qs; SI::set(sm.c1[slot], it);
          pop();
        }
      }
    }
  }
  __syncthreads();

  // ------------------------------------------------------------------ phase C
  for (int cidx = tid; cidx < tile_n; cidx += NT) {
    const long long idx = tile0 + cidx;
    if (idx >= a.ncell) continue;
    const unsigned t = (unsigned)(tile_ii0 + cidx), dj = t / nxr_u;
    const int jj = tile_jj0 + (int)dj, ii = (int)(t - dj * nxr_u);
    const int i = ii - a.ring, j = jj - a.ring;
    const int off = j * a.usj + i;          // uniform layout (see phase A)
    auto LD = [&](const DArr& d, int o) -> FT { return __ldg(reinterpret_cast<const FT*>(d.p) + o); };
    auto LDP = [&](const DArr& d) -> FT { return reinterpret_cast<const FT*>(d.p)[off]; };   // plain load: written by this thread in phase A
    auto ST = [&](const DArr& d, FT v) { if (d.p) reinterpret_cast<FT*>(d.p)[off] = v; };
    const FT Tunits = LD(a.oT, off);
    const bool act = !a.mask.p || __ldg(reinterpret_cast<const uint8_t*>(a.mask.p) + off) != 0;
    FT Qv = FT(0), Qc = FT(0), Fv = FT(0), rtx = FT(0), rty = FT(0);
    const FT us = sm.U2[cidx], ts = sm.dth[cidx], qs = sm.dq[cidx];
    if (act) {
      // the exchange state was written by this very thread in phase A (or is an input): plain loads
      const FT Ta = LDP(a.xT);
#if COFLUX_TILE_CARRY2
      const FT du = sm.du[cidx], dv = sm.dv[cidx], rho = sm.rho[cidx], cp = sm.cp[cidx];
#else
      const FT ua = LDP(a.xu);
      const FT va = LDP(a.xv);
      FT du, dv;
      if (F.velocity == COFLUX_VELOCITY_RELATIVE) {
        du = ua - (LD(a.ou, off) + LD(a.ou, off + 1)) * FT(0.5);
        dv = va - (LD(a.ov, off) + LD(a.ov, off + a.usj)) * FT(0.5);
      } else { du = ua; dv = va; }
      const FT rho = LDP(a.rtx);
      const FT cp = LDP(a.rty);
#endif
      FT taux, tauy;
      if constexpr (LEAN) {
        const FT d2 = du * du + dv * dv;
        const FT k = (d2 > FT(1e-30)) ? (-us * us) * fm::rcp(fm::sqrt(d2)) : FT(0);   // −u★²/‖Δu‖ (0 when calm)
        taux = (d2 > FT(1e-30)) ? k * du : ((d2 == FT(0)) ? FT(0) : -us * us * du / M<FT>::sqrt(d2));
        tauy = (d2 > FT(1e-30)) ? k * dv : ((d2 == FT(0)) ? FT(0) : -us * us * dv / M<FT>::sqrt(d2));
      } else {
        const FT dU = M<FT>::sqrt(du * du + dv * dv);
        taux = (dU == FT(0)) ? dU : -us * us * du / dU;
        tauy = (dU == FT(0)) ? dU : -us * us * dv / dU;
      }
      const FT LH = c.LH_v0 + (c.cp_v - c.cp_l) * (Ta - c.T_0);
      Qv = -rho * us * qs * LH;
      Qc = -rho * cp * us * ts;
      Fv = -rho * us * qs;
      rtx = rho * taux; rty = rho * tauy;
    }
    ST(a.Qv, Qv); ST(a.Qc, Qc); ST(a.Fv, Fv);
    ST(a.rtx, rtx); ST(a.rty, rty); ST(a.Tsout, Tunits);
    ST(a.ust, us); ST(a.tst, ts); ST(a.qst, qs);
    if (a.iters.p) reinterpret_cast<int32_t*>(a.iters.p)[off] = act ? SI::get(sm.c1[cidx]) : 0;
    if (a.seam_east && i == a.Nx - 1 && j >= 0 && j < a.Ny) reinterpret_cast<FT*>(a.seam_east)[j] = rtx;
    if (ASSEMBLE) {
      if (i >= 0 && i < a.Nx && j >= 0 && j < a.Ny) {
        const FT Qs = LDP(a.xQs);
        const FT Ql = LDP(a.xQl);
        const FT Mp = LDP(a.xMp);
        const FT So = LD(a.oS, off);
        const FT conc = a.conc.p ? LD(a.conc, off) : FT(0);
        const FT Qio = a.Qio.p ? LD(a.Qio, off) : FT(0);
        const FT sio = a.salt_io.p ? LD(a.salt_io, off) : FT(0);
        FT JT, JS, Qu, Qal, Qts, J0, parts[3];
        assemble_tracers<FT>(P, act, conc, So, Tunits + P.T_offset, Qs, Ql, Mp, Qc, Qv, Fv, Qio, sio, JT, JS, Qu, Qal, Qts, J0, parts);
        ST(a.JT, JT); ST(a.JS, JS); ST(a.Qu, Qu); ST(a.Qal, Qal);
        ST(a.Qts, Qts); ST(a.J0, J0);
        if (a.avg.on) avg_epilogue_u<FT>(a.avg, off, JT, JS, Qc, Qv, parts);
      }
    }
  }
}

}  // namespace coflux

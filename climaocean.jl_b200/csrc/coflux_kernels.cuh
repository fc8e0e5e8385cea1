// coflux_kernels.cuh — the CUDA kernels of the surface-flux path (sm_100a).
//
//   flux_kernel<FT, SURF, INTERP, SOLVE, ASSEMBLE>
//       (·,0,1,0,0)  interpolate_atmosphere_state!           (row a3)
//       (·,0,0,1,0)  compute_atmosphere_ocean_fluxes!        (rows a4–a6)
//       (·,1,0,1,0)  compute_atmosphere_sea_ice_fluxes!      (row a7)
//       (·,0,1,1,1)  fused update_state!: a3 + a4/a5 + tracer/radiative part of a9
//   stress_kernel<FT>        centre→face momentum part of a9 (needs the neighbours' ρτ)
//   assemble_kernel<FT>      stand-alone compute_net_ocean_fluxes! (row a9)
//   ice_ocean_kernel<FT>     compute_sea_ice_ocean_fluxes! (row a8): frazil column sweep etc.
//
// Thread mapping: one thread per surface cell of the ring-extended surface, linearised with i
// fastest, so a warp reads/writes 32 consecutive elements of every halo-padded column-major
// parent (full 32 B sectors; the arbitrary halo offset of Julia-owned parents rules out 16 B
// vector or TMA box loads: (Nx+2H)·sizeof(FT) is in general not a multiple of 16).  All input
// loads of a cell are issued before the solve (ld.global.nc), the iteration runs in registers,
// all outputs are written once at the end.
#pragma once
#include "coflux_device.cuh"

namespace coflux {

struct DArr {      // 2-D view: element (i,j) at p + (i*si + j*sj) elements; p == nullptr → absent
  char* p;
  int64_t si, sj;
};
struct DSeries {   // two bracketing time levels of a series
  char* p1;
  char* p2;
  int64_t si, sj;
};
struct DCol {      // 3-D view for column sweeps
  char* p;
  int64_t si, sj, sk;
};

template <typename FT> __device__ __forceinline__ FT ldg(const DArr& a, int i, int j) {
  return __ldg(reinterpret_cast<const FT*>(a.p) + ((int64_t)i * a.si + (int64_t)j * a.sj));
}
// streaming load: read-only, not allocated in L1 (keeps L1 for the ψ table and the gathered atmosphere tiles)
#ifndef COFLUX_STREAM_LOADS
#define COFLUX_STREAM_LOADS 0
#endif
template <typename FT> __device__ __forceinline__ FT ldgs(const DArr& a, int i, int j);
template <> __device__ __forceinline__ double ldgs<double>(const DArr& a, int i, int j) {
  const double* p = reinterpret_cast<const double*>(a.p) + ((int64_t)i * a.si + (int64_t)j * a.sj);
#if COFLUX_STREAM_LOADS
  double v; asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p)); return v;
#else
  return __ldg(p);
#endif
}
template <> __device__ __forceinline__ float ldgs<float>(const DArr& a, int i, int j) {
  const float* p = reinterpret_cast<const float*>(a.p) + ((int64_t)i * a.si + (int64_t)j * a.sj);
#if COFLUX_STREAM_LOADS
  float v; asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p)); return v;
#else
  return __ldg(p);
#endif
}
template <typename FT> __device__ __forceinline__ void stg(const DArr& a, int i, int j, FT v) {
  if (a.p) reinterpret_cast<FT*>(a.p)[(int64_t)i * a.si + (int64_t)j * a.sj] = v;
}
__device__ __forceinline__ bool is_active(const DArr& m, int i, int j) {
  if (!m.p) return true;
  return __ldg(reinterpret_cast<const uint8_t*>(m.p) + ((int64_t)i * m.si + (int64_t)j * m.sj)) != 0;
}

// Running time averages (Oceananigans WindowedTimeAverage, the schedule of omip_diagnostics.jl:152-158):
//   result ← (result·T + field·Δt) / (T + Δt),  T = time already accumulated in the window
template <typename FT> struct AvgArgs {
  int on;
  DArr JT, JS, Qc, Qv, JTao, JTio, JSio;
  FT T, dt;
};
template <typename FT> __device__ __forceinline__ void avg_update(const DArr& d, int i, int j, FT x, FT T, FT dt) {
  if (!d.p) return;
  FT* p = reinterpret_cast<FT*>(d.p) + ((int64_t)i * d.si + (int64_t)j * d.sj);
  *p = (*p * T + x * dt) / (T + dt);
}

template <typename FT> struct FluxArgs {
  int nxr, nyr, ring, Nx, Ny;
  long long cell0, ncell;   // linear cell range [cell0, ncell) of the ring-extended surface handled by this launch
  // balanced tiling of the tile kernel: every CTA takes tile_cells ≤ TILE cells, chosen by the host so that the grid is a
  // whole number of waves of resident CTAs (0: TILE cells per CTA)
  int tile_cells, pad_;
  // uniform layout (tile kernel; set by the host, which makes it a condition of tile eligibility): every 2-D surface array
  // the kernel touches has stride_i == 1 and the SAME row pitch usj (elements) — true for Oceananigans parents on one grid —
  // so a cell's element offset j·usj + i is computed once and shared by ≈ 70 loads and stores; ssj: the same for the
  // atmosphere series (one gather-offset set for all nine).  32-bit: a plane has < 2³¹ elements (checked).
  int usj, ssj, fsj;           // fsj: row pitch of fi, fj, cos θ, sin θ (ring-extended parents)
  // a3 inputs
  DSeries su, sv, sT, sq, sp, sQs, sQl, srain, ssnow;
  DArr fi, fj, cs, sn;
  FT nfrac;
  // exchange state (written when INTERP, read otherwise)
  DArr xu, xv, xT, xp, xq, xQs, xQl, xMp;
  // surface state: ocean (u,v,T,S at k = Nz-1) or ice (u, v, T_top)
  DArr ou, ov, oT, oS, mask;
  DArr ih, iS, ialb, iconc;   // sea ice (SURF == 1)
  // interface outputs
  DArr Qv, Qc, Fv, rtx, rty, Tsout, ust, tst, qst, iters, Ttop_out;
  // assembly (tracer / radiative part)
  DArr conc, Qio, salt_io;
  DArr JT, JS, Qu, Qal, Qts, J0;
  // seam push (multi-GPU, ring == 0): last column of ρτx stored to the east neighbour
  char* seam_east;     // peer pointer, Ny elements (or nullptr)
  // land freshwater (JRA55PrescribedLand: rivers + icebergs, atmosphere.jl:46) on its own source grid / time axis
  DSeries sriv, sicb;
  DArr lfi, lfj;
  FT lnfrac;
  DArr ihs;            // snow thickness on sea ice (CCSM3 albedo), optional
  AvgArgs<FT> avg;     // running time averages of the flux fields (omip_diagnostics.jl:125-158), optional
  DevParams<FT> P;
};

template <typename FT>
__device__ __forceinline__ FT interp_series(const DSeries& s, int i0, int j0, int i1, int j1, FT w00, FT w01, FT w10,
                                            FT w11, FT nfrac) {
  const int64_t o00 = (int64_t)i0 * s.si + (int64_t)j0 * s.sj, o01 = (int64_t)i0 * s.si + (int64_t)j1 * s.sj;
  const int64_t o10 = (int64_t)i1 * s.si + (int64_t)j0 * s.sj, o11 = (int64_t)i1 * s.si + (int64_t)j1 * s.sj;
  const FT* a1 = reinterpret_cast<const FT*>(s.p1);
  const FT* a2 = reinterpret_cast<const FT*>(s.p2);
  FT v100 = __ldg(a1 + o00), v101 = __ldg(a1 + o01), v110 = __ldg(a1 + o10), v111 = __ldg(a1 + o11);
  FT v200 = __ldg(a2 + o00), v201 = __ldg(a2 + o01), v210 = __ldg(a2 + o10), v211 = __ldg(a2 + o11);
  FT p1 = w00 * v100 + w01 * v101 + w10 * v110 + w11 * v111;
  FT p2 = w00 * v200 + w01 * v201 + w10 * v210 + w11 * v211;
  return p2 * nfrac + p1 * (FT(1) - nfrac);
}

// the same with the four gather offsets given (uniform series layout: computed once for all series of a cell)
template <typename FT>
__device__ __forceinline__ FT interp_series_u(const DSeries& s, int o00, int o01, int o10, int o11, FT w00, FT w01, FT w10, FT w11, FT nfrac) {
  const FT* a1 = reinterpret_cast<const FT*>(s.p1);
  const FT* a2 = reinterpret_cast<const FT*>(s.p2);
  FT v100 = __ldg(a1 + o00), v101 = __ldg(a1 + o01), v110 = __ldg(a1 + o10), v111 = __ldg(a1 + o11);
  FT v200 = __ldg(a2 + o00), v201 = __ldg(a2 + o01), v210 = __ldg(a2 + o10), v211 = __ldg(a2 + o11);
  FT p1 = w00 * v100 + w01 * v101 + w10 * v110 + w11 * v111;
  FT p2 = w00 * v200 + w01 * v201 + w10 * v210 + w11 * v211;
  return p2 * nfrac + p1 * (FT(1) - nfrac);
}

// bilinear stencil of one fractional index pair (A8)
template <typename FT> struct Bilin { int i0, j0, i1, j1; FT w00, w01, w10, w11; };
template <typename FT> __device__ __forceinline__ Bilin<FT> bilin(FT fi, FT fj) {
  Bilin<FT> b;
  b.i0 = (int)M<FT>::trunc(fi); b.j0 = (int)M<FT>::trunc(fj);
  b.i1 = b.i0 + ((fi > FT(0)) - (fi < FT(0))); b.j1 = b.j0 + ((fj > FT(0)) - (fj < FT(0)));
  const FT xi = fi - M<FT>::floor(fi), eta = fj - M<FT>::floor(fj);
  b.w00 = (FT(1) - xi) * (FT(1) - eta); b.w01 = (FT(1) - xi) * eta; b.w10 = xi * (FT(1) - eta); b.w11 = xi * eta;
  return b;
}
// rivers + icebergs interpolated to ocean cell (i, j): the land part of the exchange freshwater flux (0 without land)
template <typename FT> __device__ __forceinline__ FT land_freshwater(const FluxArgs<FT>& a, int i, int j) {
  if (!a.lfi.p) return FT(0);
  const Bilin<FT> b = bilin<FT>(ldgs<FT>(a.lfi, i, j), ldgs<FT>(a.lfj, i, j));
  FT Mr = FT(0), Mi = FT(0);
  if (a.sriv.p1) Mr = interp_series<FT>(a.sriv, b.i0, b.j0, b.i1, b.j1, b.w00, b.w01, b.w10, b.w11, a.lnfrac);
  if (a.sicb.p1) Mi = interp_series<FT>(a.sicb, b.i0, b.j0, b.i1, b.j1, b.w00, b.w01, b.w10, b.w11, a.lnfrac);
  return Mr + Mi;
}
// sea-ice albedo of cell (i, j): prescribed plane / constant, or CCSM3 from the live h_i, h_s, T_s (atmosphere.jl:31-44)
template <typename FT> __device__ __forceinline__ FT sea_ice_albedo(const DevParams<FT>& P, const DArr& ialb, const DArr& ih, const DArr& ihs,
                                                                    int i, int j, FT TsK) {
  if (P.ice_albedo_kind == COFLUX_SEA_ICE_ALBEDO_CCSM3)
    return ccsm3_albedo<FT>(P.ccsm3, ldg<FT>(ih, i, j), ihs.p ? ldg<FT>(ihs, i, j) : FT(0), TsK);
  return ialb.p ? ldg<FT>(ialb, i, j) : P.alb_i;
}

// tracer / radiative part of the net ocean flux assembly for one interior cell (A9)
template <typename FT>
__device__ __forceinline__ void assemble_tracers(const DevParams<FT>& P, bool act, FT conc, FT So, FT TsK, FT Qs, FT Ql,
                                                 FT Mp, FT Qc, FT Qv, FT Mv, FT Qio, FT salt_io, FT& JT, FT& JS, FT& Qu,
                                                 FT& Qal, FT& Qts, FT& J0, FT* parts = nullptr /* JTao, JTio, JSio */) {
  const FT rho0inv = P.rho0inv, rhofinv = P.rhofinv;
  Qu = P.emis_o * P.sigma * TsK * TsK * TsK * TsK;
  Qal = -P.emis_o * Ql;
  Qts = -(FT(1) - P.alb_o) * Qs;
  const FT Qss = P.sw_pen ? FT(0) : Qts;
  const FT SQ = Qu + Qc + Qv + Qal + Qss;
  FT SF = -Mp * rhofinv;
  SF += Mv * rhofinv;
  const FT JTao = LMath<FT>::div(SQ * rho0inv, P.c0);          // LMath::div: the lean division in Float64 (bit-identical to IEEE)
  FT JSao = -So * SF;
  if (So < P.Smin && JSao > FT(0)) JSao = FT(0);
  const FT JTao_w = (FT(1) - conc) * JTao, JTio = LMath<FT>::div(Qio * rho0inv, P.c0), JSio = salt_io * conc;
  JT = JTao_w + JTio;
  JS = (FT(1) - conc) * JSao + JSio;
  J0 = LMath<FT>::div((FT(1) - conc) * Qts * rho0inv, P.c0);
  if (!act) { JT = JS = J0 = FT(0); Qu = Qal = Qts = FT(0); }
  if (parts) { parts[0] = act ? JTao_w : FT(0); parts[1] = act ? JTio : FT(0); parts[2] = act ? JSio : FT(0); }
}

// epilogue of the fused kernels: running averages of the tracer / turbulent fluxes of one interior cell
template <typename FT>
__device__ __forceinline__ void avg_epilogue(const AvgArgs<FT>& g, int i, int j, FT JT, FT JS, FT Qc, FT Qv, const FT* parts) {
  // all the old means are loaded before the first store: seven read-modify-writes through possibly aliasing pointers would
  // otherwise run one DRAM latency after the other (measured: +0.36 ms on the 3.2 ms flux kernel at 1/12°)
  const DArr* d[7] = {&g.JT, &g.JS, &g.Qc, &g.Qv, &g.JTao, &g.JTio, &g.JSio};
  const FT x[7] = {JT, JS, Qc, Qv, parts[0], parts[1], parts[2]};
  FT* p[7];
  FT old[7];
#pragma unroll
  for (int k = 0; k < 7; ++k) {
    p[k] = d[k]->p ? reinterpret_cast<FT*>(d[k]->p) + ((int64_t)i * d[k]->si + (int64_t)j * d[k]->sj) : nullptr;
    old[k] = p[k] ? *p[k] : FT(0);
  }
#pragma unroll
  for (int k = 0; k < 7; ++k)
    if (p[k]) *p[k] = (old[k] * g.T + x[k] * g.dt) / (g.T + g.dt);
}

// the same with the cell's element offset given (uniform layout of the tile kernel)
template <typename FT>
__device__ __forceinline__ void avg_epilogue_u(const AvgArgs<FT>& g, int off, FT JT, FT JS, FT Qc, FT Qv, const FT* parts) {
  const DArr* d[7] = {&g.JT, &g.JS, &g.Qc, &g.Qv, &g.JTao, &g.JTio, &g.JSio};
  const FT x[7] = {JT, JS, Qc, Qv, parts[0], parts[1], parts[2]};
  FT old[7];
#pragma unroll
  for (int k = 0; k < 7; ++k) old[k] = d[k]->p ? reinterpret_cast<const FT*>(d[k]->p)[off] : FT(0);
#pragma unroll
  for (int k = 0; k < 7; ++k)
    if (d[k]->p) reinterpret_cast<FT*>(d[k]->p)[off] = (old[k] * g.T + x[k] * g.dt) / (g.T + g.dt);
}

// UNI: uniform parent layout (FluxArgs::usj, ssj, fsj — see flux_tile_kernel): one element offset per cell for all surface
// arrays, one gather-offset set for all series, 32-bit index arithmetic.  Used for the fused ocean path when the host has
// verified the layout (`:ncar`: 1.88 → see profiles/README.md); UNI = false addresses every array through its own strides.
template <typename FT, int SURF, bool INTERP, bool SOLVE, bool ASSEMBLE, bool UNI = false>
__global__ void __launch_bounds__(128) flux_kernel(const __grid_constant__ FluxArgs<FT> a) {
  static_assert(!UNI || SURF == 0, "the uniform-layout form covers the ocean surface arrays only");
  const long long idx = a.cell0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= a.ncell) return;
  int ii, jj;
  if constexpr (UNI) { const unsigned u = (unsigned)idx, q = u / (unsigned)a.nxr; jj = (int)q; ii = (int)(u - q * (unsigned)a.nxr); }
  else { jj = (int)(idx / a.nxr); ii = (int)(idx - (long long)jj * a.nxr); }
  const int i = ii - a.ring, j = jj - a.ring;
  const int off = UNI ? j * a.usj + i : 0;
  auto L = [&](const DArr& d, int di = 0, int dj = 0) -> FT {
    if constexpr (UNI) return __ldg(reinterpret_cast<const FT*>(d.p) + (off + di + dj * a.usj));
    else return ldg<FT>(d, i + di, j + dj);
  };
  auto S = [&](const DArr& d, FT v) {
    if constexpr (UNI) { if (d.p) reinterpret_cast<FT*>(d.p)[off] = v; }
    else stg<FT>(d, i, j, v);
  };
  auto wet = [&]() -> bool {
    if constexpr (UNI) return !a.mask.p || __ldg(reinterpret_cast<const uint8_t*>(a.mask.p) + off) != 0;
    else return is_active(a.mask, i, j);
  };

  FT ua, va, Ta, pa, qa, Qs, Ql, Mp;
  if (INTERP) {
    const int foff = UNI ? j * a.fsj + i : 0;
    FT fi, fj;
    if constexpr (UNI) { fi = __ldg(reinterpret_cast<const FT*>(a.fi.p) + foff); fj = __ldg(reinterpret_cast<const FT*>(a.fj.p) + foff); }
    else { fi = ldg<FT>(a.fi, i, j); fj = ldg<FT>(a.fj, i, j); }
    const int i0 = (int)M<FT>::trunc(fi), j0 = (int)M<FT>::trunc(fj);
    const int i1 = i0 + ((fi > FT(0)) - (fi < FT(0))), j1 = j0 + ((fj > FT(0)) - (fj < FT(0)));
    const FT xi = fi - M<FT>::floor(fi), eta = fj - M<FT>::floor(fj);
    const FT w00 = (FT(1) - xi) * (FT(1) - eta), w01 = (FT(1) - xi) * eta, w10 = xi * (FT(1) - eta), w11 = xi * eta;
    const int o00 = j0 * a.ssj + i0, o01 = j1 * a.ssj + i0, o10 = j0 * a.ssj + i1, o11 = j1 * a.ssj + i1;
    auto SER = [&](const DSeries& d) -> FT {
      if constexpr (UNI) return interp_series_u<FT>(d, o00, o01, o10, o11, w00, w01, w10, w11, a.nfrac);
      else return interp_series<FT>(d, i0, j0, i1, j1, w00, w01, w10, w11, a.nfrac);
    };
    ua = SER(a.su);
    va = SER(a.sv);
    Ta = SER(a.sT);
    qa = SER(a.sq);
    pa = SER(a.sp);
    Qs = SER(a.sQs);
    Ql = SER(a.sQl);
    Mp = FT(0);
    if (a.srain.p1) Mp += SER(a.srain);
    if (a.ssnow.p1) Mp += SER(a.ssnow);
    if (a.lfi.p) Mp += land_freshwater<FT>(a, i, j);
    if (a.cs.p && a.sn.p) {
      FT cs, sn;
      if constexpr (UNI) { cs = __ldg(reinterpret_cast<const FT*>(a.cs.p) + foff); sn = __ldg(reinterpret_cast<const FT*>(a.sn.p) + foff); }
      else { cs = ldg<FT>(a.cs, i, j); sn = ldg<FT>(a.sn, i, j); }
      const FT ur = ua * cs + va * sn, vr = -ua * sn + va * cs;
      ua = ur; va = vr;
    }
    S(a.xu, ua); S(a.xv, va); S(a.xT, Ta); S(a.xp, pa);
    S(a.xq, qa); S(a.xQs, Qs); S(a.xQl, Ql); S(a.xMp, Mp);
  } else {
    ua = L(a.xu); va = L(a.xv); Ta = L(a.xT); pa = L(a.xp);
    qa = L(a.xq); Qs = L(a.xQs); Ql = L(a.xQl);
    Mp = (ASSEMBLE && a.xMp.p) ? L(a.xMp) : FT(0);
  }
  if (!SOLVE) return;

  const DevParams<FT>& P = a.P;
  CellIn<FT> in;
  in.ua = ua; in.va = va; in.Ta = Ta; in.pa = pa; in.qa = qa; in.Qs = Qs; in.Ql = Ql;
  in.us = (L(a.ou) + L(a.ou, 1, 0)) * FT(0.5);
  in.vs = (L(a.ov) + L(a.ov, 0, 1)) * FT(0.5);
  const FT Tunits = L(a.oT);
  in.Ts0 = Tunits + P.T_offset;
  in.So = (SURF == 0) ? L(a.oS) : FT(0);
  bool act = wet();
  if (SURF == 1) {
    in.h_ice = ldg<FT>(a.ih, i, j);
    in.S_ice = ldg<FT>(a.iS, i, j);
    in.albedo = sea_ice_albedo<FT>(P, a.ialb, a.ih, a.ihs, i, j, in.Ts0);
    const FT conc = ldg<FT>(a.iconc, i, j);
    act = act && (conc > FT(0)) && (in.h_ice > FT(0));
  } else {
    in.h_ice = in.S_ice = in.albedo = FT(0);
  }

  FT Qv = FT(0), Qc = FT(0), Fv = FT(0), rtx = FT(0), rty = FT(0), Tsout = Tunits, us = FT(0), ts = FT(0), qs = FT(0);
  int its = 0;
  if (act) {
    CellOut<FT> o;
    if (SURF == 0) solve_cell<FT, 0>(P, P.ao, in, o);
    else solve_cell<FT, 1>(P, P.ai, in, o);
    const FT dU = LMath<FT>::sqrt(o.du * o.du + o.dv * o.dv);      // LMath: the lean square root / division in Float64 (bit-identical to IEEE)
    const FT taux = (dU == FT(0)) ? dU : LMath<FT>::div(-o.ustar * o.ustar * o.du, dU);
    const FT tauy = (dU == FT(0)) ? dU : LMath<FT>::div(-o.ustar * o.ustar * o.dv, dU);
    const ThermoC<FT>& c = P.th;
    const FT LH = (SURF == 0) ? c.LH_v0 + (c.cp_v - c.cp_l) * (Ta - c.T_0) : c.LH_s0 + (c.cp_v - c.cp_i) * (Ta - c.T_0);
    Qv = -o.rho_a * o.ustar * o.qstar * LH;
    Qc = -o.rho_a * o.cp_a * o.ustar * o.tstar;
    Fv = -o.rho_a * o.ustar * o.qstar;
    rtx = o.rho_a * taux; rty = o.rho_a * tauy;
    Tsout = o.Ts - P.T_offset;
    us = o.ustar; ts = o.tstar; qs = o.qstar; its = o.it;
  }
  S(a.Qv, Qv); S(a.Qc, Qc); S(a.Fv, Fv);
  S(a.rtx, rtx); S(a.rty, rty); S(a.Tsout, Tsout);
  S(a.ust, us); S(a.tst, ts); S(a.qst, qs);
  if (SURF == 1) S(a.Ttop_out, Tsout);
  if (a.iters.p) {
    if constexpr (UNI) reinterpret_cast<int32_t*>(a.iters.p)[off] = its;
    else reinterpret_cast<int32_t*>(a.iters.p)[(int64_t)i * a.iters.si + (int64_t)j * a.iters.sj] = its;
  }
  if (a.seam_east && i == a.Nx - 1 && j >= 0 && j < a.Ny) reinterpret_cast<FT*>(a.seam_east)[j] = rtx;

  if (ASSEMBLE) {
    if (i >= 0 && i < a.Nx && j >= 0 && j < a.Ny) {
      const FT conc = a.conc.p ? L(a.conc) : FT(0);
      const FT Qio = a.Qio.p ? L(a.Qio) : FT(0);
      const FT sio = a.salt_io.p ? L(a.salt_io) : FT(0);
      FT JT, JS, Qu, Qal, Qts, J0, parts[3];
      assemble_tracers<FT>(P, wet(), conc, in.So, Tsout + P.T_offset, Qs, Ql, Mp, Qc, Qv, Fv, Qio, sio,
                           JT, JS, Qu, Qal, Qts, J0, parts);
      S(a.JT, JT); S(a.JS, JS); S(a.Qu, Qu); S(a.Qal, Qal);
      S(a.Qts, Qts); S(a.J0, J0);
      if (a.avg.on) {
        if constexpr (UNI) avg_epilogue_u<FT>(a.avg, off, JT, JS, Qc, Qv, parts);
        else avg_epilogue<FT>(a.avg, i, j, JT, JS, Qc, Qv, parts);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Lane-refill form of the stand-alone interface solve (rows a4–a7 without interpolation / assembly), for the
// solves whose pass count varies strongly from cell to cell — above all the atmosphere–sea-ice solve with a skin
// temperature: 29 passes on average, but one cell in nine runs to maxiter (limit cycle of the clamped T_s update), so
// with one cell per thread virtually every warp waits for a 100-pass lane.  Here every lane owns a CellSolver; a lane
// whose cell has converged stores its result and pops the next cell of the CTA's tile from a shared counter.  The
// loads / stores of a popped cell are not coalesced (≈ 25 words per cell against ≈ 29 passes × 2 000 instructions).
// Arithmetic per cell is exactly that of flux_kernel → bit-identical results.
// ---------------------------------------------------------------------------------------------
#ifndef COFLUX_REFILL_BATCH
#define COFLUX_REFILL_BATCH 8
#endif
template <typename FT, int SURF, int TILE>
__global__ void __launch_bounds__(128) flux_refill_kernel(const __grid_constant__ FluxArgs<FT> a) {
  __shared__ int head;
  if (threadIdx.x == 0) head = 0;
  __syncthreads();
  const DevParams<FT>& P = a.P;
  const FluxP<FT>& F = (SURF == 0) ? P.ao : P.ai;
  const long long tile0 = a.cell0 + (long long)blockIdx.x * TILE;
  const int ntile = (int)((a.ncell - tile0 < TILE) ? (a.ncell - tile0) : TILE);
  CellSolver<FT, SURF> s;
  int cur = -1, ci = 0, cj = 0;
  FT Tunits = FT(0);

  auto store = [&](bool act) {
    FT Qv = FT(0), Qc = FT(0), Fv = FT(0), rtx = FT(0), rty = FT(0), Tsout = Tunits, us = FT(0), ts = FT(0), qs = FT(0);
    int its = 0;
    if (act) {
      CellOut<FT> o;
      s.finish(o);
      const FT dU = M<FT>::sqrt(o.du * o.du + o.dv * o.dv);
      const FT taux = (dU == FT(0)) ? dU : -o.ustar * o.ustar * o.du / dU;
      const FT tauy = (dU == FT(0)) ? dU : -o.ustar * o.ustar * o.dv / dU;
      const ThermoC<FT>& c = P.th;
      const FT Ta = s.in.Ta;
      const FT LH = (SURF == 0) ? c.LH_v0 + (c.cp_v - c.cp_l) * (Ta - c.T_0) : c.LH_s0 + (c.cp_v - c.cp_i) * (Ta - c.T_0);
      Qv = -o.rho_a * o.ustar * o.qstar * LH;
      Qc = -o.rho_a * o.cp_a * o.ustar * o.tstar;
      Fv = -o.rho_a * o.ustar * o.qstar;
      rtx = o.rho_a * taux; rty = o.rho_a * tauy;
      Tsout = o.Ts - P.T_offset;
      us = o.ustar; ts = o.tstar; qs = o.qstar; its = o.it;
    }
    const int i = ci, j = cj;
    stg<FT>(a.Qv, i, j, Qv); stg<FT>(a.Qc, i, j, Qc); stg<FT>(a.Fv, i, j, Fv);
    stg<FT>(a.rtx, i, j, rtx); stg<FT>(a.rty, i, j, rty); stg<FT>(a.Tsout, i, j, Tsout);
    stg<FT>(a.ust, i, j, us); stg<FT>(a.tst, i, j, ts); stg<FT>(a.qst, i, j, qs);
    if (SURF == 1) stg<FT>(a.Ttop_out, i, j, Tsout);
    if (a.iters.p) reinterpret_cast<int32_t*>(a.iters.p)[(int64_t)i * a.iters.si + (int64_t)j * a.iters.sj] = its;
    if (a.seam_east && i == a.Nx - 1 && j >= 0 && j < a.Ny) reinterpret_cast<FT*>(a.seam_east)[j] = rtx;
  };
  // pop cells until one needs passes (inactive cells and cells that stop at pass 0 are finished on the spot)
  auto pop = [&]() {
    for (;;) {
      cur = atomicAdd(&head, 1);
      if (cur >= ntile) { cur = -1; return; }
      const long long idx = tile0 + cur;
      const int jj = (int)(idx / a.nxr);
      const int ii = (int)(idx - (long long)jj * a.nxr);
      ci = ii - a.ring; cj = jj - a.ring;
      const int i = ci, j = cj;
      CellIn<FT> in;
      in.ua = ldg<FT>(a.xu, i, j); in.va = ldg<FT>(a.xv, i, j); in.Ta = ldg<FT>(a.xT, i, j); in.pa = ldg<FT>(a.xp, i, j);
      in.qa = ldg<FT>(a.xq, i, j); in.Qs = ldg<FT>(a.xQs, i, j); in.Ql = ldg<FT>(a.xQl, i, j);
      in.us = (ldg<FT>(a.ou, i, j) + ldg<FT>(a.ou, i + 1, j)) * FT(0.5);
      in.vs = (ldg<FT>(a.ov, i, j) + ldg<FT>(a.ov, i, j + 1)) * FT(0.5);
      Tunits = ldg<FT>(a.oT, i, j);
      in.Ts0 = Tunits + P.T_offset;
      in.So = (SURF == 0) ? ldg<FT>(a.oS, i, j) : FT(0);
      bool act = is_active(a.mask, i, j);
      if (SURF == 1) {
        in.h_ice = ldg<FT>(a.ih, i, j);
        in.S_ice = ldg<FT>(a.iS, i, j);
        in.albedo = sea_ice_albedo<FT>(P, a.ialb, a.ih, a.ihs, i, j, in.Ts0);
        const FT conc = ldg<FT>(a.iconc, i, j);
        act = act && (conc > FT(0)) && (in.h_ice > FT(0));
      } else {
        in.h_ice = in.S_ice = in.albedo = FT(0);
      }
      if (act) {
        s.init(P, F, in);
        if (s.go) return;
      }
      store(act);
    }
  };
  // Finished lanes store and refill TOGETHER: popping a cell costs ≈ 600 instructions (loads, both thermodynamic states)
  // and a lane finishes every ≈ 29 passes, so refilling each lane on its own would run that code in almost every pass of
  // the warp with one lane active.  A finished lane therefore waits until COFLUX_REFILL_BATCH lanes are waiting (or nobody
  // is iterating any more); the wait idles ≈ 4 of 32 lanes, the batching divides the refill cost by the batch size.
  bool waiting = false;          // this lane's cell has converged, its result is still in registers
  pop();
  for (;;) {
    const bool iterating = (cur >= 0) && !waiting;
    if (iterating) {
      s.pass(P, F);
      if (!s.go) waiting = true;
    }
    const unsigned want = __ballot_sync(0xffffffffu, waiting);
    const unsigned busy = __ballot_sync(0xffffffffu, (cur >= 0) && !waiting);
    if (want == 0u && busy == 0u) break;
    if (__popc(want) >= COFLUX_REFILL_BATCH || busy == 0u) {
      if (waiting) { store(true); waiting = false; pop(); }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Tile form of the atmosphere–sea-ice solve (row a7) for the sea-ice parameter sets that take CellSolver::pass_ice
// (skin temperature, fixed roughness lengths, SHEBA / Large–Yeager ψ).  Same three phases as flux_tile_kernel:
//   A  per cell, all lanes busy: loads and CellSolver::init (both thermodynamic states — the expensive part of starting a
//      cell); the 16 numbers a pass needs go to shared memory, the cell is queued;
//   B  lanes pop a cell (17 shared-memory loads rebuild the solver), iterate pass() until it stops, write the state
//      back and pop the next one — so a warp is not held up by its slowest cell (29 passes on average, up to 100);
//   C  per cell: fluxes and coalesced stores.
// Every arithmetic operation is CellSolver's, i.e. results are bit-identical to flux_kernel<FT,1,…>.
// ---------------------------------------------------------------------------------------------
template <typename FT, int TILE> struct IceTileSmem {
  FT Ta[TILE], pa[TILE], Qs[TILE], Ql[TILE], S_ice[TILE], h_ice[TILE], albedo[TILE];   // CellIn fields a pass reads
  FT rho[TILE], cp[TILE], qv[TILE], theta_a[TILE], du2dv2[TILE];                        // hoisted by init()
  FT us[TILE], ts[TILE], qs[TILE], Ts[TILE];                                             // state / result
  int it[TILE];
  unsigned short queue[TILE];
  int n_queued, head;
};
#ifndef COFLUX_ICE_MIN_BLOCKS
#define COFLUX_ICE_MIN_BLOCKS 4
#endif
template <typename FT, int TILE>
__global__ void __launch_bounds__(128, COFLUX_ICE_MIN_BLOCKS) ice_tile_kernel(const __grid_constant__ FluxArgs<FT> a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  IceTileSmem<FT, TILE>& sm = *reinterpret_cast<IceTileSmem<FT, TILE>*>(smem_raw);
  const DevParams<FT>& P = a.P;
  const FluxP<FT>& F = P.ai;
  const int tid = threadIdx.x;
  const long long tile0 = a.cell0 + (long long)blockIdx.x * TILE;
  if (tid == 0) { sm.n_queued = 0; sm.head = 0; }
  __syncthreads();

  // ------------------------------------------------------------------ phase A
  for (int cidx = tid; cidx < TILE; cidx += 128) {
    const long long idx = tile0 + cidx;
    if (idx >= a.ncell) { sm.it[cidx] = -1; continue; }
    const int jj = (int)(idx / a.nxr);
    const int ii = (int)(idx - (long long)jj * a.nxr);
    const int i = ii - a.ring, j = jj - a.ring;
    CellIn<FT> in;
    in.ua = ldg<FT>(a.xu, i, j); in.va = ldg<FT>(a.xv, i, j); in.Ta = ldg<FT>(a.xT, i, j); in.pa = ldg<FT>(a.xp, i, j);
    in.qa = ldg<FT>(a.xq, i, j); in.Qs = ldg<FT>(a.xQs, i, j); in.Ql = ldg<FT>(a.xQl, i, j);
    in.us = (ldg<FT>(a.ou, i, j) + ldg<FT>(a.ou, i + 1, j)) * FT(0.5);
    in.vs = (ldg<FT>(a.ov, i, j) + ldg<FT>(a.ov, i, j + 1)) * FT(0.5);
    in.Ts0 = ldg<FT>(a.oT, i, j) + P.T_offset;
    in.So = FT(0);
    in.h_ice = ldg<FT>(a.ih, i, j);
    in.S_ice = ldg<FT>(a.iS, i, j);
    in.albedo = sea_ice_albedo<FT>(P, a.ialb, a.ih, a.ihs, i, j, in.Ts0);
    const FT conc = ldg<FT>(a.iconc, i, j);
    const bool act = is_active(a.mask, i, j) && (conc > FT(0)) && (in.h_ice > FT(0));
    FT us = FT(0), ts = FT(0), qs = FT(0), Ts = in.Ts0;
    int it = 0;
    if (act) {
      CellSolver<FT, 1> s;
      s.init(P, F, in);
      us = s.ustar; ts = s.tstar; qs = s.qstar; Ts = s.Ts;
      sm.Ta[cidx] = in.Ta; sm.pa[cidx] = in.pa; sm.Qs[cidx] = in.Qs; sm.Ql[cidx] = in.Ql; sm.S_ice[cidx] = in.S_ice;
      sm.h_ice[cidx] = in.h_ice; sm.albedo[cidx] = in.albedo;
      sm.rho[cidx] = s.atm.rho; sm.cp[cidx] = s.atm.cp_m; sm.qv[cidx] = s.atm.q_vap; sm.theta_a[cidx] = s.theta_a;
      sm.du2dv2[cidx] = s.du2dv2;
      if (s.go) sm.queue[atomicAdd(&sm.n_queued, 1)] = (unsigned short)cidx;
    } else {
      sm.rho[cidx] = FT(0); sm.cp[cidx] = FT(0); sm.Ta[cidx] = in.Ta;
      it = -2;                       // inactive: phase C writes zeros
    }
    sm.us[cidx] = us; sm.ts[cidx] = ts; sm.qs[cidx] = qs; sm.Ts[cidx] = Ts; sm.it[cidx] = it;
  }
  __syncthreads();
  const int n_total = sm.n_queued;

  // ------------------------------------------------------------------ phase B: lane refill
  {
    CellSolver<FT, 1> s;
    int slot = -1;
    // everything init() derives from the parameters alone
    {
      CellIn<FT> z{};
      z.Ta = FT(280); z.pa = FT(1e5); z.qa = FT(1e-3); z.Ts0 = FT(270); z.h_ice = FT(1);
      s.in = z;
      s.x = FT(1);
      s.delta = P.th.eps - FT(1);
      s.ly = false; s.U_ly = FT(0); s.rcdn_ly = FT(0); s.lnh10 = FT(0);
      s.fixed = (F.stop_kind == COFLUX_STOP_FIXED_ITERATIONS);
      s.ice_fast = true;
      s.lnh_lu = M<FT>::log(P.h / F.mr.fixed); s.lnh_lt = M<FT>::log(P.h / F.tr.fixed); s.lnh_lq = M<FT>::log(P.h / F.qr.fixed);
      s.du = s.dv = FT(0);
    }
    auto pop = [&]() {
      const int pos = atomicAdd(&sm.head, 1);
      slot = -1;
      if (pos < n_total) {
        slot = sm.queue[pos];
        s.in.Ta = sm.Ta[slot]; s.in.pa = sm.pa[slot]; s.in.Qs = sm.Qs[slot]; s.in.Ql = sm.Ql[slot]; s.in.S_ice = sm.S_ice[slot];
        s.in.h_ice = sm.h_ice[slot]; s.in.albedo = sm.albedo[slot];
        s.atm.rho = sm.rho[slot]; s.atm.cp_m = sm.cp[slot]; s.atm.q_vap = sm.qv[slot]; s.theta_a = sm.theta_a[slot];
        s.du2dv2 = sm.du2dv2[slot];
        s.ustar = sm.us[slot]; s.tstar = sm.ts[slot]; s.qstar = sm.qs[slot]; s.Ts = sm.Ts[slot];
        s.it = 0; s.go = true;
        s.su = s.ustar; s.st = s.tstar; s.sq = s.qstar; s.sT = s.Ts; s.sr = s.rcdn_ly;
        s.snap_it = 0; s.window = 1; s.stop_at = -1;
      }
    };
    pop();
    while (__any_sync(0xffffffffu, slot >= 0)) {
      if (slot >= 0) {
        s.pass(P, F);
        if (!s.go) {
          sm.us[slot] = s.ustar; sm.ts[slot] = s.tstar; sm.qs[slot] = s.qstar; sm.Ts[slot] = s.Ts; sm.it[slot] = s.it;
          pop();
        }
      }
    }
  }
  __syncthreads();

  // ------------------------------------------------------------------ phase C
  for (int cidx = tid; cidx < TILE; cidx += 128) {
    const long long idx = tile0 + cidx;
    if (idx >= a.ncell) continue;
    const int jj = (int)(idx / a.nxr);
    const int ii = (int)(idx - (long long)jj * a.nxr);
    const int i = ii - a.ring, j = jj - a.ring;
    const int it = sm.it[cidx];
    const bool act = it != -2;
    const FT Tunits = ldg<FT>(a.oT, i, j);
    FT Qv = FT(0), Qc = FT(0), Fv = FT(0), rtx = FT(0), rty = FT(0), Tsout = Tunits, us = FT(0), ts = FT(0), qs = FT(0);
    if (act) {
      us = sm.us[cidx]; ts = sm.ts[cidx]; qs = sm.qs[cidx];
      const FT ua = ldg<FT>(a.xu, i, j), va = ldg<FT>(a.xv, i, j);
      FT du, dv;
      if (F.velocity == COFLUX_VELOCITY_RELATIVE) {
        du = ua - (ldg<FT>(a.ou, i, j) + ldg<FT>(a.ou, i + 1, j)) * FT(0.5);
        dv = va - (ldg<FT>(a.ov, i, j) + ldg<FT>(a.ov, i, j + 1)) * FT(0.5);
      } else { du = ua; dv = va; }
      const FT rho = sm.rho[cidx], cp = sm.cp[cidx], Ta = sm.Ta[cidx];
      const FT dU = M<FT>::sqrt(du * du + dv * dv);
      const FT taux = (dU == FT(0)) ? dU : -us * us * du / dU;
      const FT tauy = (dU == FT(0)) ? dU : -us * us * dv / dU;
      const ThermoC<FT>& c = P.th;
      const FT LH = c.LH_s0 + (c.cp_v - c.cp_i) * (Ta - c.T_0);
      Qv = -rho * us * qs * LH;
      Qc = -rho * cp * us * ts;
      Fv = -rho * us * qs;
      rtx = rho * taux; rty = rho * tauy;
      Tsout = sm.Ts[cidx] - P.T_offset;
    }
    stg<FT>(a.Qv, i, j, Qv); stg<FT>(a.Qc, i, j, Qc); stg<FT>(a.Fv, i, j, Fv);
    stg<FT>(a.rtx, i, j, rtx); stg<FT>(a.rty, i, j, rty); stg<FT>(a.Tsout, i, j, Tsout);
    stg<FT>(a.ust, i, j, us); stg<FT>(a.tst, i, j, ts); stg<FT>(a.qst, i, j, qs);
    stg<FT>(a.Ttop_out, i, j, Tsout);
    if (a.iters.p) reinterpret_cast<int32_t*>(a.iters.p)[(int64_t)i * a.iters.si + (int64_t)j * a.iters.sj] = act ? it : 0;
    if (a.seam_east && i == a.Nx - 1 && j >= 0 && j < a.Ny) reinterpret_cast<FT*>(a.seam_east)[j] = rtx;
  }
}

// ---------------------------------------------------------------------------------------------
// Queue form of the same solve: the cells of a tile wait in GLOBAL memory, not in shared memory.
// With 2–100 passes per cell (mean 29 in Float64, one cell in eight on a limit cycle) and the two cells per lane that
// 134 B of shared memory per queued cell allow, ice_tile_kernel keeps only 15.5 of 32 lanes busy: every tile ends with a
// long tail of a few slow cells (ncu, round 2).  Here phase A parks the five numbers init() derives (ρ_a, c_p,m, q_v, θ_a,
// ‖Δu‖²) and the albedo in the cell's own elements of six OUTPUT arrays (Q_v, Q_c, F_v, ρτx, ρτy, T_s — overwritten with
// the results when the cell finishes); everything else a pass needs is an input and is simply read again.  Shared memory
// holds only the 2-byte queue entries, so a tile is thousands of cells (tens per lane) and the tail all but disappears.
// Popping a cell costs 14 scattered loads, finishing one 6 loads + 11 scattered stores — against ≈ 29 passes of ≈ 800
// instructions.  Phase C is gone: the lane that finishes a cell computes and stores its fluxes.  All arithmetic is
// CellSolver's, in the same order: results are bit-identical to ice_tile_kernel and flux_kernel<FT,1,…>.
// MEASURED (B200, round 2): lane occupancy does improve, the time does not — 19.9 ms against 16.9 ms for ice_tile_kernel at
// 1/12° with 92 % ice cover.  About one lane of every warp pops a cell per pass, and its scattered loads stall the whole
// warp for a memory round trip; the shared-memory pop of the tile form costs ≈ 30 cycles.  Opt-in: COFLUX_ICE_QUEUE=1.
// ---------------------------------------------------------------------------------------------
template <int TILE> struct IceQueueSmem {
  unsigned short queue[TILE];
  int n_queued, head;
};
template <typename FT> __device__ __forceinline__ FT ldc(const DArr& a, int i, int j) {      // coherent load (parked values)
  return reinterpret_cast<const FT*>(a.p)[(int64_t)i * a.si + (int64_t)j * a.sj];
}
// fluxes and stores of one finished (or inactive) cell — phase C of ice_tile_kernel, cell by cell
template <typename FT>
__device__ __forceinline__ void ice_finish_cell(const FluxArgs<FT>& a, int i, int j, bool act, FT us, FT ts, FT qs, FT Ts, FT rho, FT cp, FT Ta, int it) {
  const DevParams<FT>& P = a.P;
  const FluxP<FT>& F = P.ai;
  const FT Tunits = ldg<FT>(a.oT, i, j);
  FT Qv = FT(0), Qc = FT(0), Fv = FT(0), rtx = FT(0), rty = FT(0), Tsout = Tunits;
  if (act) {
    const FT ua = ldg<FT>(a.xu, i, j), va = ldg<FT>(a.xv, i, j);
    FT du, dv;
    if (F.velocity == COFLUX_VELOCITY_RELATIVE) {
      du = ua - (ldg<FT>(a.ou, i, j) + ldg<FT>(a.ou, i + 1, j)) * FT(0.5);
      dv = va - (ldg<FT>(a.ov, i, j) + ldg<FT>(a.ov, i, j + 1)) * FT(0.5);
    } else { du = ua; dv = va; }
    const FT dU = M<FT>::sqrt(du * du + dv * dv);
    const FT taux = (dU == FT(0)) ? dU : -us * us * du / dU;
    const FT tauy = (dU == FT(0)) ? dU : -us * us * dv / dU;
    const ThermoC<FT>& c = P.th;
    const FT LH = c.LH_s0 + (c.cp_v - c.cp_i) * (Ta - c.T_0);
    Qv = -rho * us * qs * LH;
    Qc = -rho * cp * us * ts;
    Fv = -rho * us * qs;
    rtx = rho * taux; rty = rho * tauy;
    Tsout = Ts - P.T_offset;
  } else { us = ts = qs = FT(0); }
  stg<FT>(a.Qv, i, j, Qv); stg<FT>(a.Qc, i, j, Qc); stg<FT>(a.Fv, i, j, Fv);
  stg<FT>(a.rtx, i, j, rtx); stg<FT>(a.rty, i, j, rty); stg<FT>(a.Tsout, i, j, Tsout);
  stg<FT>(a.ust, i, j, us); stg<FT>(a.tst, i, j, ts); stg<FT>(a.qst, i, j, qs);
  stg<FT>(a.Ttop_out, i, j, Tsout);
  if (a.iters.p) reinterpret_cast<int32_t*>(a.iters.p)[(int64_t)i * a.iters.si + (int64_t)j * a.iters.sj] = act ? it : 0;
  if (a.seam_east && i == a.Nx - 1 && j >= 0 && j < a.Ny) reinterpret_cast<FT*>(a.seam_east)[j] = rtx;
}
template <typename FT, int TILE>
__global__ void __launch_bounds__(128, COFLUX_ICE_MIN_BLOCKS) ice_queue_kernel(const __grid_constant__ FluxArgs<FT> a) {
  __shared__ IceQueueSmem<TILE> sm;
  const DevParams<FT>& P = a.P;
  const FluxP<FT>& F = P.ai;
  const int tid = threadIdx.x;
  const int tile_n = (a.tile_cells > 0 && a.tile_cells < TILE) ? a.tile_cells : TILE;
  const long long tile0 = a.cell0 + (long long)blockIdx.x * tile_n;
  const int tile_jj0 = (int)(tile0 / a.nxr);
  const int tile_ii0 = (int)(tile0 - (long long)tile_jj0 * a.nxr);
  const unsigned nxr_u = (unsigned)a.nxr;
  if (tid == 0) { sm.n_queued = 0; sm.head = 0; }
  __syncthreads();

  // ------------------------------------------------------------------ phase A: init() of every cell, all lanes busy
  for (int cidx = tid; cidx < tile_n; cidx += 128) {
    if (tile0 + cidx >= a.ncell) break;
    const unsigned t = (unsigned)(tile_ii0 + cidx), dj = t / nxr_u;
    const int i = (int)(t - dj * nxr_u) - a.ring, j = tile_jj0 + (int)dj - a.ring;
    CellIn<FT> in;
    in.ua = ldg<FT>(a.xu, i, j); in.va = ldg<FT>(a.xv, i, j); in.Ta = ldg<FT>(a.xT, i, j); in.pa = ldg<FT>(a.xp, i, j);
    in.qa = ldg<FT>(a.xq, i, j); in.Qs = ldg<FT>(a.xQs, i, j); in.Ql = ldg<FT>(a.xQl, i, j);
    in.us = (ldg<FT>(a.ou, i, j) + ldg<FT>(a.ou, i + 1, j)) * FT(0.5);
    in.vs = (ldg<FT>(a.ov, i, j) + ldg<FT>(a.ov, i, j + 1)) * FT(0.5);
    in.Ts0 = ldg<FT>(a.oT, i, j) + P.T_offset;
    in.So = FT(0);
    in.h_ice = ldg<FT>(a.ih, i, j);
    in.S_ice = ldg<FT>(a.iS, i, j);
    in.albedo = sea_ice_albedo<FT>(P, a.ialb, a.ih, a.ihs, i, j, in.Ts0);
    const FT conc = ldg<FT>(a.iconc, i, j);
    const bool act = is_active(a.mask, i, j) && (conc > FT(0)) && (in.h_ice > FT(0));
    if (act) {
      CellSolver<FT, 1> s;
      s.init(P, F, in);
      if (s.go) {
        stg<FT>(a.Qv, i, j, s.atm.rho); stg<FT>(a.Qc, i, j, s.atm.cp_m); stg<FT>(a.Fv, i, j, s.atm.q_vap);
        stg<FT>(a.rtx, i, j, s.theta_a); stg<FT>(a.rty, i, j, s.du2dv2); stg<FT>(a.Tsout, i, j, in.albedo);
        sm.queue[atomicAdd(&sm.n_queued, 1)] = (unsigned short)cidx;
      } else {
        ice_finish_cell<FT>(a, i, j, true, s.ustar, s.tstar, s.qstar, s.Ts, s.atm.rho, s.atm.cp_m, in.Ta, s.it);
      }
    } else {
      ice_finish_cell<FT>(a, i, j, false, FT(0), FT(0), FT(0), in.Ts0, FT(0), FT(0), in.Ta, 0);
    }
  }
  __syncthreads();
  const int n_total = sm.n_queued;

  // ------------------------------------------------------------------ phase B: lane refill; a finished cell is stored by its lane
  CellSolver<FT, 1> s;
  int ci = 0, cj = 0;
  bool busy = false;
  {   // everything init() derives from the parameters alone
    CellIn<FT> z{};
    z.Ta = FT(280); z.pa = FT(1e5); z.qa = FT(1e-3); z.Ts0 = FT(270); z.h_ice = FT(1);
    s.in = z;
    s.x = FT(1);
    s.delta = P.th.eps - FT(1);
    s.ly = false; s.U_ly = FT(0); s.rcdn_ly = FT(0); s.lnh10 = FT(0);
    s.fixed = (F.stop_kind == COFLUX_STOP_FIXED_ITERATIONS);
    s.ice_fast = true;
    s.lnh_lu = M<FT>::log(P.h / F.mr.fixed); s.lnh_lt = M<FT>::log(P.h / F.tr.fixed); s.lnh_lq = M<FT>::log(P.h / F.qr.fixed);
    s.du = s.dv = FT(0);
  }
  auto pop = [&]() {
    const int pos = atomicAdd(&sm.head, 1);
    busy = false;
    if (pos < n_total) {
      busy = true;
      const unsigned t = (unsigned)(tile_ii0 + (int)sm.queue[pos]), dj = t / nxr_u;
      ci = (int)(t - dj * nxr_u) - a.ring; cj = tile_jj0 + (int)dj - a.ring;
      s.in.Ta = ldg<FT>(a.xT, ci, cj); s.in.pa = ldg<FT>(a.xp, ci, cj); s.in.Qs = ldg<FT>(a.xQs, ci, cj); s.in.Ql = ldg<FT>(a.xQl, ci, cj);
      s.in.S_ice = ldg<FT>(a.iS, ci, cj); s.in.h_ice = ldg<FT>(a.ih, ci, cj);
      s.atm.rho = ldc<FT>(a.Qv, ci, cj); s.atm.cp_m = ldc<FT>(a.Qc, ci, cj); s.atm.q_vap = ldc<FT>(a.Fv, ci, cj);
      s.theta_a = ldc<FT>(a.rtx, ci, cj); s.du2dv2 = ldc<FT>(a.rty, ci, cj); s.in.albedo = ldc<FT>(a.Tsout, ci, cj);
      s.ustar = F.init; s.tstar = F.init; s.qstar = F.init;
      s.Ts = ldg<FT>(a.oT, ci, cj) + P.T_offset;
      s.it = 0; s.go = true;
      s.su = s.ustar; s.st = s.tstar; s.sq = s.qstar; s.sT = s.Ts; s.sr = s.rcdn_ly;
      s.snap_it = 0; s.window = 1; s.stop_at = -1;
    }
  };
  pop();
  while (__any_sync(0xffffffffu, busy)) {
    if (busy) {
      s.pass(P, F);
      if (!s.go) {
        ice_finish_cell<FT>(a, ci, cj, true, s.ustar, s.tstar, s.qstar, s.Ts, s.atm.rho, s.atm.cp_m, s.in.Ta, s.it);
        pop();
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// centre → face momentum fluxes (A9)
// ---------------------------------------------------------------------------------------------
// by-products of the net fluxes for the ocean mixing closures (KPP/kpp_surface_forcing.jl:18-29,
// NEMOTKE/nemo_tke_surface_forcing.jl:14-22); all pointers optional
template <typename FT> struct ClosureArgs {
  int on;
  DArr JT, JS, alpha, beta;
  DArr ustar, ustar2, tke, Bo;
  FT umin, emin, Cb, g;
};
template <typename FT>
__device__ __forceinline__ void closure_forcing(const ClosureArgs<FT>& c, int i, int j, FT tx, FT ty) {
  const FT ustar2 = M<FT>::sqrt(tx * tx + ty * ty);                 // u★² = |τ| (kinematic stress)
  stg<FT>(c.ustar2, i, j, ustar2);
  stg<FT>(c.ustar, i, j, M<FT>::max(M<FT>::sqrt(ustar2), c.umin));   // max(√√(τx²+τy²), u★_min)
  stg<FT>(c.tke, i, j, M<FT>::max(c.emin, c.Cb * ustar2));
  if (c.Bo.p && c.alpha.p && c.beta.p && c.JT.p && c.JS.p) {
    const FT JT = ldg<FT>(c.JT, i, j), JS = ldg<FT>(c.JS, i, j);
    stg<FT>(c.Bo, i, j, -(c.g * (ldg<FT>(c.alpha, i, j) * JT - ldg<FT>(c.beta, i, j) * JS)));
  }
}

template <typename FT> struct StressArgs {
  int Nx, Ny, wrap_x;      // wrap_x: i-1 at i == 0 → Nx-1 (single-slab periodic, ring == 0)
  long long cell0, cell1;  // linear interior cell range [cell0, cell1) handled by this launch
  DArr rtx, rty, conc, tx_io, ty_io, mask;
  DArr taux, tauy;
  const char* seam_west;   // ρτx of the west neighbour's last column (Ny elements), or nullptr
  FT rho0, rho0inv;        // rho0inv = 1/ρ₀, divided on the host
  ClosureArgs<FT> closure;
  DArr avg_tx, avg_ty;     // running time averages of τx, τy (optional)
  FT avg_T, avg_dt;
};
template <typename FT>
__device__ __forceinline__ void assemble_stress(const StressArgs<FT>& a, int i, int j, FT& tx, FT& ty) {
  const FT rho0inv = a.rho0inv;
  const int iw = (a.wrap_x && i == 0) ? a.Nx - 1 : i - 1;
  const FT rtx_c = ldg<FT>(a.rtx, i, j);
  const FT rtx_w = (a.seam_west && i == 0) ? __ldg(reinterpret_cast<const FT*>(a.seam_west) + j) : ldg<FT>(a.rtx, iw, j);
  const FT rty_c = ldg<FT>(a.rty, i, j), rty_s = ldg<FT>(a.rty, i, j - 1);
  FT cx = FT(0), cy = FT(0);
  if (a.conc.p) {
    const FT c = ldg<FT>(a.conc, i, j);
    cx = FT(0.5) * (ldg<FT>(a.conc, iw, j) + c);
    cy = FT(0.5) * (ldg<FT>(a.conc, i, j - 1) + c);
  }
  const FT txao = (rtx_w + rtx_c) * FT(0.5) * rho0inv;
  const FT tyao = (rty_s + rty_c) * FT(0.5) * rho0inv;
  const FT txio = a.tx_io.p ? ldg<FT>(a.tx_io, i, j) * rho0inv * cx : FT(0);
  const FT tyio = a.ty_io.p ? ldg<FT>(a.ty_io, i, j) * rho0inv * cy : FT(0);
  tx = (FT(1) - cx) * txao + txio;
  ty = (FT(1) - cy) * tyao + tyio;
  const bool act = is_active(a.mask, i, j);
  if (!act || !is_active(a.mask, iw, j)) tx = FT(0);
  if (!act || !is_active(a.mask, i, j - 1)) ty = FT(0);
}
template <typename FT> __global__ void __launch_bounds__(256) stress_kernel(const __grid_constant__ StressArgs<FT> a) {
  const long long idx = a.cell0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= a.cell1) return;
  int i, j;
  if (a.cell1 < (1LL << 31)) { const unsigned u = (unsigned)idx, q = u / (unsigned)a.Nx; j = (int)q; i = (int)(u - q * (unsigned)a.Nx); }   // 32-bit division
  else { j = (int)(idx / a.Nx); i = (int)(idx - (long long)j * a.Nx); }
  FT tx, ty;
  assemble_stress<FT>(a, i, j, tx, ty);
  stg<FT>(a.taux, i, j, tx);
  stg<FT>(a.tauy, i, j, ty);
  if (a.closure.on) closure_forcing<FT>(a.closure, i, j, tx, ty);
  if (a.avg_tx.p || a.avg_ty.p) {          // both old means in flight before the first store
    FT* px = a.avg_tx.p ? reinterpret_cast<FT*>(a.avg_tx.p) + ((int64_t)i * a.avg_tx.si + (int64_t)j * a.avg_tx.sj) : nullptr;
    FT* py = a.avg_ty.p ? reinterpret_cast<FT*>(a.avg_ty.p) + ((int64_t)i * a.avg_ty.si + (int64_t)j * a.avg_ty.sj) : nullptr;
    const FT ox = px ? *px : FT(0), oy = py ? *py : FT(0);
    if (px) *px = (ox * a.avg_T + tx * a.avg_dt) / (a.avg_T + a.avg_dt);
    if (py) *py = (oy * a.avg_T + ty * a.avg_dt) / (a.avg_T + a.avg_dt);
  }
}
// stand-alone closure front end: τx, τy read back from the net fluxes
template <typename FT> struct ClosureKernelArgs {
  int Nx, Ny;
  DArr taux, tauy;
  ClosureArgs<FT> closure;
};
template <typename FT> __global__ void __launch_bounds__(256) closure_forcing_kernel(const __grid_constant__ ClosureKernelArgs<FT> a) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)a.Nx * a.Ny) return;
  const int j = (int)(idx / a.Nx), i = (int)(idx - (long long)j * a.Nx);
  closure_forcing<FT>(a.closure, i, j, ldg<FT>(a.taux, i, j), ldg<FT>(a.tauy, i, j));
}

// stand-alone compute_net_ocean_fluxes! (reads the interface fluxes back from memory)
template <typename FT> struct AssembleArgs {
  StressArgs<FT> s;
  DArr oS, Ts, xQs, xQl, xMp, Qc, Qv, Fv, Qio, salt_io;
  DArr JT, JS, Qu, Qal, Qts, J0;
  DevParams<FT> P;
};
template <typename FT> __global__ void __launch_bounds__(256) assemble_kernel(const __grid_constant__ AssembleArgs<FT> a) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)a.s.Nx * a.s.Ny) return;
  const int j = (int)(idx / a.s.Nx), i = (int)(idx - (long long)j * a.s.Nx);
  const FT conc = a.s.conc.p ? ldg<FT>(a.s.conc, i, j) : FT(0);
  const FT Qio = a.Qio.p ? ldg<FT>(a.Qio, i, j) : FT(0);
  const FT sio = a.salt_io.p ? ldg<FT>(a.salt_io, i, j) : FT(0);
  FT JT, JS, Qu, Qal, Qts, J0;
  assemble_tracers<FT>(a.P, is_active(a.s.mask, i, j), conc, ldg<FT>(a.oS, i, j), ldg<FT>(a.Ts, i, j) + a.P.T_offset,
                       ldg<FT>(a.xQs, i, j), ldg<FT>(a.xQl, i, j), ldg<FT>(a.xMp, i, j), ldg<FT>(a.Qc, i, j),
                       ldg<FT>(a.Qv, i, j), ldg<FT>(a.Fv, i, j), Qio, sio, JT, JS, Qu, Qal, Qts, J0);
  FT tx, ty;
  assemble_stress<FT>(a.s, i, j, tx, ty);
  stg<FT>(a.s.taux, i, j, tx); stg<FT>(a.s.tauy, i, j, ty);
  stg<FT>(a.JT, i, j, JT); stg<FT>(a.JS, i, j, JS); stg<FT>(a.Qu, i, j, Qu); stg<FT>(a.Qal, i, j, Qal);
  stg<FT>(a.Qts, i, j, Qts); stg<FT>(a.J0, i, j, J0);
}

// ---------------------------------------------------------------------------------------------
// sea-ice–ocean fluxes (A10): quadratic stress at the faces, frazil column sweep, interface heat
// (ice bath or three-equation), salt flux from thickness change.  HBM-bound: 2·Nz words/column.
// ---------------------------------------------------------------------------------------------
template <typename FT> struct IceOceanArgs {
  int Nx, Ny, Nz;
  DCol T, S;            // full columns; T is read and conditionally written
  DCol dz;              // si = sj = 0 for a 1-D Δz(k)
  DArr ou, ov;          // ocean u, v at k = Nz-1
  DArr iu, iv, ih, ihm, iconc, iS;
  DArr Qf, Qio, Js, tx, ty;
  DArr avg_JTf;         // running time average of the frazil temperature flux Q_f / (ρ₀ c₀) (optional)
  FT avg_T, avg_dt;
  FT dt;
  DevParams<FT> P;
};
template <typename FT>
__device__ __forceinline__ FT io_stress_x(const IceOceanArgs<FT>& a, int i, int j) {
  const FT dux = ldg<FT>(a.iu, i, j) - ldg<FT>(a.ou, i, j);
  const FT viF = FT(0.25) * (ldg<FT>(a.iv, i - 1, j) + ldg<FT>(a.iv, i, j) + ldg<FT>(a.iv, i - 1, j + 1) + ldg<FT>(a.iv, i, j + 1));
  const FT voF = FT(0.25) * (ldg<FT>(a.ov, i - 1, j) + ldg<FT>(a.ov, i, j) + ldg<FT>(a.ov, i - 1, j + 1) + ldg<FT>(a.ov, i, j + 1));
  const FT dvx = viF - voF;
  return a.P.rho0 * a.P.io.Cd * M<FT>::sqrt(dux * dux + dvx * dvx) * dux;
}
template <typename FT>
__device__ __forceinline__ FT io_stress_y(const IceOceanArgs<FT>& a, int i, int j) {
  const FT dvy = ldg<FT>(a.iv, i, j) - ldg<FT>(a.ov, i, j);
  const FT uiF = FT(0.25) * (ldg<FT>(a.iu, i, j - 1) + ldg<FT>(a.iu, i, j) + ldg<FT>(a.iu, i + 1, j - 1) + ldg<FT>(a.iu, i + 1, j));
  const FT uoF = FT(0.25) * (ldg<FT>(a.ou, i, j - 1) + ldg<FT>(a.ou, i, j) + ldg<FT>(a.ou, i + 1, j - 1) + ldg<FT>(a.ou, i + 1, j));
  const FT duy = uiF - uoF;
  return a.P.rho0 * a.P.io.Cd * M<FT>::sqrt(duy * duy + dvy * dvy) * dvy;
}

// levels loaded per batch of the frazil sweep (memory-level parallelism; tuned on B200 at Nz = 75:
// Float64 8 → 63 %, 12 → 70 %, 16 → 56 % of the measured HBM peak)
#ifndef COFLUX_IO_UNR64
#define COFLUX_IO_UNR64 12
#endif
#ifndef COFLUX_IO_UNR32
#define COFLUX_IO_UNR32 12
#endif
#ifndef COFLUX_IO_BLOCK
#define COFLUX_IO_BLOCK 128
#endif
template <typename FT> __global__ void __launch_bounds__(COFLUX_IO_BLOCK) ice_ocean_kernel(const __grid_constant__ IceOceanArgs<FT> a) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)a.Nx * a.Ny) return;
  const int j = (int)(idx / a.Nx), i = (int)(idx - (long long)j * a.Nx);
  const DevParams<FT>& P = a.P;
  const FT rho0 = P.rho0, c0 = P.c0, T0 = P.io.T0, m = P.io.slope;

  const FT taux = io_stress_x<FT>(a, i, j), tauy = io_stress_y<FT>(a, i, j);
  stg<FT>(a.tx, i, j, taux);
  stg<FT>(a.ty, i, j, tauy);

  // frazil sweep, top to bottom; loads batched UNR levels at a time for memory-level parallelism
  const int64_t base = (int64_t)i * a.T.si + (int64_t)j * a.T.sj;
  const int64_t baseS = (int64_t)i * a.S.si + (int64_t)j * a.S.sj;
  const int64_t basez = (int64_t)i * a.dz.si + (int64_t)j * a.dz.sj;
  FT* Tp = reinterpret_cast<FT*>(a.T.p);
  const FT* Sp = reinterpret_cast<const FT*>(a.S.p);
  const FT* zp = reinterpret_cast<const FT*>(a.dz.p);
  FT dQ = FT(0);
  FT TN = FT(0), SN = FT(0);
  constexpr int UNR = (sizeof(FT) == 8) ? COFLUX_IO_UNR64 : COFLUX_IO_UNR32;
  for (int k0 = a.Nz - 1; k0 >= 0; k0 -= UNR) {
    FT Tk[UNR], Sk[UNR], zk[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int k = k0 - u;
      if (k >= 0) {
        Tk[u] = Tp[base + (int64_t)k * a.T.sk];
        Sk[u] = __ldg(Sp + baseS + (int64_t)k * a.S.sk);
        zk[u] = __ldg(zp + basez + (int64_t)k * a.dz.sk);
      }
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const int k = k0 - u;
      if (k >= 0) {
        const FT Tm = T0 - m * Sk[u];
        const bool freezing = Tk[u] < Tm;
        const FT dE = rho0 * c0 * (Tm - Tk[u]);
        if (freezing) {
          Tp[base + (int64_t)k * a.T.sk] = Tm;
          dQ -= dE * zk[u] / a.dt;
        }
        if (k == a.Nz - 1) { TN = freezing ? Tm : Tk[u]; SN = Sk[u]; }
      }
    }
  }
  const FT conc = ldg<FT>(a.iconc, i, j), Si = ldg<FT>(a.iS, i, j);
  const FT Tm = T0 - m * SN;
  FT Qio;
  if (P.io.heat_flux == COFLUX_ICE_OCEAN_THREE_EQUATION) {
    FT ustar;
    if (P.io.friction == COFLUX_FRICTION_VELOCITY_MOMENTUM_BASED) {
      const FT txe = (i + 1 < a.Nx) ? io_stress_x<FT>(a, i + 1, j) : taux;
      const FT tyn = (j + 1 < a.Ny) ? io_stress_y<FT>(a, i, j + 1) : tauy;
      const FT tx = FT(0.5) * (taux + txe), ty = FT(0.5) * (tauy + tyn);
      ustar = M<FT>::sqrt(M<FT>::sqrt(tx * tx + ty * ty) / rho0);
      ustar = M<FT>::max(ustar, P.io.ustar_min);
    } else {
      ustar = P.io.ustar_const;
    }
    const FT gT = P.io.alpha_h * ustar, gS = P.io.alpha_s * ustar;
    const FT A = rho0 * c0 * gT / (P.io.rho_i * P.io.L_f);
    const FT qa = A * m;
    const FT qb = A * (TN - T0) - A * m * Si + gS;
    const FT qc = -(A * (TN - T0) * Si + gS * SN);
    const FT disc = qb * qb - FT(4) * qa * qc;
    const FT Sb = (-qb + M<FT>::sqrt(M<FT>::max(disc, FT(0)))) / (FT(2) * qa);
    const FT Tb = T0 - m * Sb;
    Qio = rho0 * c0 * gT * (TN - Tb) * conc;
  } else {
    const FT dE = rho0 * c0 * (Tm - TN);
    Qio = -dE * P.io.um_star * conc;
  }
  const FT h = ldg<FT>(a.ih, i, j);
  const FT hm = reinterpret_cast<const FT*>(a.ihm.p)[(int64_t)i * a.ihm.si + (int64_t)j * a.ihm.sj];
  const FT Js = (h - hm) / a.dt * (Si - SN);
  stg<FT>(a.Qf, i, j, dQ);
  avg_update<FT>(a.avg_JTf, i, j, dQ / rho0 / c0, a.avg_T, a.avg_dt);
  stg<FT>(a.Qio, i, j, Qio);
  stg<FT>(a.Js, i, j, Js);
  stg<FT>(a.ihm, i, j, h);
}

// ---------------------------------------------------------------------------------------------
// land freshwater, stand-alone: exchange Mp += rivers + icebergs (the fused kernels do this in their phase A)
// ---------------------------------------------------------------------------------------------
template <typename FT> __global__ void __launch_bounds__(256) land_kernel(const __grid_constant__ FluxArgs<FT> a) {
  const long long idx = a.cell0 + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= a.ncell) return;
  const int jj = (int)(idx / a.nxr);
  const int ii = (int)(idx - (long long)jj * a.nxr);
  const int i = ii - a.ring, j = jj - a.ring;
  FT* p = reinterpret_cast<FT*>(a.xMp.p) + ((int64_t)i * a.xMp.si + (int64_t)j * a.xMp.sj);
  *p += land_freshwater<FT>(a, i, j);
}

// ---------------------------------------------------------------------------------------------
// compute_net_sea_ice_fluxes!: heat into the ice from above and below; the atmosphere–ice stress at the
// velocity points of the ice model.  Pure streaming: ≈ 12 words read, 2–4 written per cell.
// ---------------------------------------------------------------------------------------------
template <typename FT> struct NetIceArgs {
  int Nx, Ny, wrap_x;
  DArr Qs, Ql;                    // exchange state
  DArr Ttop, conc, ih, ihs, ialb; // sea ice
  DArr Qc, Qv, rtx, rty;          // atmosphere–sea-ice interface fluxes
  DArr Qf, Qi;                    // sea-ice–ocean fluxes
  DArr mask;
  DArr top, bottom, top_u, top_v;
  DevParams<FT> P;
};
template <typename FT> __global__ void __launch_bounds__(256) net_sea_ice_kernel(const __grid_constant__ NetIceArgs<FT> a) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)a.Nx * a.Ny) return;
  const int j = (int)(idx / a.Nx), i = (int)(idx - (long long)j * a.Nx);
  const DevParams<FT>& P = a.P;
  const bool act = is_active(a.mask, i, j);
  const FT TsK = ldg<FT>(a.Ttop, i, j) + P.T_offset;
  const FT conc = ldg<FT>(a.conc, i, j);
  const FT alpha = sea_ice_albedo<FT>(P, a.ialb, a.ih, a.ihs, i, j, TsK);
  const FT Qu = P.emis_i * P.sigma * TsK * TsK * TsK * TsK;
  const FT Qd = -(FT(1) - alpha) * ldg<FT>(a.Qs, i, j) - P.emis_i * ldg<FT>(a.Ql, i, j);
  const FT SQt = (conc > FT(0)) ? (Qd + Qu + ldg<FT>(a.Qc, i, j) + ldg<FT>(a.Qv, i, j)) : FT(0);
  const FT SQb = (a.Qf.p ? ldg<FT>(a.Qf, i, j) : FT(0)) + (a.Qi.p ? ldg<FT>(a.Qi, i, j) : FT(0));
  stg<FT>(a.top, i, j, act ? SQt : FT(0));
  stg<FT>(a.bottom, i, j, act ? SQb : FT(0));
  if (a.top_u.p) {
    const int iw = (a.wrap_x && i == 0) ? a.Nx - 1 : i - 1;
    const FT v = (ldg<FT>(a.rtx, iw, j) + ldg<FT>(a.rtx, i, j)) * FT(0.5);
    stg<FT>(a.top_u, i, j, (act && is_active(a.mask, iw, j)) ? v : FT(0));
  }
  if (a.top_v.p) {
    const FT v = (ldg<FT>(a.rty, i, j - 1) + ldg<FT>(a.rty, i, j)) * FT(0.5);
    stg<FT>(a.top_v, i, j, (act && is_active(a.mask, i, j - 1)) ? v : FT(0));
  }
}

// ---------------------------------------------------------------------------------------------
// time-averaged flux diagnostics, stand-alone (reads the flux fields back; the fused path accumulates in its epilogues)
// ---------------------------------------------------------------------------------------------
template <typename FT> struct FluxAvgArgs {
  int Nx, Ny;
  DArr tx, ty, JT, JS, Qc, Qv, conc, Qio, salt_io, Qf;
  DArr a_tx, a_ty, a_JT, a_JS, a_Qc, a_Qv, a_JTao, a_JTio, a_JSio, a_JTf;
  FT T, dt, rho0, c0;
};
template <typename FT> __global__ void __launch_bounds__(256) flux_average_kernel(const __grid_constant__ FluxAvgArgs<FT> a) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)a.Nx * a.Ny) return;
  const int j = (int)(idx / a.Nx), i = (int)(idx - (long long)j * a.Nx);
  if (a.tx.p) avg_update<FT>(a.a_tx, i, j, ldg<FT>(a.tx, i, j), a.T, a.dt);
  if (a.ty.p) avg_update<FT>(a.a_ty, i, j, ldg<FT>(a.ty, i, j), a.T, a.dt);
  const FT JT = a.JT.p ? ldg<FT>(a.JT, i, j) : FT(0);
  if (a.JT.p) avg_update<FT>(a.a_JT, i, j, JT, a.T, a.dt);
  if (a.JS.p) avg_update<FT>(a.a_JS, i, j, ldg<FT>(a.JS, i, j), a.T, a.dt);
  if (a.Qc.p) avg_update<FT>(a.a_Qc, i, j, ldg<FT>(a.Qc, i, j), a.T, a.dt);
  if (a.Qv.p) avg_update<FT>(a.a_Qv, i, j, ldg<FT>(a.Qv, i, j), a.T, a.dt);
  const FT rho0inv = FT(1) / a.rho0;
  const FT JTio = a.Qio.p ? ldg<FT>(a.Qio, i, j) * rho0inv / a.c0 : FT(0);
  avg_update<FT>(a.a_JTio, i, j, JTio, a.T, a.dt);
  if (a.JT.p) avg_update<FT>(a.a_JTao, i, j, JT - JTio, a.T, a.dt);
  const FT conc = a.conc.p ? ldg<FT>(a.conc, i, j) : FT(0);
  avg_update<FT>(a.a_JSio, i, j, a.salt_io.p ? ldg<FT>(a.salt_io, i, j) * conc : FT(0), a.T, a.dt);
  if (a.Qf.p) avg_update<FT>(a.a_JTf, i, j, ldg<FT>(a.Qf, i, j) / a.rho0 / a.c0, a.T, a.dt);
}

// ---------------------------------------------------------------------------------------------
// sea-ice–ocean fluxes, bulk-asynchronous form (round 2).  The register-staged kernel above keeps 24 loads per thread in
// flight and stops at 69 % of the measured HBM peak: its in-flight bytes are bounded by registers (128 per thread, 16 warps
// per SM).  Here the T and S columns of a CTA's W adjacent cells are streamed through shared memory by the bulk-copy engine
// (cp.async.bulk global → shared, completion on an mbarrier: SASS UBLKCP), KB levels per stage, STAGES stages deep, issued
// by one warp and costing no registers: ≈ 3 × 66 KB in flight per SM.  A thread then reads its column from shared memory
// (consecutive words, conflict-free) top → bottom, exactly the arithmetic of ice_ocean_kernel → bit-identical results.
// Alignment: cp.async.bulk needs 16-byte aligned source, destination and size; a row segment of a halo-padded parent starts
// at an arbitrary element, so each copy starts at the aligned element at or before the segment (≤ 16/sizeof(FT) − 1 extra
// elements, inside the parent because every parent has a halo) and the consumer indexes past that shift.
// ---------------------------------------------------------------------------------------------
// Tile shape, A/B-measured on B200 at 1/12°, Nz = 75 (profiles/README.md; fraction of the measured HBM peak, Float64):
//   W × KB × STAGES   128×8×4 59 %   256×8×3 73 %   256×8×2 90 %   288×8×2 93 %   384×8×2 90 %   256×4×6 67 %   288×5×5 40 %
// Row segments of ≥ 2 KB per copy, few deep stages, three CTAs per SM.  (Register-staged kernel: 69 %.)
#ifndef COFLUX_IOB_W
#define COFLUX_IOB_W 288            /* Float64 (4320 = 15 × 288) */
#endif
#ifndef COFLUX_IOB_W32
#define COFLUX_IOB_W32 256          /* Float32 */
#endif
#ifndef COFLUX_IOB_KB
#define COFLUX_IOB_KB 8
#endif
#ifndef COFLUX_IOB_STAGES
#define COFLUX_IOB_STAGES 2
#endif
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
template <typename FT, int W, int KB, int STAGES> struct IceOceanBulkSmem {
  static constexpr int PADE = 16 / (int)sizeof(FT);             // elements per 16 bytes
  static constexpr int ROW = W + 2 * PADE;                      // aligned superset of a W-element segment
  alignas(128) FT T[STAGES][KB][ROW];
  alignas(128) FT S[STAGES][KB][ROW];
  alignas(8) unsigned long long full[STAGES];
};
template <typename FT, int W, int KB, int STAGES>
__global__ void __launch_bounds__(W) ice_ocean_bulk_kernel(const __grid_constant__ IceOceanArgs<FT> a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  using SM = IceOceanBulkSmem<FT, W, KB, STAGES>;
  SM& sm = *reinterpret_cast<SM*>(smem_raw);
  static_assert(2 * KB <= 32, "one warp issues the copies of a stage: one lane per (array, level)");
  const int tid = threadIdx.x;
  const int tiles_x = (a.Nx + W - 1) / W;
  const int j = (int)(blockIdx.x / tiles_x);
  const int i0 = (int)(blockIdx.x - (long long)j * tiles_x) * W;
  const int i = i0 + tid;
  const int wcols = min(W, a.Nx - i0);                          // columns of this tile
  const bool valid = tid < wcols;
  const DevParams<FT>& P = a.P;
  const FT rho0 = P.rho0, c0 = P.c0, T0 = P.io.T0, m = P.io.slope;
  const int nblk = (a.Nz + KB - 1) / KB;
  FT* Tp = reinterpret_cast<FT*>(a.T.p);
  const FT* Sp = reinterpret_cast<const FT*>(a.S.p);
  const int64_t rowT = (int64_t)i0 + (int64_t)j * a.T.sj, rowS = (int64_t)i0 + (int64_t)j * a.S.sj;   // element index of (i0, j, 0)

  if (tid == 0)
    for (int s = 0; s < STAGES; ++s) mbar_init(smem_u32(&sm.full[s]), 2 * KB);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();

  // one lane per (array, level) of a stage: aligned source, size, destination
  auto issue = [&](int blk, int s) {
    if (tid < 2 * KB) {
      const int which = tid / KB, u = tid - which * KB;
      const int k = a.Nz - 1 - blk * KB - u;
      const unsigned bar = smem_u32(&sm.full[s]);
      if (k >= 0) {
        const FT* src = which ? (Sp + rowS + (int64_t)k * a.S.sk) : (Tp + rowT + (int64_t)k * a.T.sk);
        const unsigned mis = (unsigned)((uintptr_t)src & 15u);                     // bytes past the aligned address
        const unsigned bytes = (unsigned)((mis + (unsigned)wcols * sizeof(FT) + 15u) & ~15u);
        mbar_arrive_expect_tx(bar, bytes);
        bulk_g2s(smem_u32(which ? &sm.S[s][u][0] : &sm.T[s][u][0]), reinterpret_cast<const char*>(src) - mis, bytes, bar);
      } else {
        mbar_arrive_expect_tx(bar, 0u);                                            // level below the bottom: nothing to copy
      }
    }
  };
  for (int b = 0; b < STAGES && b < nblk; ++b) issue(b, b);

  // 2-D terms of the column while the first stages land
  FT taux = FT(0), tauy = FT(0);
  if (valid) {
    taux = io_stress_x<FT>(a, i, j); tauy = io_stress_y<FT>(a, i, j);
    stg<FT>(a.tx, i, j, taux);
    stg<FT>(a.ty, i, j, tauy);
  }
  FT dQ = FT(0), TN = FT(0), SN = FT(0);
  const int64_t colT = (int64_t)i * a.T.si + (int64_t)j * a.T.sj;
  for (int b = 0; b < nblk; ++b) {
    const int s = b % STAGES;
    mbar_wait(smem_u32(&sm.full[s]), (unsigned)((b / STAGES) & 1));
    if (valid) {
#pragma unroll
      for (int u = 0; u < KB; ++u) {
        const int k = a.Nz - 1 - b * KB - u;
        if (k >= 0) {
          const unsigned misT = (unsigned)(((uintptr_t)(Tp + rowT + (int64_t)k * a.T.sk) & 15u) / sizeof(FT));
          const unsigned misS = (unsigned)(((uintptr_t)(Sp + rowS + (int64_t)k * a.S.sk) & 15u) / sizeof(FT));
          const FT Tk = sm.T[s][u][misT + tid], Sk = sm.S[s][u][misS + tid];
          const FT zk = __ldg(reinterpret_cast<const FT*>(a.dz.p) + ((int64_t)i * a.dz.si + (int64_t)j * a.dz.sj + (int64_t)k * a.dz.sk));
          const FT Tm = T0 - m * Sk;
          const bool freezing = Tk < Tm;
          const FT dE = rho0 * c0 * (Tm - Tk);
          if (freezing) {
            Tp[colT + (int64_t)k * a.T.sk] = Tm;
            dQ -= dE * zk / a.dt;
          }
          if (k == a.Nz - 1) { TN = freezing ? Tm : Tk; SN = Sk; }
        }
      }
    }
    __syncthreads();                                   // every thread is done with stage s: refill it
    if (b + STAGES < nblk) issue(b + STAGES, s);
  }
  if (!valid) return;
  const FT conc = ldg<FT>(a.iconc, i, j), Si = ldg<FT>(a.iS, i, j);
  const FT Tm = T0 - m * SN;
  FT Qio;
  if (P.io.heat_flux == COFLUX_ICE_OCEAN_THREE_EQUATION) {
    FT ustar;
    if (P.io.friction == COFLUX_FRICTION_VELOCITY_MOMENTUM_BASED) {
      const FT txe = (i + 1 < a.Nx) ? io_stress_x<FT>(a, i + 1, j) : taux;
      const FT tyn = (j + 1 < a.Ny) ? io_stress_y<FT>(a, i, j + 1) : tauy;
      const FT tx = FT(0.5) * (taux + txe), ty = FT(0.5) * (tauy + tyn);
      ustar = M<FT>::sqrt(M<FT>::sqrt(tx * tx + ty * ty) / rho0);
      ustar = M<FT>::max(ustar, P.io.ustar_min);
    } else {
      ustar = P.io.ustar_const;
    }
    const FT gT = P.io.alpha_h * ustar, gS = P.io.alpha_s * ustar;
    const FT A = rho0 * c0 * gT / (P.io.rho_i * P.io.L_f);
    const FT qa = A * m;
    const FT qb = A * (TN - T0) - A * m * Si + gS;
    const FT qc = -(A * (TN - T0) * Si + gS * SN);
    const FT disc = qb * qb - FT(4) * qa * qc;
    const FT Sb = (-qb + M<FT>::sqrt(M<FT>::max(disc, FT(0)))) / (FT(2) * qa);
    const FT Tb = T0 - m * Sb;
    Qio = rho0 * c0 * gT * (TN - Tb) * conc;
  } else {
    const FT dE = rho0 * c0 * (Tm - TN);
    Qio = -dE * P.io.um_star * conc;
  }
  const FT h = ldg<FT>(a.ih, i, j);
  const FT hm = reinterpret_cast<const FT*>(a.ihm.p)[(int64_t)i * a.ihm.si + (int64_t)j * a.ihm.sj];
  const FT Js = (h - hm) / a.dt * (Si - SN);
  stg<FT>(a.Qf, i, j, dQ);
  avg_update<FT>(a.avg_JTf, i, j, dQ / rho0 / c0, a.avg_T, a.avg_dt);
  stg<FT>(a.Qio, i, j, Qio);
  stg<FT>(a.Js, i, j, Js);
  stg<FT>(a.ihm, i, j, h);
}

// ---------------------------------------------------------------------------------------------
// NormalizeSalinity (omip_simulation.jl:187-220): area-weighted mean of the combined surface salinity flux,
// removed from the whole parent of the bulk-flux field.  HBM-bound streaming: 1–2 words read per cell for the
// sums, 1 read + 1 write per parent element for the subtraction.  The reduction is a fixed-order tree (thread →
// warp shuffle → CTA → one thread over the CTA partials), accumulated in Float64 for either precision.
// ---------------------------------------------------------------------------------------------
template <typename FT> struct SaltSumArgs {
  int Nx, Ny;
  DArr flux, add, area, mask;
  double* partial;      // [2 * gridDim.x]
};
// One (virtual) thread's share of the two sums: elements first, first + stride, …, four loads in flight; the order
// in which a thread adds its elements is fixed, so the result is bit-reproducible.  (i, j) advance incrementally —
// a 64-bit division per element would make this streaming kernel compute bound.
template <typename FT> __device__ __forceinline__ void salt_thread_sums(const SaltSumArgs<FT>& a, long long idx, long long stride, double& num, double& den) {
  const long long n = (long long)a.Nx * a.Ny;
  num = 0.0; den = 0.0;
  const int di = (int)(stride % a.Nx), dj = (int)(stride / a.Nx);
  int j = (int)(idx / a.Nx), i = (int)(idx - (long long)j * a.Nx);
  while (idx < n) {
    double f[4], A[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      f[u] = 0.0; A[u] = 0.0;
      if (idx < n && is_active(a.mask, i, j)) {
        f[u] = (double)ldg<FT>(a.flux, i, j);
        if (a.add.p) f[u] += (double)ldg<FT>(a.add, i, j);
        A[u] = (double)ldg<FT>(a.area, i, j);
      }
      idx += stride; i += di; j += dj;
      if (i >= a.Nx) { i -= a.Nx; ++j; }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) { num = fma(f[u], A[u], num); den += A[u]; }
  }
}
// warp shuffle → CTA (8 warps) → partial[vb]; fixed order
__device__ __forceinline__ void salt_block_partial(double num, double den, double* sn, double* sd, double* partial, int vb) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { num += __shfl_down_sync(0xffffffffu, num, o); den += __shfl_down_sync(0xffffffffu, den, o); }
  if ((threadIdx.x & 31) == 0) { sn[threadIdx.x >> 5] = num; sd[threadIdx.x >> 5] = den; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double tn = 0.0, td = 0.0;
    for (int w = 0; w < 8; ++w) { tn += sn[w]; td += sd[w]; }
    partial[2 * vb] = tn; partial[2 * vb + 1] = td;
  }
}
template <typename FT> __global__ void __launch_bounds__(256) salt_sums_kernel(const __grid_constant__ SaltSumArgs<FT> a) {
  double num, den;
  salt_thread_sums<FT>(a, (long long)blockIdx.x * blockDim.x + threadIdx.x, (long long)gridDim.x * blockDim.x, num, den);
  __shared__ double sn[8], sd[8];
  salt_block_partial(num, den, sn, sd, a.partial, blockIdx.x);
}
// one warp, fixed order: lane l adds partials l, l+32, …; then a shuffle tree
__device__ __forceinline__ void salt_reduce_partials(const double* partial, int nblocks, double& num, double& den) {
  num = 0.0; den = 0.0;
  for (int b = threadIdx.x & 31; b < nblocks; b += 32) { num += partial[2 * b]; den += partial[2 * b + 1]; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { num += __shfl_down_sync(0xffffffffu, num, o); den += __shfl_down_sync(0xffffffffu, den, o); }
}
__global__ void __launch_bounds__(32) salt_sums_final_kernel(const double* partial, int nblocks, double* sums) {
  double num, den;
  salt_reduce_partials(partial, nblocks, num, den);
  if (threadIdx.x == 0) { sums[0] = num; sums[1] = den; }
}
template <typename FT> struct SubMeanArgs {
  char* p;              // first element of the parent
  int64_t si, sj;
  int ni, nj;           // parent extents (interior + 2 halos)
  const double* sums;   // {Σ f·Az, Σ Az}, or (nblocks > 0) the CTA partials of salt_sums_kernel
  int nblocks;
};
// parent element idx, idx + stride, …: four loads in flight, then four stores
template <typename FT> __device__ __forceinline__ void subtract_thread(const SubMeanArgs<FT>& a, FT mean, long long idx, long long stride) {
  const long long n = (long long)a.ni * a.nj;
  const int di = (int)(stride % a.ni), dj = (int)(stride / a.ni);
  int j = (int)(idx / a.ni), i = (int)(idx - (long long)j * a.ni);
  while (idx < n) {
    FT v[4];
    FT* q[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      q[u] = nullptr;
      if (idx < n) {
        q[u] = reinterpret_cast<FT*>(a.p) + ((int64_t)i * a.si + (int64_t)j * a.sj);
        v[u] = *q[u];
      }
      idx += stride; i += di; j += dj;
      if (i >= a.ni) { i -= a.ni; ++j; }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) if (q[u]) *q[u] = v[u] - mean;
  }
}
template <typename FT> __global__ void __launch_bounds__(256) subtract_mean_kernel(const __grid_constant__ SubMeanArgs<FT> a) {
  __shared__ double tot[2];
  if (a.nblocks > 0) {                 // single-slab form: every CTA reduces the partials itself (same order as the
    if (threadIdx.x < 32) {            // final kernel → same bits), which saves a launch
      double num, den;
      salt_reduce_partials(a.sums, a.nblocks, num, den);
      if (threadIdx.x == 0) { tot[0] = num; tot[1] = den; }
    }
  } else if (threadIdx.x == 0) {
    tot[0] = a.sums[0]; tot[1] = a.sums[1];
  }
  __syncthreads();
  const double den = tot[1];
  const FT mean = (den != 0.0) ? (FT)(tot[0] / den) : FT(0);
  subtract_thread<FT>(a, mean, (long long)blockIdx.x * blockDim.x + threadIdx.x, (long long)gridDim.x * blockDim.x);
}

// NormalizeSalinity in ONE cooperative launch: every CTA sums `vblocks / gridDim.x` virtual blocks of the two-launch
// form (same partials, same order → same bits), a grid barrier, then every CTA reduces the partials itself and
// subtracts the mean from its share of the parent.  The second pass finds the flux field in L2 (62 MB of 126 MB at
// 1/12° F64), so DRAM sees 17 B/cell of reads and 8 B/cell of writes.  The barrier is a self-resetting
// count + generation pair, so a captured graph can replay the launch.
template <typename FT> struct NormalizeArgs {
  SaltSumArgs<FT> s;
  SubMeanArgs<FT> m;
  int vblocks;
  unsigned* bar;        // {arrived, generation}
};
template <typename FT> __global__ void __launch_bounds__(256) normalize_salinity_kernel(const __grid_constant__ NormalizeArgs<FT> a) {
  __shared__ double sn[8], sd[8], tot[2];
  const int V = a.vblocks / (int)gridDim.x;
  const long long vstride = (long long)a.vblocks * blockDim.x;
  for (int v = 0; v < V; ++v) {
    const int vb = (int)blockIdx.x * V + v;
    double num, den;
    salt_thread_sums<FT>(a.s, (long long)vb * blockDim.x + threadIdx.x, vstride, num, den);
    salt_block_partial(num, den, sn, sd, a.s.partial, vb);
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    volatile unsigned* gen = a.bar + 1;
    const unsigned g = *gen;           // cannot move on before this CTA has arrived
    __threadfence();
    if (atomicAdd(a.bar, 1u) == gridDim.x - 1) {
      a.bar[0] = 0;
      __threadfence();
      atomicAdd(a.bar + 1, 1u);
    } else {
      while (*gen == g) __nanosleep(64);
    }
    __threadfence();
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    double num = 0.0, den = 0.0;
    for (int b = threadIdx.x; b < a.vblocks; b += 32) { num += __ldcg(a.s.partial + 2 * b); den += __ldcg(a.s.partial + 2 * b + 1); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { num += __shfl_down_sync(0xffffffffu, num, o); den += __shfl_down_sync(0xffffffffu, den, o); }
    if (threadIdx.x == 0) { tot[0] = num; tot[1] = den; }
  }
  __syncthreads();
  const double den = tot[1];
  const FT mean = (den != 0.0) ? (FT)(tot[0] / den) : FT(0);
  subtract_thread<FT>(a.m, mean, (long long)blockIdx.x * blockDim.x + threadIdx.x, (long long)gridDim.x * blockDim.x);
}

}  // namespace coflux

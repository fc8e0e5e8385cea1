// coflux_solve_stream.cuh — persistent, warp-specialised form of the atmosphere–ocean flux kernel (round 2).
//
// What the tile kernel (coflux_solve_tile.cuh) still lost after its ψ table moved to shared memory (ncu, round 2,
// profiles/README.md): every tile DRAINS — a lane handles ≈ 3 cells, the last cells of a tile run with most lanes of
// their warps idle (25 of 32 lanes active on average in phase B), and the warps then wait for each other at the closing
// barrier (13 % of the stall samples), after which the memory-latency-bound phases A and C run with nothing to overlap
// them inside the CTA.  Here ONE CTA per SM lives for the whole launch and its warps are specialised:
//
//   service warps   work on CHUNKS of 32 consecutive cells, lane ↔ cell, coalesced:
//                     A  loads, atmosphere interpolation, exchange-state stores, both thermodynamic states, the first
//                        similarity pass (always the stable block: lock step, all lanes busy); the cell's invariants go
//                        to a shared-memory slot, its slot id to the queue of its stability class;
//                     C  when the last cell of a chunk has converged: fluxes, net-flux assembly, coalesced stores;
//   solver warps    lanes pop cells from ONE class queue (a warp serves one class at a time, so its lanes share the
//                   code path), iterate, write the result to the slot and pop again.  The queues never run dry in steady
//                   state, so a warp has idle lanes only at the very end of the launch.
//
// Chunk buffers cycle FREE → filled by A → cells queued / in flight → all done → C → FREE.  All hand-offs are shared-
// memory ring queues: a producer reserves entries with one atomicAdd per warp and then writes them; a consumer reserves
// with a compare-and-swap per warp and waits for the (already reserved) entry to turn valid.  __threadfence_block()
// orders slot data before the queue entry / completion counter that publishes it.  There is no CTA-wide barrier after
// start-up and no dependence between CTAs (chunks are dealt round-robin: chunk g belongs to CTA g mod gridDim.x).
//
// Arithmetic per cell is exactly that of the tile kernel (same functions, same order): results are bit-identical.
#pragma once
#include "coflux_solve_tile.cuh"

namespace coflux {

#ifndef COFLUX_STREAM_NT
#define COFLUX_STREAM_NT 768          /* threads per CTA (one CTA per SM) */
#endif
#ifndef COFLUX_STREAM_SERVICE
#define COFLUX_STREAM_SERVICE 7       /* service warps of the 24 (phases A + C are ≈ 27 % of the instructions, latency bound) */
#endif
#ifndef COFLUX_STREAM_BUFS64
#define COFLUX_STREAM_BUFS64 48       /* chunk buffers (32 cells each) in Float64: 1536 slots */
#endif
#ifndef COFLUX_STREAM_BUFS32
#define COFLUX_STREAM_BUFS32 56
#endif

template <typename FT, int NB, bool VARNU> struct StreamSmem {
  static constexpr int NS = NB * 32;                       // cell slots
  static constexpr int QCAP = 2048, QMASK = QCAP - 1;      // class queues (≥ NS entries, power of two)
  static constexpr int CCAP = 64, CMASK = CCAP - 1;        // chunk-ready queue (≥ NB)
  static_assert(NS <= QCAP && NB <= CCAP && NB <= 64, "queue capacities");
  // cell slots (SoA: lane l of a chunk buffer b owns slot 32 b + l → consecutive words, no bank conflicts in A / C)
  //   queued      U2 dth dq c1 c2 (+nu inu): invariants;  us1 chi1: state after the first pass (us1 < 0: not run)
  //   in flight   U2 dth dq: Brent snapshot, c1: packed Brent state (the lane holds the invariants in registers)
  //   finished    U2 dth dq ← u★ θ★ q★,  c1 ← iteration count
  //   du dv rho cp Ta: carried from phase A to phase C
  FT U2[NS], dth[NS], dq[NS], c1[NS], c2[NS], us1[NS], chi1[NS];
  FT du[NS], dv[NS], rho[NS], cp[NS], Ta[NS];
  FT nu[VARNU ? NS : 1], inu[VARNU ? NS : 1];
  int chunk_id[NB];                 // CTA-local chunk number held by the buffer
  int n_queued[NB], n_done[NB];
  unsigned short q[2][QCAP];        // [0] unstable, [1] stable: slot ids, 0xFFFF = not yet written
  unsigned short qc[CCAP];          // chunk buffers whose cells have all converged
  unsigned int head[2], tail[2], headc, tailc;
  unsigned long long free_mask;     // bit b: chunk buffer b is free
  int next_chunk, a_done, c_done;
};

template <typename FT, int SPEC> struct StreamTraits {
  static constexpr bool F64 = (sizeof(FT) == 8);
  static constexpr bool VARNU = (SPEC != 1);
  static constexpr int NT = COFLUX_STREAM_NT;
  static constexpr int NW = NT / 32;
  static constexpr int NSERVICE = COFLUX_STREAM_SERVICE;
  static constexpr int NB = F64 ? COFLUX_STREAM_BUFS64 : COFLUX_STREAM_BUFS32;
  static constexpr int PSI_BYTES = COFLUX_PSI_SM_ROWS * 16 * (int)sizeof(FT);
  static_assert(SPEC == 1 || SPEC == 2, "the streaming kernel runs the lean pass (OMIP parameter sets)");
  static_assert(NSERVICE >= 1 && NSERVICE < NW, "need service and solver warps");
};

// volatile views of the shared scheduler words
__device__ __forceinline__ unsigned ld_vol(const unsigned* p) { return *reinterpret_cast<const volatile unsigned*>(p); }
__device__ __forceinline__ int ld_vol(const int* p) { return *reinterpret_cast<const volatile int*>(p); }
__device__ __forceinline__ unsigned long long ld_vol(const unsigned long long* p) { return *reinterpret_cast<const volatile unsigned long long*>(p); }

template <typename FT, bool INTERP, bool ASSEMBLE, int SPEC>
__global__ void __launch_bounds__(StreamTraits<FT, SPEC>::NT, 1) flux_stream_kernel(const __grid_constant__ FluxArgs<FT> a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  using TT = StreamTraits<FT, SPEC>;
  constexpr bool VARNU = TT::VARNU;
  constexpr bool TABS = TT::F64;
  constexpr int NB = TT::NB;
  using SM = StreamSmem<FT, NB, VARNU>;
  using MP = std::conditional_t<TABS, MLeanD, M<FT>>;
  using SI = SlotInt<FT>;
  SM& sm = *reinterpret_cast<SM*>(smem_raw);
  __shared__ __align__(16) double s_lgt[TABS ? 256 : 2];
  __shared__ double s_ext[TABS ? 64 : 2];
  __shared__ __align__(128) unsigned char s_psi[TT::PSI_BYTES];
  const DevParams<FT>& P = a.P;
  const FluxP<FT>& F = P.ao;
  const ThermoC<FT>& c = P.th;
  const FastConsts<FT>& K = P.K;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr unsigned FULL = 0xffffffffu;

  // chunks of 32 cells, dealt round-robin to the CTAs
  const long long ncells = a.ncell - a.cell0;
  const int nchunks = (int)((ncells + 31) / 32);
  const int nchunks_cta = (nchunks > (int)blockIdx.x) ? (nchunks - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  // ---------------------------------------------------------------- start-up (the only CTA-wide barrier)
  if (tid == 0) {
    sm.head[0] = sm.head[1] = sm.tail[0] = sm.tail[1] = 0; sm.headc = sm.tailc = 0;
    sm.free_mask = (NB == 64) ? ~0ull : ((1ull << NB) - 1ull);
    sm.next_chunk = 0; sm.a_done = 0; sm.c_done = 0;
  }
  for (int k = tid; k < SM::QCAP; k += TT::NT) { sm.q[0][k] = 0xFFFF; sm.q[1][k] = 0xFFFF; }
  if (tid < SM::CCAP) sm.qc[tid] = 0xFFFF;
  if (TABS) {
    for (int k = tid; k < 256; k += TT::NT) s_lgt[k] = (&COFLUX_LOG_TABLE[0][0])[k];
    if (tid < 64) s_ext[tid] = COFLUX_EXP_TABLE[tid];
  }
  if constexpr (sizeof(FT) == 8) {
    const double2* src = reinterpret_cast<const double2*>(&COFLUX_PSI_TABLE_F64[COFLUX_PSI_SM_ROW0][0][0]);
    double2* dst = reinterpret_cast<double2*>(s_psi);
    for (int k = tid; k < COFLUX_PSI_SM_ROWS * 8; k += TT::NT) { const int r = k >> 3, p = k & 7; dst[(r << 3) + (p ^ (r & 7))] = __ldg(src + k); }
  } else {
    const float4* src = reinterpret_cast<const float4*>(&COFLUX_PSI_TABLE_F32[COFLUX_PSI_SM_ROW0][0][0]);
    float4* dst = reinterpret_cast<float4*>(s_psi);
    for (int k = tid; k < COFLUX_PSI_SM_ROWS * 4; k += TT::NT) { const int r = k >> 2, p = k & 3; dst[(r << 2) + (p ^ ((r >> 1) & 3))] = __ldg(src + k); }
  }
  __syncthreads();
  const LeanTabs tb{s_lgt, s_ext, (unsigned)__cvta_generic_to_shared(s_psi)};
  const FT delta = c.eps - FT(1);
  const bool fixed = (F.stop_kind == COFLUX_STOP_FIXED_ITERATIONS);

  // publish a chunk buffer whose cells are all finished
  auto push_ready = [&](int buf) {
    __threadfence_block();          // everything observed through the completion counter before the entry that publishes it
    const unsigned pos = atomicAdd(&sm.tailc, 1u) & SM::CMASK;
    *reinterpret_cast<volatile unsigned short*>(&sm.qc[pos]) = (unsigned short)buf;
  };

  if (warp < TT::NSERVICE) {
    // ================================================================ service warps: phases A and C, chunk-wise
    auto phase_c = [&](int buf) {
      const int slot = buf * 32 + lane;
      const long long idx = a.cell0 + ((long long)blockIdx.x + (long long)sm.chunk_id[buf] * gridDim.x) * 32 + lane;
      if (idx < a.ncell) {
        const int jj = (int)(idx / a.nxr);
        const int ii = (int)(idx - (long long)jj * a.nxr);
        const int i = ii - a.ring, j = jj - a.ring;
        const FT Tunits = ldg<FT>(a.oT, i, j);
        const bool act = is_active(a.mask, i, j);
        FT Qv = FT(0), Qc = FT(0), Fv = FT(0), rtx = FT(0), rty = FT(0);
        const FT us = sm.U2[slot], ts = sm.dth[slot], qs = sm.dq[slot];
        if (act) {
          const FT du = sm.du[slot], dv = sm.dv[slot], rho = sm.rho[slot], cp = sm.cp[slot], Ta = sm.Ta[slot];
          const FT d2 = du * du + dv * dv;
          const FT k = (d2 > FT(1e-30)) ? (-us * us) * fm::rcp(fm::sqrt(d2)) : FT(0);   // −u★²/‖Δu‖ (0 when calm)
          const FT taux = (d2 > FT(1e-30)) ? k * du : ((d2 == FT(0)) ? FT(0) : -us * us * du / M<FT>::sqrt(d2));
          const FT tauy = (d2 > FT(1e-30)) ? k * dv : ((d2 == FT(0)) ? FT(0) : -us * us * dv / M<FT>::sqrt(d2));
          const FT LH = c.LH_v0 + (c.cp_v - c.cp_l) * (Ta - c.T_0);
          Qv = -rho * us * qs * LH;
          Qc = -rho * cp * us * ts;
          Fv = -rho * us * qs;
          rtx = rho * taux; rty = rho * tauy;
        }
        stg<FT>(a.Qv, i, j, Qv); stg<FT>(a.Qc, i, j, Qc); stg<FT>(a.Fv, i, j, Fv);
        stg<FT>(a.rtx, i, j, rtx); stg<FT>(a.rty, i, j, rty); stg<FT>(a.Tsout, i, j, Tunits);
        stg<FT>(a.ust, i, j, us); stg<FT>(a.tst, i, j, ts); stg<FT>(a.qst, i, j, qs);
        if (a.iters.p) reinterpret_cast<int32_t*>(a.iters.p)[(int64_t)i * a.iters.si + (int64_t)j * a.iters.sj] = act ? SI::get(sm.c1[slot]) : 0;
        if (a.seam_east && i == a.Nx - 1 && j >= 0 && j < a.Ny) reinterpret_cast<FT*>(a.seam_east)[j] = rtx;
        if (ASSEMBLE) {
          if (i >= 0 && i < a.Nx && j >= 0 && j < a.Ny) {
            // the exchange state of this cell was stored by a lane of THIS CTA in phase A, published through the
            // shared-memory queues (fences on both sides); .cg loads take it from L2, where the stores went
            const FT Qs = __ldcg(reinterpret_cast<const FT*>(a.xQs.p) + ((int64_t)i * a.xQs.si + (int64_t)j * a.xQs.sj));
            const FT Ql = __ldcg(reinterpret_cast<const FT*>(a.xQl.p) + ((int64_t)i * a.xQl.si + (int64_t)j * a.xQl.sj));
            const FT Mp = __ldcg(reinterpret_cast<const FT*>(a.xMp.p) + ((int64_t)i * a.xMp.si + (int64_t)j * a.xMp.sj));
            const FT So = ldg<FT>(a.oS, i, j);
            const FT conc = a.conc.p ? ldg<FT>(a.conc, i, j) : FT(0);
            const FT Qio = a.Qio.p ? ldg<FT>(a.Qio, i, j) : FT(0);
            const FT sio = a.salt_io.p ? ldg<FT>(a.salt_io, i, j) : FT(0);
            FT JT, JS, Qu, Qal, Qts, J0, parts[3];
            assemble_tracers<FT>(P, act, conc, So, Tunits + P.T_offset, Qs, Ql, Mp, Qc, Qv, Fv, Qio, sio, JT, JS, Qu, Qal, Qts, J0, parts);
            stg<FT>(a.JT, i, j, JT); stg<FT>(a.JS, i, j, JS); stg<FT>(a.Qu, i, j, Qu); stg<FT>(a.Qal, i, j, Qal);
            stg<FT>(a.Qts, i, j, Qts); stg<FT>(a.J0, i, j, J0);
            if (a.avg.on) avg_epilogue<FT>(a.avg, i, j, JT, JS, Qc, Qv, parts);
          }
        }
      }
      __syncwarp();
      if (lane == 0) {
        atomicOr(&sm.free_mask, 1ull << buf);
        atomicAdd(&sm.c_done, 1);
      }
    };

    auto phase_a = [&](int buf, int k) {
      const int slot = buf * 32 + lane;
      const long long idx = a.cell0 + ((long long)blockIdx.x + (long long)k * gridDim.x) * 32 + lane;
      bool queued = false, unstable = false;
      if (idx < a.ncell) {
        const int jj = (int)(idx / a.nxr);
        const int ii = (int)(idx - (long long)jj * a.nxr);
        const int i = ii - a.ring, j = jj - a.ring;
        FT ua, va, Ta, pa, qa;
        if (INTERP) {
          const FT fi = ldgs<FT>(a.fi, i, j), fj = ldgs<FT>(a.fj, i, j);
          const int i0 = (int)M<FT>::trunc(fi), j0 = (int)M<FT>::trunc(fj);
          const int i1 = i0 + ((fi > FT(0)) - (fi < FT(0))), j1 = j0 + ((fj > FT(0)) - (fj < FT(0)));
          const FT xi = fi - M<FT>::floor(fi), eta = fj - M<FT>::floor(fj);
          const FT w00 = (FT(1) - xi) * (FT(1) - eta), w01 = (FT(1) - xi) * eta, w10 = xi * (FT(1) - eta), w11 = xi * eta;
          ua = interp_series<FT>(a.su, i0, j0, i1, j1, w00, w01, w10, w11, a.nfrac);
          va = interp_series<FT>(a.sv, i0, j0, i1, j1, w00, w01, w10, w11, a.nfrac);
          Ta = interp_series<FT>(a.sT, i0, j0, i1, j1, w00, w01, w10, w11, a.nfrac);
          qa = interp_series<FT>(a.sq, i0, j0, i1, j1, w00, w01, w10, w11, a.nfrac);
          pa = interp_series<FT>(a.sp, i0, j0, i1, j1, w00, w01, w10, w11, a.nfrac);
          const FT Qs = interp_series<FT>(a.sQs, i0, j0, i1, j1, w00, w01, w10, w11, a.nfrac);
          const FT Ql = interp_series<FT>(a.sQl, i0, j0, i1, j1, w00, w01, w10, w11, a.nfrac);
          FT Mp = FT(0);
          if (a.srain.p1) Mp += interp_series<FT>(a.srain, i0, j0, i1, j1, w00, w01, w10, w11, a.nfrac);
          if (a.ssnow.p1) Mp += interp_series<FT>(a.ssnow, i0, j0, i1, j1, w00, w01, w10, w11, a.nfrac);
          if (a.lfi.p) Mp += land_freshwater<FT>(a, i, j);
          if (a.cs.p && a.sn.p) {
            const FT cs = ldgs<FT>(a.cs, i, j), sn = ldgs<FT>(a.sn, i, j);
            const FT ur = ua * cs + va * sn, vr = -ua * sn + va * cs;
            ua = ur; va = vr;
          }
          stg<FT>(a.xu, i, j, ua); stg<FT>(a.xv, i, j, va); stg<FT>(a.xT, i, j, Ta); stg<FT>(a.xp, i, j, pa);
          stg<FT>(a.xq, i, j, qa); stg<FT>(a.xQs, i, j, Qs); stg<FT>(a.xQl, i, j, Ql); stg<FT>(a.xMp, i, j, Mp);
        } else {
          ua = ldgs<FT>(a.xu, i, j); va = ldgs<FT>(a.xv, i, j); Ta = ldgs<FT>(a.xT, i, j); pa = ldgs<FT>(a.xp, i, j);
          qa = ldgs<FT>(a.xq, i, j);
        }
        const bool act = is_active(a.mask, i, j);
        bool finished_in_a = false;
        if (act) {
          const FT uo = (ldgs<FT>(a.ou, i, j) + ldgs<FT>(a.ou, i + 1, j)) * FT(0.5);
          const FT vo = (ldgs<FT>(a.ov, i, j) + ldgs<FT>(a.ov, i, j + 1)) * FT(0.5);
          const FT Ts = ldgs<FT>(a.oT, i, j) + P.T_offset;
          const FT So = ldgs<FT>(a.oS, i, j);
          FT du, dv;
          if (F.velocity == COFLUX_VELOCITY_RELATIVE) { du = ua - uo; dv = va - vo; } else { du = ua; dv = va; }
          const FT U2 = du * du + dv * dv;
          const Thermo<FT> atm = phase_equil_pTq<FT, MP>(c, pa, Ta, qa);
          const FT s = MP::div(So, FT(1000));
          const FT x = MP::div(FT(1) - s, FT(1) - s + P.wmf_alpha * s);
          const FT theta_a = Ta + MP::div(P.g * P.h, atm.cp_m);
          const SurfaceState<FT> S = surface_state<FT, 0, MP>(P, F, atm, pa, theta_a, x, Ts);
          sm.du[slot] = du; sm.dv[slot] = dv; sm.rho[slot] = atm.rho; sm.cp[slot] = atm.cp_m; sm.Ta[slot] = Ta;
          if (fixed ? (F.maxit > 0) : true) {
            queued = true;
            const FT gTv = P.g * fm::rcp(S.T_v);               // b★ = c1·θ★ + c2·q★
            const FT c1 = gTv * (FT(1) + delta * S.q_vap), c2 = gTv * (delta * S.T_v);
            const FT inv_nu = VARNU ? fm::rcp(S.nu_m) : K.inv_nu;
            if (VARNU) { sm.nu[slot] = S.nu_m; sm.inu[slot] = inv_nu; }
            // first pass, in lock step (every lane busy, one code path)
            FT us = F.init, ts = F.init, qs = F.init, chi = FT(0);
            bool pre = F.init > FT(0) && (c1 + c2) > FT(0);
            if (pre) {
              LeanCell<FT> lc;
              lc.U2 = U2; lc.dth = S.dtheta; lc.dq = S.dq; lc.cb1 = c1; lc.cb2 = c2;
              { const FT v = fm::fma_(F.ugmin, F.ugmin, U2); lc.Ustab = (v > FT(0)) ? fm::sqrt(v) : FT(0); }
              lc.bnu = VARNU ? F.mr.beta_s * S.nu_m : K.bnu; lc.inv_nu = inv_nu;
              pre = iterate_lean<FT, SPEC, true>(P, F, K, tb, lc, us, ts, qs, &chi);
            }
            if (pre && !keep_going<FT>(F, 1, us, ts, qs, F.init, F.init, F.init)) {     // done after one pass
              queued = false; finished_in_a = true;
              sm.U2[slot] = us; sm.dth[slot] = ts; sm.dq[slot] = qs; SI::set(sm.c1[slot], 1);
            }
            if (queued) {
              sm.us1[slot] = pre ? us : FT(-1); sm.chi1[slot] = chi;
              sm.U2[slot] = U2; sm.dth[slot] = S.dtheta; sm.dq[slot] = S.dq; sm.c1[slot] = c1; sm.c2[slot] = c2;
              unstable = (S.dtheta * c1 + c2 * S.dq) < FT(0);   // sign of the buoyancy scale of every later pass
            }
          }
        }
        if (!queued && !finished_in_a) {      // land, or a zero-pass solve
          const FT r0 = act ? F.init : FT(0);
          sm.U2[slot] = r0; sm.dth[slot] = r0; sm.dq[slot] = r0; SI::set(sm.c1[slot], 0);
        }
      }
      const unsigned mq = __ballot_sync(FULL, queued), mu = __ballot_sync(FULL, queued && unstable);
      const unsigned ms = mq & ~mu;
      if (lane == 0) { sm.chunk_id[buf] = k; sm.n_queued[buf] = __popc(mq); sm.n_done[buf] = 0; }
      __syncwarp();
      if (mq == 0u) { if (lane == 0) push_ready(buf); return; }   // nothing to solve in this chunk: straight to phase C
      unsigned bu = 0, bs = 0;
      if (lane == 0) {
        if (mu) bu = atomicAdd(&sm.tail[0], (unsigned)__popc(mu));
        if (ms) bs = atomicAdd(&sm.tail[1], (unsigned)__popc(ms));
      }
      bu = __shfl_sync(FULL, bu, 0); bs = __shfl_sync(FULL, bs, 0);
      __threadfence_block();                               // slot data (and the chunk counters) before the queue entries
      if (queued) {
        const unsigned lt = (1u << lane) - 1u;
        if (unstable) *reinterpret_cast<volatile unsigned short*>(&sm.q[0][(bu + __popc(mu & lt)) & SM::QMASK]) = (unsigned short)slot;
        else *reinterpret_cast<volatile unsigned short*>(&sm.q[1][(bs + __popc(ms & lt)) & SM::QMASK]) = (unsigned short)slot;
      }
      __syncwarp();
    };

    for (;;) {
      // 1. a chunk whose cells have all converged → phase C
      int buf = -1;
      if (lane == 0) {
        for (;;) {
          const unsigned h = ld_vol(&sm.headc), t = ld_vol(&sm.tailc);
          if ((int)(t - h) <= 0) break;
          if (atomicCAS(&sm.headc, h, h + 1u) == h) {
            volatile unsigned short* e = &sm.qc[h & SM::CMASK];
            unsigned short v;
            while ((v = *e) == 0xFFFF) {}
            *e = 0xFFFF;
            buf = v;
            break;
          }
        }
      }
      buf = __shfl_sync(FULL, buf, 0);
      if (buf >= 0) { __threadfence_block(); phase_c(buf); continue; }
      // 2. a free buffer and an unclaimed chunk → phase A
      int k = -1;
      if (lane == 0 && ld_vol(&sm.next_chunk) < nchunks_cta) {
        for (;;) {
          const unsigned long long m = ld_vol(&sm.free_mask);
          if (m == 0ull) break;
          const int b = __ffsll((long long)m) - 1;
          if (atomicCAS(&sm.free_mask, m, m & ~(1ull << b)) == m) {
            k = atomicAdd(&sm.next_chunk, 1);
            if (k >= nchunks_cta) { k = -1; atomicOr(&sm.free_mask, 1ull << b); }
            else buf = b;
            break;
          }
        }
      }
      k = __shfl_sync(FULL, k, 0); buf = __shfl_sync(FULL, buf, 0);
      if (k >= 0) {
        phase_a(buf, k);
        if (lane == 0) { __threadfence_block(); atomicAdd(&sm.a_done, 1); }
        continue;
      }
      if (ld_vol(&sm.c_done) >= nchunks_cta) break;
      __nanosleep(200);
    }
  } else {
    // ================================================================ solver warps: lane refill from ONE class queue
    constexpr int BRENT_FROM = (sizeof(FT) == 8) ? 24 : 6;
    constexpr int NSOLVE = TT::NW - TT::NSERVICE;
    int cls = ((warp - TT::NSERVICE) * 4 < NSOLVE * 3) ? 0 : 1;      // initial guess: three quarters of the warps on the unstable class
    int slot = -1, it = 0;
    LeanCell<FT> lc{};
    FT nu = F.mr.visc.nu, us = 0, ts = 0, qs = 0;
    if (!VARNU) { lc.bnu = K.bnu; lc.inv_nu = K.inv_nu; }
    auto brent = [&](bool go) -> bool {
      if (it == BRENT_FROM) {
        sm.U2[slot] = us; sm.dth[slot] = ts; sm.dq[slot] = qs;
        SI::set(sm.c1[slot], it | (1 << 8));
        return go;
      }
      const int packed = SI::get(sm.c1[slot]);
      const int snap_it = packed & 0xff, window = (packed >> 8) & 0xff, stop_at = (packed >> 16) - 1;
      if (stop_at >= 0) {
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
    for (;;) {
      const unsigned active = __ballot_sync(FULL, slot >= 0);
      if (active != FULL) {
        const unsigned idle = ~active;
        int base = 0, take = 0;
        if (lane == 0) {
          for (;;) {
            const unsigned h = ld_vol(&sm.head[cls]), t = ld_vol(&sm.tail[cls]);
            const int avail = (int)(t - h), want = __popc(idle);
            take = avail < want ? avail : want;
            if (take <= 0) { take = 0; break; }
            if (atomicCAS(&sm.head[cls], h, h + (unsigned)take) == h) { base = (int)h; break; }
          }
        }
        take = __shfl_sync(FULL, take, 0); base = __shfl_sync(FULL, base, 0);
        if (take > 0) {
          const int rank = __popc(idle & ((1u << lane) - 1u));
          if (slot < 0 && rank < take) {
            volatile unsigned short* e = &sm.q[cls][(unsigned)(base + rank) & SM::QMASK];
            unsigned short v;
            while ((v = *e) == 0xFFFF) {}                   // reserved by its producer, written a few instructions later
            *e = 0xFFFF;
            __threadfence_block();
            slot = v;
            lc.U2 = sm.U2[slot]; lc.dth = sm.dth[slot]; lc.dq = sm.dq[slot]; lc.cb1 = sm.c1[slot]; lc.cb2 = sm.c2[slot];
            { const FT v2 = fm::fma_(F.ugmin, F.ugmin, lc.U2); lc.Ustab = (v2 > FT(0)) ? fm::sqrt(v2) : FT(0); }
            if (VARNU) { nu = sm.nu[slot]; lc.inv_nu = sm.inu[slot]; lc.bnu = F.mr.beta_s * nu; }
            us = sm.us1[slot];
            if (us >= FT(0)) { const FT chi = sm.chi1[slot]; ts = chi * lc.dth; qs = chi * lc.dq; it = 1; }
            else { us = ts = qs = F.init; it = 0; }
          }
        } else if (active == 0u) {
          // the whole warp is idle and its class queue is empty: serve the other class if it has cells, leave when the
          // producers are finished and both queues are empty, else wait
          int go_on = 1;
          if (lane == 0) {
            const int other = (int)(ld_vol(&sm.tail[cls ^ 1]) - ld_vol(&sm.head[cls ^ 1]));
            if (other > 0) go_on = 2;
            else if (ld_vol(&sm.a_done) >= nchunks_cta &&
                     (int)(ld_vol(&sm.tail[0]) - ld_vol(&sm.head[0])) <= 0 && (int)(ld_vol(&sm.tail[1]) - ld_vol(&sm.head[1])) <= 0) go_on = 0;
          }
          go_on = __shfl_sync(FULL, go_on, 0);
          if (go_on == 0) break;
          if (go_on == 2) cls ^= 1; else __nanosleep(200);
          continue;
        }
      }
      if (slot >= 0) {
        const FT u0 = us, t0 = ts, q0 = qs;
        if (__builtin_expect(!iterate_lean<FT, SPEC>(P, F, K, tb, lc, us, ts, qs), 0)) {
          const D3<FT> r = lean_cold_pass<FT, SPEC>(&P, lc.U2, lc.dth, lc.dq, lc.cb1, lc.cb2, nu, u0, t0, q0);
          us = r.u; ts = r.t; qs = r.q;
        }
        ++it;
        bool go = keep_going<FT>(F, it, us, ts, qs, u0, t0, q0);
        if (__builtin_expect(it >= BRENT_FROM && !fixed && F.maxit < 250, 0)) go = brent(go);
        if (!go) {
          sm.U2[slot] = us; sm.dth[slot] = ts; sm.dq[slot] = qs; SI::set(sm.c1[slot], it);
          __threadfence_block();                           // the result before the completion count
          const int buf = slot >> 5;
          if (atomicAdd(&sm.n_done[buf], 1) + 1 == ld_vol(&sm.n_queued[buf])) push_ready(buf);
          slot = -1;
        }
      }
    }
  }
}

}  // namespace coflux

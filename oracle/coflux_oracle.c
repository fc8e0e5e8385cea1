/* coflux_oracle.c — CPU oracle for the surface-flux hot path.
 *
 * ==========================================================================================
 *  TEST INFRASTRUCTURE.  PARITY UNPINNED.
 *
 *  This is a CPU restatement of the algorithm ClimaOcean's `update_state!(::OceanSeaIceModel)`
 *  executes per coupling step.  The reference snapshot (/root/reference, ClimaOcean v0.10.0)
 *  does NOT contain that arithmetic: it lives in the un-vendored dependency NumericalEarth.jl,
 *  "pinned" only as rev = "main" (Project.toml:21,31-32,48), and neither Julia nor that source
 *  is available in this environment.  The reference ships no golden vector, known-answer test
 *  or fixture for any flux value (test/ *.jl files).  This file therefore follows
 *  SURVEY.md Appendix A — the published formulas the reference cites (Edson et al. 2013, Large &
 *  Yeager 2009, Grachev et al. 2007, Paulson 1970, Holland & Jenkins 1999; docs/climaocean.bib)
 *  reconciled with the in-tree parameter surface (src/OMIPConfigurations/omip_simulation.jl:
 *  40-113,123-164; atmosphere.jl:13-49; omip_diagnostics.jl:77-89).  It is anchored on analytic
 *  known answers and mpmath cross-checks in tests/, NOT on outputs of the reference.
 *
 *  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 *  load this library — as the checker or the timed CPU baseline, never as the product path.
 * ==========================================================================================
 *
 * Build: make -C oracle   (gcc -O2 -fopenmp -ffp-contract=off; no fast-math)
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stddef.h>
#include "../include/coflux.h"

/* A8 (time part): bracket `time` in a series window.  LINEAR extrapolates with the end
 * intervals, CLAMP clamps the weight to [0,1], CYCLICAL wraps with the given period. */
int oracle_time_indices(const double* times, int32_t Nt, int32_t mode, double period, double time,
                        int32_t* n1, int32_t* n2, double* frac) {
  if (!times || Nt < 1) return -1;
  if (Nt == 1) { *n1 = *n2 = 0; *frac = 0.0; return 0; }
  double t = time;
  if (mode == COFLUX_TIME_CYCLICAL) {
    double dtl = times[Nt - 1] - times[Nt - 2];
    double T = (period > 0.0) ? period : (times[Nt - 1] - times[0] + dtl);
    double rel = fmod(t - times[0], T);
    if (rel < 0.0) rel += T;
    t = times[0] + rel;
    if (t >= times[Nt - 1]) { /* wrap interval between the last and the first level */
      *n1 = Nt - 1; *n2 = 0;
      *frac = (t - times[Nt - 1]) / (times[0] + T - times[Nt - 1]);
      return 0;
    }
  }
  int n = 0;
  while (n < Nt - 2 && t >= times[n + 1]) ++n;
  double f = (t - times[n]) / (times[n + 1] - times[n]);
  if (mode == COFLUX_TIME_CLAMP) { if (f < 0.0) f = 0.0; if (f > 1.0) f = 1.0; }
  *n1 = n; *n2 = n + 1; *frac = f;
  return 0;
}

/* ---- Float64 instantiation ---- */
#define FT double
#define SUF(name) name##_f64
#define LOG log
#define EXP exp
#define SQRT sqrt
#define CBRT cbrt
#define ATAN atan
#define POW pow
#define FABS fabs
#define FLOOR floor
#define FMIN fmin
#define FMAX fmax
#define TRUNC trunc
#include "oracle_impl.h"
#undef FT
#undef SUF
#undef LOG
#undef EXP
#undef SQRT
#undef CBRT
#undef ATAN
#undef POW
#undef FABS
#undef FLOOR
#undef FMIN
#undef FMAX
#undef TRUNC

/* ---- Float32 instantiation ---- */
#define FT float
#define SUF(name) name##_f32
#define LOG logf
#define EXP expf
#define SQRT sqrtf
#define CBRT cbrtf
#define ATAN atanf
#define POW powf
#define FABS fabsf
#define FLOOR floorf
#define FMIN fminf
#define FMAX fmaxf
#define TRUNC truncf
#include "oracle_impl.h"

int oracle_abi_version(void) { return COFLUX_ABI_VERSION; }

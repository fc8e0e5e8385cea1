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

/* ------------------------------------------------------------------------------------------------
 * The reference's parameter defaults and the three flux configurations of build_coupled_model
 * (/root/reference/src/OMIPConfigurations/omip_simulation.jl:40-113, 123-164; atmosphere.jl:42-43), stated on the oracle
 * side so that the CPU arm of bench.py (`--impl reference`) never has to load the product library.
 * tests/test_abi_cpu.py holds these byte for byte to coflux_default_config / coflux_apply_flux_configuration.
 * ---------------------------------------------------------------------------------------------- */
#include <string.h>
static void visc_const(coflux_air_viscosity* v, double nu) {
  memset(v, 0, sizeof(*v));
  v->kind = COFLUX_VISCOSITY_CONSTANT; v->nu = nu;
  v->c0 = 1.326e-5; v->c1 = v->c0 * 6.542e-3; v->c2 = v->c0 * 8.301e-6; v->c3 = -v->c0 * 4.84e-9;   /* COARE polynomial */
}
static void similarity_defaults(coflux_flux_params* f, int stability) {
  memset(f, 0, sizeof(*f));
  f->formulation = COFLUX_FLUXES_SIMILARITY_THEORY; f->stability_functions = stability;
  f->similarity_form = COFLUX_PROFILE_LOGARITHMIC; f->velocity_formulation = COFLUX_VELOCITY_RELATIVE;
  f->stop_kind = COFLUX_STOP_CONVERGENCE; f->max_iterations = 100; f->interface_temperature = COFLUX_TEMPERATURE_BULK;
  f->tolerance = 1e-8; f->von_karman_constant = 0.4; f->turbulent_prandtl_number = 1.0; f->gustiness_parameter = 1.0;
  f->minimum_gustiness = 0.0; f->initial_scale = 1e-4; f->ly_minimum_wind = 0.5; f->skin_max_delta_T = 5.0;
  coflux_momentum_roughness* m = &f->momentum_roughness;
  m->kind = COFLUX_ROUGHNESS_CHARNOCK; m->wave_formulation = COFLUX_WAVES_CONSTANT; m->fixed_length = 1e-4;
  m->gravity_wave_parameter = 0.02;                       /* omip_simulation.jl:263 */
  m->wind_a1 = 0.0017; m->wind_a2 = -0.005; m->wind_umax = 19.0; m->wind_alpha_min = 0.0;
  m->smooth_wall_parameter = 0.11; m->maximum_length = 1.0; m->gravitational_acceleration = 9.81;
  visc_const(&m->viscosity, 1.5e-5);
  coflux_scalar_roughness s;
  memset(&s, 0, sizeof(s));
  s.kind = COFLUX_ROUGHNESS_REYNOLDS_SCALING; s.fixed_length = 1e-4; s.reynolds_A = 5.85e-5; s.reynolds_b = 0.72; s.maximum_length = 1.6e-4;
  visc_const(&s.viscosity, 1.5e-5);
  f->temperature_roughness = s; f->water_vapor_roughness = s;
}
static void fixed_roughness(coflux_flux_params* f, double lu, double lt, double lq) {
  f->momentum_roughness.kind = COFLUX_ROUGHNESS_FIXED; f->momentum_roughness.fixed_length = lu;
  f->temperature_roughness.kind = COFLUX_ROUGHNESS_FIXED; f->temperature_roughness.fixed_length = lt;
  f->water_vapor_roughness.kind = COFLUX_ROUGHNESS_FIXED; f->water_vapor_roughness.fixed_length = lq;
}
static void default_sea_ice_fluxes(coflux_flux_params* f) {
  similarity_defaults(f, COFLUX_STABILITY_SHEBA_PAULSON);
  f->interface_temperature = COFLUX_TEMPERATURE_SKIN;
  fixed_roughness(f, 1e-4, 1e-4, 1e-4);
}
int oracle_default_config(coflux_config* cfg, int32_t Nx, int32_t Ny, int32_t Nz, int32_t dtype) {
  if (!cfg) return -1;
  memset(cfg, 0, sizeof(*cfg));
  cfg->abi_version = COFLUX_ABI_VERSION; cfg->dtype = dtype; cfg->device = 0;
  cfg->grid.Nx = Nx; cfg->grid.Ny = Ny; cfg->grid.Nz = Nz; cfg->grid.ring = 1; cfg->grid.periodic_x = 1;
  similarity_defaults(&cfg->atmosphere_ocean, COFLUX_STABILITY_EDSON);
  default_sea_ice_fluxes(&cfg->atmosphere_sea_ice);
  coflux_ice_ocean_params* io = &cfg->ice_ocean;
  io->heat_flux = COFLUX_ICE_OCEAN_ICE_BATH; io->friction_velocity = COFLUX_FRICTION_VELOCITY_CONSTANT;
  io->characteristic_melting_speed = 1e-5; io->liquidus_freshwater_melting_temperature = 0.0; io->liquidus_slope = 0.054;
  io->heat_transfer_coefficient = 0.0095; io->salt_transfer_coefficient = 0.0095 / 35.0; io->constant_friction_velocity = 0.002;
  io->minimum_friction_velocity = 1e-4; io->ice_density = 900.0; io->ice_latent_heat = 334e3; io->ice_ocean_drag_coefficient = 5.5e-3;
  io->ice_conductivity = 2.0; io->ice_consolidation_thickness = 0.05;
  coflux_thermodynamics* t = &cfg->atmosphere.thermodynamics;
  t->gas_constant = 8.3144598; t->dry_air_molar_mass = 0.02897; t->water_molar_mass = 0.018015; t->dry_air_adiabatic_exponent = 2.0 / 7.0;
  t->water_vapor_heat_capacity = 1859; t->liquid_water_heat_capacity = 4181; t->ice_heat_capacity = 2100;
  t->reference_vaporization_enthalpy = 2500800; t->reference_sublimation_enthalpy = 2834400; t->reference_temperature = 273.16;
  t->triple_point_temperature = 273.16; t->triple_point_pressure = 611.657; t->water_freezing_temperature = 273.15;
  t->total_ice_nucleation_temperature = 233;
  cfg->atmosphere.surface_layer_height = 10.0; cfg->atmosphere.boundary_layer_height = 512.0; cfg->atmosphere.gravitational_acceleration = 9.81;
  coflux_ocean_properties* o = &cfg->ocean;
  o->reference_density = 1026.0; o->heat_capacity = 3991.86795711963;      /* visualize/common.jl:17-18 */
  o->freshwater_density = 1000.0; o->minimum_salinity = 1.0;               /* omip_simulation.jl:125 */
  o->temperature_units = COFLUX_TEMPERATURE_CELSIUS; o->salt_water_molar_mass = 18.02;
  const double mm[4] = {35.45, 22.99, 96.06, 24.31}, mf[4] = {0.56, 0.31, 0.08, 0.05};
  for (int k = 0; k < 4; ++k) { o->constituent_molar_mass[k] = mm[k]; o->constituent_mass_fraction[k] = mf[k]; }
  coflux_radiation_properties* r = &cfg->radiation;
  r->stefan_boltzmann_constant = 5.67e-8; r->ocean_albedo = 0.06; r->ocean_emissivity = 1.0;   /* atmosphere.jl:43 */
  r->sea_ice_emissivity = 1.0; r->sea_ice_albedo = 0.7; r->shortwave_penetrates = 1;
  r->sea_ice_albedo_kind = COFLUX_SEA_ICE_ALBEDO_PRESCRIBED;
  coflux_ccsm3_albedo* a = &r->ccsm3;
  a->ice_visible = 0.78; a->ice_near_infrared = 0.36; a->snow_visible = 0.98; a->snow_near_infrared = 0.70; a->thickness_scale = 0.3;
  a->melt_temperature_range = 1.5; a->ice_melt_change = 0.075; a->snow_visible_melt_change = 0.10; a->snow_near_infrared_melt_change = 0.15;
  a->snow_patchiness = 0.02; a->ocean_albedo = 0.06; a->visible_fraction = 0.52; a->melting_temperature = 273.15;
  return 0;
}
int oracle_apply_flux_configuration(coflux_config* cfg, const char* name, int32_t velocity) {
  if (!cfg || !name) return -1;
  const int is_default = !strcmp(name, "default");        /* `:default` returns before velocity_formulation is looked at (:127-133) */
  if (!is_default && velocity != COFLUX_VELOCITY_RELATIVE && velocity != COFLUX_VELOCITY_WIND) return -1;
  coflux_flux_params* ao = &cfg->atmosphere_ocean;
  coflux_flux_params* ai = &cfg->atmosphere_sea_ice;
  if (is_default) {
    similarity_defaults(ao, COFLUX_STABILITY_EDSON);
    default_sea_ice_fluxes(ai);
    cfg->ice_ocean.heat_flux = COFLUX_ICE_OCEAN_ICE_BATH; cfg->ice_ocean.friction_velocity = COFLUX_FRICTION_VELOCITY_CONSTANT;
    return 0;
  }
  if (!strcmp(name, "corrected")) {
    similarity_defaults(ao, COFLUX_STABILITY_EDSON);                                    /* :40-50 */
    ao->similarity_form = COFLUX_PROFILE_COARE_LOGARITHMIC; ao->minimum_gustiness = 0.5;
    ao->momentum_roughness.wave_formulation = COFLUX_WAVES_WIND_DEPENDENT;
    ao->momentum_roughness.viscosity.kind = COFLUX_VISCOSITY_TEMPERATURE_POLY;
    ao->temperature_roughness.viscosity.kind = COFLUX_VISCOSITY_TEMPERATURE_POLY;
    ao->water_vapor_roughness.viscosity.kind = COFLUX_VISCOSITY_TEMPERATURE_POLY;
    similarity_defaults(ai, COFLUX_STABILITY_SHEBA_PAULSON);                            /* :62-69 */
    ai->similarity_form = COFLUX_PROFILE_COARE_LOGARITHMIC; ai->minimum_gustiness = 0.2; ai->interface_temperature = COFLUX_TEMPERATURE_SKIN;
    fixed_roughness(ai, 5e-4, 5e-5, 5e-5);
  } else if (!strcmp(name, "ncar")) {
    similarity_defaults(ao, COFLUX_STABILITY_LARGE_YEAGER);                             /* :86-89 */
    ao->formulation = COFLUX_FLUXES_COEFFICIENT_LARGE_YEAGER; ao->stop_kind = COFLUX_STOP_FIXED_ITERATIONS; ao->max_iterations = 5;
    similarity_defaults(ai, COFLUX_STABILITY_LARGE_YEAGER);                             /* :105-113 */
    ai->similarity_form = COFLUX_PROFILE_COARE_LOGARITHMIC; ai->gustiness_parameter = 0.0; ai->minimum_gustiness = 0.5;
    ai->interface_temperature = COFLUX_TEMPERATURE_SKIN;
    fixed_roughness(ai, 5e-4, 5e-4, 5e-4);
  } else {
    return -1;
  }
  cfg->ice_ocean.heat_flux = COFLUX_ICE_OCEAN_THREE_EQUATION;                           /* :77 */
  cfg->ice_ocean.friction_velocity = COFLUX_FRICTION_VELOCITY_MOMENTUM_BASED;
  ao->velocity_formulation = velocity; ai->velocity_formulation = velocity;
  return 0;
}

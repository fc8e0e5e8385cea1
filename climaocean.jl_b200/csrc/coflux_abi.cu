// coflux_abi.cu — the C ABI (include/coflux.h) over the sm_100a kernels.
//
// Host responsibilities: parameter validation, conversion of the POD parameter mirrors to the
// device parameter block (in FT arithmetic), translation of Oceananigans-style halo-padded array
// descriptors to kernel views, kernel launches on the caller's stream, the HOST-buffer end-to-end
// entry, and the multi-GPU seam.  There is no CPU compute path in this file: every entry point
// either launches CUDA kernels or fails with a status code.
#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <vector>
#include "coflux_solve_stream.cuh"
#include <cstdlib>
#include <unistd.h>

using namespace coflux;

// ---------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
#define CUDA_TRY(expr)                                                                                  \
  do {                                                                                                  \
    cudaError_t e__ = (expr);                                                                           \
    if (e__ != cudaSuccess) return fail(COFLUX_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e__)); \
  } while (0)

#define COFLUX_STR_(x) #x
#define COFLUX_STR(x) COFLUX_STR_(x)
extern "C" const char* coflux_last_error(void) { return g_err; }
extern "C" int coflux_abi_version(void) { return COFLUX_ABI_VERSION; }
extern "C" const char* coflux_build_info(void) {
  return "coflux abi 1; target sm_100a; nvcc " COFLUX_STR(__CUDACC_VER_MAJOR__) "." COFLUX_STR(__CUDACC_VER_MINOR__)
         "; no CPU fallback";
}

extern "C" int coflux_sizeof(const char* name) {
  if (!name) return -1;
#define SZ(n) if (!strcmp(name, #n)) return (int)sizeof(coflux_##n)
  SZ(array); SZ(air_viscosity); SZ(momentum_roughness); SZ(scalar_roughness); SZ(flux_params); SZ(thermodynamics);
  SZ(atmosphere_properties); SZ(ocean_properties); SZ(radiation_properties); SZ(ice_ocean_params); SZ(grid_desc);
  SZ(config); SZ(atmos_series); SZ(exchange_state); SZ(ocean_surface); SZ(interface_fluxes); SZ(sea_ice_state);
  SZ(ocean_columns); SZ(ice_ocean_fluxes); SZ(net_ocean_fluxes); SZ(update_inputs); SZ(update_outputs); SZ(host_step); SZ(salinity_normalization); SZ(closure_forcing);
  SZ(land_series); SZ(ccsm3_albedo); SZ(net_sea_ice_fluxes); SZ(flux_averages);
#undef SZ
  return -1;
}

// ---------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------
struct HostStage {   // device staging planes + pipeline resources for coflux_update_state_host
  static const int MAX_CHUNKS = 16;
  int halo = -1;
  size_t plane_bytes = 0;
  char* in[4] = {nullptr, nullptr, nullptr, nullptr};     // ocean u v T S
  char* xch[8] = {};                                     // exchange state
  char* ao[6] = {};                                      // Qv Qc Fv ρτx ρτy Ts
  char* net[8] = {};                                     // τx τy JT JS Qu Qal Qts J0
  cudaStream_t stream = nullptr, stream2 = nullptr;      // compute (chunks alternate, so that consecutive chunk kernels overlap)
  cudaStream_t copy_in = nullptr, copy_out = nullptr;    // H2D / D2H
  cudaEvent_t ev_in[MAX_CHUNKS] = {}, ev_k[MAX_CHUNKS] = {}, ev_f[MAX_CHUNKS] = {};
};
// Multi-GPU seam (mode B): this context's buffer, IPC-mapped by both neighbours.
//   [0]   uint32 data_flag[2]  — written by the WEST neighbour: step id whose ρτx column (parity) is complete
//   [8]   uint32 ack           — written by the EAST neighbour: last step whose column it has consumed
//   [256] data[2][Ny]          — ρτx of the west neighbour's last interior column, double-buffered by step parity
struct Seam {
  bool attached = false;
  int rank = 0, world = 1;
  char* local = nullptr;        // cudaMalloc'd
  char* east = nullptr;         // peer mapping of the east neighbour's buffer (we store our column + flag there)
  char* west = nullptr;         // peer mapping of the west neighbour's buffer (we store our ack there)
  bool east_is_ipc = false, west_is_ipc = false;
  size_t bytes = 0;
  unsigned int step = 0;
};
static const size_t SEAM_DATA_OFFSET = 256;
// cuStreamWriteValue32 / cuStreamWaitValue32, resolved at run time through cudaGetDriverEntryPoint
typedef int (*cuStreamMemOp32_t)(void* stream, unsigned long long addr, unsigned int value, unsigned int flags);
static cuStreamMemOp32_t g_write32 = nullptr, g_wait32 = nullptr;
static const unsigned int COFLUX_WAIT_GEQ = 0x0;   // CU_STREAM_WAIT_VALUE_GEQ: (int32)(*addr - value) >= 0
struct Profile {
  bool on = false;
  static const int RING = 64;             // events are read back lazily; a ring bounds their number
  cudaEvent_t ev[RING][3] = {};
  int pending = 0;
  double flux_ms = 0, stress_ms = 0;
  long long calls = 0;
};
static const int SALT_BLOCKS = 148 * 8;   // CTAs of the salinity-flux reduction (fixed: the summation order must not depend on the grid size)
struct coflux_ctx {
  bool closure_on = false;               // coflux_attach_closure_forcing: by-products emitted by the stress kernel
  coflux_closure_forcing closure;
  bool avg_on = false;                   // coflux_attach_flux_averages: running averages updated by the fused kernels
  coflux_flux_averages avg;
  int sm_count = 0;
  double* salt_ws = nullptr;             // [2 * SALT_BLOCKS] CTA partials + [2] totals of the salinity-flux reduction
  coflux_config cfg;
  Profile prof;
  int device;
  DevParams<double> P64;
  DevParams<float> P32;
  long long launches = 0;
  HostStage stage;
  Seam seam;
};

static bool finite_all(const double* v, int n) {
  for (int k = 0; k < n; ++k)
    if (!std::isfinite(v[k])) return false;
  return true;
}

// ---------------------------------------------------------------------------------------------
// defaults (omip_simulation.jl:40-113 for the presets)
// ---------------------------------------------------------------------------------------------
static coflux_air_viscosity constant_viscosity(double nu) {
  coflux_air_viscosity v;
  memset(&v, 0, sizeof(v));
  v.kind = COFLUX_VISCOSITY_CONSTANT;
  v.nu = nu;
  v.c0 = 1.326e-5; v.c1 = v.c0 * 6.542e-3; v.c2 = v.c0 * 8.301e-6; v.c3 = -v.c0 * 4.84e-9;
  return v;
}
static coflux_air_viscosity temperature_dependent_viscosity() {
  coflux_air_viscosity v = constant_viscosity(1.5e-5);
  v.kind = COFLUX_VISCOSITY_TEMPERATURE_POLY;
  return v;
}
static coflux_flux_params default_similarity_fluxes(int stability) {
  coflux_flux_params f;
  memset(&f, 0, sizeof(f));
  f.formulation = COFLUX_FLUXES_SIMILARITY_THEORY;
  f.stability_functions = stability;
  f.similarity_form = COFLUX_PROFILE_LOGARITHMIC;
  f.velocity_formulation = COFLUX_VELOCITY_RELATIVE;
  f.stop_kind = COFLUX_STOP_CONVERGENCE;
  f.max_iterations = 100;
  f.interface_temperature = COFLUX_TEMPERATURE_BULK;
  f.tolerance = 1e-8;
  f.von_karman_constant = 0.4;
  f.turbulent_prandtl_number = 1.0;
  f.gustiness_parameter = 1.0;
  f.minimum_gustiness = 0.0;
  f.initial_scale = 1e-4;
  f.ly_minimum_wind = 0.5;
  f.skin_max_delta_T = 5.0;
  coflux_momentum_roughness& m = f.momentum_roughness;
  m.kind = COFLUX_ROUGHNESS_CHARNOCK;
  m.wave_formulation = COFLUX_WAVES_CONSTANT;
  m.fixed_length = 1e-4;
  m.gravity_wave_parameter = 0.02;   // ":default — Edson/COARE with constant Charnock 0.02" (omip_simulation.jl:263)
  m.wind_a1 = 0.0017; m.wind_a2 = -0.005; m.wind_umax = 19.0; m.wind_alpha_min = 0.0;
  m.smooth_wall_parameter = 0.11;
  m.maximum_length = 1.0;
  m.gravitational_acceleration = 9.81;
  m.viscosity = constant_viscosity(1.5e-5);
  coflux_scalar_roughness s;
  memset(&s, 0, sizeof(s));
  s.kind = COFLUX_ROUGHNESS_REYNOLDS_SCALING;
  s.fixed_length = 1e-4;
  s.reynolds_A = 5.85e-5; s.reynolds_b = 0.72;
  s.maximum_length = 1.6e-4;
  s.viscosity = constant_viscosity(1.5e-5);
  f.temperature_roughness = s;
  f.water_vapor_roughness = s;
  return f;
}
static void set_fixed_roughness(coflux_flux_params& f, double lu, double lt, double lq) {
  f.momentum_roughness.kind = COFLUX_ROUGHNESS_FIXED; f.momentum_roughness.fixed_length = lu;
  f.temperature_roughness.kind = COFLUX_ROUGHNESS_FIXED; f.temperature_roughness.fixed_length = lt;
  f.water_vapor_roughness.kind = COFLUX_ROUGHNESS_FIXED; f.water_vapor_roughness.fixed_length = lq;
}

extern "C" int coflux_default_config(coflux_config* cfg, int32_t Nx, int32_t Ny, int32_t Nz, int32_t dtype) {
  if (!cfg) return fail(COFLUX_ERR_INVALID_ARGUMENT, "cfg is NULL");
  memset(cfg, 0, sizeof(*cfg));
  cfg->abi_version = COFLUX_ABI_VERSION;
  cfg->dtype = dtype;
  cfg->device = 0;
  cfg->grid.Nx = Nx; cfg->grid.Ny = Ny; cfg->grid.Nz = Nz; cfg->grid.ring = 1; cfg->grid.periodic_x = 1;
  cfg->atmosphere_ocean = default_similarity_fluxes(COFLUX_STABILITY_EDSON);
  cfg->atmosphere_sea_ice = default_similarity_fluxes(COFLUX_STABILITY_SHEBA_PAULSON);
  cfg->atmosphere_sea_ice.interface_temperature = COFLUX_TEMPERATURE_SKIN;
  set_fixed_roughness(cfg->atmosphere_sea_ice, 1e-4, 1e-4, 1e-4);
  coflux_ice_ocean_params& io = cfg->ice_ocean;
  io.heat_flux = COFLUX_ICE_OCEAN_ICE_BATH;
  io.friction_velocity = COFLUX_FRICTION_VELOCITY_CONSTANT;
  io.characteristic_melting_speed = 1e-5;
  io.liquidus_freshwater_melting_temperature = 0.0;
  io.liquidus_slope = 0.054;
  io.heat_transfer_coefficient = 0.0095;
  io.salt_transfer_coefficient = 0.0095 / 35.0;
  io.constant_friction_velocity = 0.002;
  io.minimum_friction_velocity = 1e-4;
  io.ice_density = 900.0;
  io.ice_latent_heat = 334e3;
  io.ice_ocean_drag_coefficient = 5.5e-3;
  io.ice_conductivity = 2.0;
  io.ice_consolidation_thickness = 0.05;
  coflux_thermodynamics& t = cfg->atmosphere.thermodynamics;
  t.gas_constant = 8.3144598; t.dry_air_molar_mass = 0.02897; t.water_molar_mass = 0.018015;
  t.dry_air_adiabatic_exponent = 2.0 / 7.0;
  t.water_vapor_heat_capacity = 1859; t.liquid_water_heat_capacity = 4181; t.ice_heat_capacity = 2100;
  t.reference_vaporization_enthalpy = 2500800; t.reference_sublimation_enthalpy = 2834400;
  t.reference_temperature = 273.16; t.triple_point_temperature = 273.16; t.triple_point_pressure = 611.657;
  t.water_freezing_temperature = 273.15; t.total_ice_nucleation_temperature = 233;
  cfg->atmosphere.surface_layer_height = 10.0;
  cfg->atmosphere.boundary_layer_height = 512.0;
  cfg->atmosphere.gravitational_acceleration = 9.81;
  coflux_ocean_properties& o = cfg->ocean;
  o.reference_density = 1026.0;             // visualize/common.jl:17
  o.heat_capacity = 3991.86795711963;       // visualize/common.jl:18
  o.freshwater_density = 1000.0;
  o.minimum_salinity = 1.0;                 // omip_simulation.jl:125
  o.temperature_units = COFLUX_TEMPERATURE_CELSIUS;
  o.salt_water_molar_mass = 18.02;
  const double mm[4] = {35.45, 22.99, 96.06, 24.31}, mf[4] = {0.56, 0.31, 0.08, 0.05};
  for (int k = 0; k < 4; ++k) { o.constituent_molar_mass[k] = mm[k]; o.constituent_mass_fraction[k] = mf[k]; }
  coflux_radiation_properties& r = cfg->radiation;
  r.stefan_boltzmann_constant = 5.67e-8;
  r.ocean_albedo = 0.06; r.ocean_emissivity = 1.0;   // atmosphere.jl:43 (OMIP-2 ocean surface)
  r.sea_ice_emissivity = 1.0; r.sea_ice_albedo = 0.7;
  r.shortwave_penetrates = 1;
  r.sea_ice_albedo_kind = COFLUX_SEA_ICE_ALBEDO_PRESCRIBED;
  coflux_ccsm3_albedo& a = r.ccsm3;      // CCSM3 / CICE "ccsm3" constants (Briegleb et al. 2004)
  a.ice_visible = 0.78; a.ice_near_infrared = 0.36; a.snow_visible = 0.98; a.snow_near_infrared = 0.70;
  a.thickness_scale = 0.3; a.melt_temperature_range = 1.5; a.ice_melt_change = 0.075;
  a.snow_visible_melt_change = 0.10; a.snow_near_infrared_melt_change = 0.15; a.snow_patchiness = 0.02;
  a.ocean_albedo = 0.06; a.visible_fraction = 0.52; a.melting_temperature = 273.15;
  return COFLUX_OK;
}

extern "C" int coflux_apply_flux_configuration(coflux_config* cfg, const char* name, int32_t velocity) {
  if (!cfg || !name) return fail(COFLUX_ERR_INVALID_ARGUMENT, "NULL argument");
  // `:default` returns before velocity_formulation is looked at (omip_simulation.jl:127-133): the ComponentInterfaces
  // default (relative velocity) is used and the argument is neither validated nor applied
  const bool is_default = !strcmp(name, "default");
  if (!is_default && velocity != COFLUX_VELOCITY_RELATIVE && velocity != COFLUX_VELOCITY_WIND)
    return fail(COFLUX_ERR_INVALID_ARGUMENT, "Unknown velocity_formulation: %d. Options: relative (0), wind (1)", velocity);
  if (is_default) {
    cfg->atmosphere_ocean = default_similarity_fluxes(COFLUX_STABILITY_EDSON);
    cfg->atmosphere_sea_ice = default_similarity_fluxes(COFLUX_STABILITY_SHEBA_PAULSON);
    cfg->atmosphere_sea_ice.interface_temperature = COFLUX_TEMPERATURE_SKIN;
    set_fixed_roughness(cfg->atmosphere_sea_ice, 1e-4, 1e-4, 1e-4);
    cfg->ice_ocean.heat_flux = COFLUX_ICE_OCEAN_ICE_BATH;
    cfg->ice_ocean.friction_velocity = COFLUX_FRICTION_VELOCITY_CONSTANT;
  } else if (!strcmp(name, "corrected")) {
    // corrected_atmosphere_ocean_fluxes (omip_simulation.jl:40-50)
    coflux_flux_params ao = default_similarity_fluxes(COFLUX_STABILITY_EDSON);
    ao.similarity_form = COFLUX_PROFILE_COARE_LOGARITHMIC;
    ao.minimum_gustiness = 0.5;
    ao.momentum_roughness.wave_formulation = COFLUX_WAVES_WIND_DEPENDENT;
    ao.momentum_roughness.viscosity = temperature_dependent_viscosity();
    ao.temperature_roughness.viscosity = temperature_dependent_viscosity();
    ao.water_vapor_roughness.viscosity = temperature_dependent_viscosity();
    cfg->atmosphere_ocean = ao;
    // corrected_atmosphere_sea_ice_fluxes (omip_simulation.jl:62-69)
    coflux_flux_params ai = default_similarity_fluxes(COFLUX_STABILITY_SHEBA_PAULSON);
    ai.similarity_form = COFLUX_PROFILE_COARE_LOGARITHMIC;
    ai.minimum_gustiness = 0.2;
    ai.interface_temperature = COFLUX_TEMPERATURE_SKIN;
    set_fixed_roughness(ai, 5e-4, 5e-5, 5e-5);
    cfg->atmosphere_sea_ice = ai;
    // corrected_ice_ocean_heat_flux (omip_simulation.jl:77)
    cfg->ice_ocean.heat_flux = COFLUX_ICE_OCEAN_THREE_EQUATION;
    cfg->ice_ocean.friction_velocity = COFLUX_FRICTION_VELOCITY_MOMENTUM_BASED;
  } else if (!strcmp(name, "ncar")) {
    // ncar_atmosphere_ocean_fluxes (omip_simulation.jl:86-89)
    coflux_flux_params ao = default_similarity_fluxes(COFLUX_STABILITY_LARGE_YEAGER);
    ao.formulation = COFLUX_FLUXES_COEFFICIENT_LARGE_YEAGER;
    ao.stop_kind = COFLUX_STOP_FIXED_ITERATIONS;
    ao.max_iterations = 5;
    cfg->atmosphere_ocean = ao;
    // ncar_atmosphere_sea_ice_fluxes (omip_simulation.jl:105-113)
    coflux_flux_params ai = default_similarity_fluxes(COFLUX_STABILITY_LARGE_YEAGER);
    ai.similarity_form = COFLUX_PROFILE_COARE_LOGARITHMIC;
    ai.gustiness_parameter = 0.0;
    ai.minimum_gustiness = 0.5;
    ai.interface_temperature = COFLUX_TEMPERATURE_SKIN;
    set_fixed_roughness(ai, 5e-4, 5e-4, 5e-4);
    cfg->atmosphere_sea_ice = ai;
    cfg->ice_ocean.heat_flux = COFLUX_ICE_OCEAN_THREE_EQUATION;
    cfg->ice_ocean.friction_velocity = COFLUX_FRICTION_VELOCITY_MOMENTUM_BASED;
  } else {
    return fail(COFLUX_ERR_INVALID_ARGUMENT, "Unknown flux_configuration: %s. Options: default, corrected, ncar", name);
  }
  if (!is_default) {
    cfg->atmosphere_ocean.velocity_formulation = velocity;
    cfg->atmosphere_sea_ice.velocity_formulation = velocity;
  }
  return COFLUX_OK;
}

// ---------------------------------------------------------------------------------------------
// validation + device parameter block
// ---------------------------------------------------------------------------------------------
static int validate_viscosity(const coflux_air_viscosity& v, const char* who) {
  if (v.kind != COFLUX_VISCOSITY_CONSTANT && v.kind != COFLUX_VISCOSITY_TEMPERATURE_POLY)
    return fail(COFLUX_ERR_INVALID_ARGUMENT, "%s: unknown air viscosity kind %d", who, v.kind);
  const double vals[5] = {v.nu, v.c0, v.c1, v.c2, v.c3};
  if (!finite_all(vals, 5)) return fail(COFLUX_ERR_INVALID_ARGUMENT, "%s: non-finite air viscosity parameter", who);
  if (v.kind == COFLUX_VISCOSITY_CONSTANT && !(v.nu > 0)) return fail(COFLUX_ERR_INVALID_ARGUMENT, "%s: viscosity must be > 0", who);
  return COFLUX_OK;
}
static int validate_flux(const coflux_flux_params& f, const char* who) {
  if (f.formulation < 0 || f.formulation > COFLUX_FLUXES_COEFFICIENT_LARGE_YEAGER)
    return fail(COFLUX_ERR_INVALID_ARGUMENT, "%s: unknown flux formulation %d", who, f.formulation);
  if (f.stability_functions < 0 || f.stability_functions > COFLUX_STABILITY_NEUTRAL)
    return fail(COFLUX_ERR_INVALID_ARGUMENT, "%s: unknown stability functions %d", who, f.stability_functions);
  if (f.similarity_form < 0 || f.similarity_form > COFLUX_PROFILE_COARE_LOGARITHMIC)
    return fail(COFLUX_ERR_INVALID_ARGUMENT, "%s: unknown similarity form %d", who, f.similarity_form);
  if (f.velocity_formulation < 0 || f.velocity_formulation > COFLUX_VELOCITY_WIND)
    return fail(COFLUX_ERR_INVALID_ARGUMENT, "%s: Unknown velocity_formulation: %d. Options: relative (0), wind (1)", who, f.velocity_formulation);
  if (f.stop_kind < 0 || f.stop_kind > COFLUX_STOP_FIXED_ITERATIONS)
    return fail(COFLUX_ERR_INVALID_ARGUMENT, "%s: unknown solver stop criteria %d", who, f.stop_kind);
  if (f.interface_temperature < 0 || f.interface_temperature > COFLUX_TEMPERATURE_SKIN)
    return fail(COFLUX_ERR_INVALID_ARGUMENT, "%s: unknown interface temperature formulation %d", who, f.interface_temperature);
  if (f.skin_temperature_update != COFLUX_SKIN_CLAMPED_EXPLICIT && f.skin_temperature_update != COFLUX_SKIN_LINEARIZED_LONGWAVE)
    return fail(COFLUX_ERR_INVALID_ARGUMENT, "%s: unknown skin temperature update %d", who, f.skin_temperature_update);
  if (f.max_iterations < 0 || f.max_iterations > 100000)
    return fail(COFLUX_ERR_INVALID_ARGUMENT, "%s: max_iterations out of range (%d)", who, f.max_iterations);
  const double vals[8] = {f.tolerance, f.von_karman_constant, f.turbulent_prandtl_number, f.gustiness_parameter,
                          f.minimum_gustiness, f.initial_scale, f.ly_minimum_wind, f.skin_max_delta_T};
  if (!finite_all(vals, 8)) return fail(COFLUX_ERR_INVALID_ARGUMENT, "%s: non-finite scalar parameter", who);
  if (!(f.von_karman_constant > 0)) return fail(COFLUX_ERR_INVALID_ARGUMENT, "%s: von_karman_constant must be > 0", who);
  if (f.turbulent_prandtl_number != 1.0)
    return fail(COFLUX_ERR_UNSUPPORTED, "%s: turbulent_prandtl_number != 1 is not supported", who);
  if (f.formulation == COFLUX_FLUXES_COEFFICIENT_LARGE_YEAGER && !(f.ly_minimum_wind > 0))
    return fail(COFLUX_ERR_INVALID_ARGUMENT, "%s: ly_minimum_wind must be > 0", who);
  const coflux_momentum_roughness& m = f.momentum_roughness;
  if (m.kind != COFLUX_ROUGHNESS_FIXED && m.kind != COFLUX_ROUGHNESS_CHARNOCK)
    return fail(COFLUX_ERR_INVALID_ARGUMENT, "%s: unknown momentum roughness kind %d", who, m.kind);
  if (m.wave_formulation < 0 || m.wave_formulation > COFLUX_WAVES_WIND_DEPENDENT)
    return fail(COFLUX_ERR_INVALID_ARGUMENT, "%s: unknown wave formulation %d", who, m.wave_formulation);
  const double mv[10] = {m.fixed_length, m.gravity_wave_parameter, m.wind_a1, m.wind_a2, m.wind_umax, m.wind_alpha_min,
                         m.smooth_wall_parameter, m.maximum_length, m.gravitational_acceleration, 0.0};
  if (!finite_all(mv, 10)) return fail(COFLUX_ERR_INVALID_ARGUMENT, "%s: non-finite momentum roughness parameter", who);
  if (m.kind == COFLUX_ROUGHNESS_FIXED && !(m.fixed_length > 0))
    return fail(COFLUX_ERR_INVALID_ARGUMENT, "%s: fixed momentum roughness length must be > 0", who);
  int rc = validate_viscosity(m.viscosity, who);
  if (rc) return rc;
  const coflux_scalar_roughness* ss[2] = {&f.temperature_roughness, &f.water_vapor_roughness};
  for (int k = 0; k < 2; ++k) {
    const coflux_scalar_roughness& s = *ss[k];
    if (s.kind != COFLUX_ROUGHNESS_FIXED && s.kind != COFLUX_ROUGHNESS_REYNOLDS_SCALING)
      return fail(COFLUX_ERR_INVALID_ARGUMENT, "%s: unknown scalar roughness kind %d", who, s.kind);
    const double sv[4] = {s.fixed_length, s.reynolds_A, s.reynolds_b, s.maximum_length};
    if (!finite_all(sv, 4)) return fail(COFLUX_ERR_INVALID_ARGUMENT, "%s: non-finite scalar roughness parameter", who);
    if (s.kind == COFLUX_ROUGHNESS_FIXED && !(s.fixed_length > 0))
      return fail(COFLUX_ERR_INVALID_ARGUMENT, "%s: fixed scalar roughness length must be > 0", who);
    rc = validate_viscosity(s.viscosity, who);
    if (rc) return rc;
  }
  return COFLUX_OK;
}
static int validate_config(const coflux_config* c) {
  if (c->abi_version != COFLUX_ABI_VERSION)
    return fail(COFLUX_ERR_INVALID_ARGUMENT, "abi_version %d != library %d", c->abi_version, COFLUX_ABI_VERSION);
  if (c->dtype != COFLUX_F32 && c->dtype != COFLUX_F64) return fail(COFLUX_ERR_INVALID_ARGUMENT, "dtype must be 32 or 64 (got %d)", c->dtype);
  if (c->grid.Nx < 1 || c->grid.Ny < 1 || c->grid.Nz < 1) return fail(COFLUX_ERR_INVALID_ARGUMENT, "grid size must be positive");
  if (c->grid.ring != 0 && c->grid.ring != 1) return fail(COFLUX_ERR_INVALID_ARGUMENT, "grid.ring must be 0 or 1");
  int rc = validate_flux(c->atmosphere_ocean, "atmosphere_ocean");
  if (rc) return rc;
  rc = validate_flux(c->atmosphere_sea_ice, "atmosphere_sea_ice");
  if (rc) return rc;
  if (c->atmosphere_ocean.interface_temperature != COFLUX_TEMPERATURE_BULK)
    return fail(COFLUX_ERR_UNSUPPORTED, "atmosphere_ocean: only the bulk interface temperature is supported over the ocean");
  const coflux_ice_ocean_params& io = c->ice_ocean;
  if (io.heat_flux < 0 || io.heat_flux > COFLUX_ICE_OCEAN_THREE_EQUATION) return fail(COFLUX_ERR_INVALID_ARGUMENT, "unknown sea_ice_ocean_heat_flux %d", io.heat_flux);
  if (io.friction_velocity < 0 || io.friction_velocity > COFLUX_FRICTION_VELOCITY_MOMENTUM_BASED)
    return fail(COFLUX_ERR_INVALID_ARGUMENT, "unknown friction velocity formulation %d", io.friction_velocity);
  const double iv[12] = {io.characteristic_melting_speed, io.liquidus_freshwater_melting_temperature, io.liquidus_slope,
                         io.heat_transfer_coefficient, io.salt_transfer_coefficient, io.constant_friction_velocity,
                         io.minimum_friction_velocity, io.ice_density, io.ice_latent_heat, io.ice_ocean_drag_coefficient,
                         io.ice_conductivity, io.ice_consolidation_thickness};
  if (!finite_all(iv, 12)) return fail(COFLUX_ERR_INVALID_ARGUMENT, "non-finite sea-ice–ocean parameter");
  const coflux_thermodynamics& t = c->atmosphere.thermodynamics;
  const double tv[14] = {t.gas_constant, t.dry_air_molar_mass, t.water_molar_mass, t.dry_air_adiabatic_exponent,
                         t.water_vapor_heat_capacity, t.liquid_water_heat_capacity, t.ice_heat_capacity,
                         t.reference_vaporization_enthalpy, t.reference_sublimation_enthalpy, t.reference_temperature,
                         t.triple_point_temperature, t.triple_point_pressure, t.water_freezing_temperature,
                         t.total_ice_nucleation_temperature};
  if (!finite_all(tv, 14)) return fail(COFLUX_ERR_INVALID_ARGUMENT, "non-finite thermodynamics parameter");
  for (int k = 0; k < 14; ++k)
    if (!(tv[k] > 0)) return fail(COFLUX_ERR_INVALID_ARGUMENT, "thermodynamics parameters must be > 0");
  const double av[3] = {c->atmosphere.surface_layer_height, c->atmosphere.boundary_layer_height, c->atmosphere.gravitational_acceleration};
  if (!finite_all(av, 3) || !(av[0] > 0) || !(av[1] > 0) || !(av[2] > 0))
    return fail(COFLUX_ERR_INVALID_ARGUMENT, "atmosphere heights / gravity must be finite and > 0");
  const coflux_ocean_properties& o = c->ocean;
  const double ov[5] = {o.reference_density, o.heat_capacity, o.freshwater_density, o.minimum_salinity, o.salt_water_molar_mass};
  if (!finite_all(ov, 5) || !(ov[0] > 0) || !(ov[1] > 0) || !(ov[2] > 0))
    return fail(COFLUX_ERR_INVALID_ARGUMENT, "ocean properties must be finite and > 0");
  if (!finite_all(o.constituent_molar_mass, 4) || !finite_all(o.constituent_mass_fraction, 4))
    return fail(COFLUX_ERR_INVALID_ARGUMENT, "non-finite salinity constituent");
  if (o.temperature_units != COFLUX_TEMPERATURE_CELSIUS && o.temperature_units != COFLUX_TEMPERATURE_KELVIN)
    return fail(COFLUX_ERR_INVALID_ARGUMENT, "unknown ocean temperature units %d", o.temperature_units);
  const coflux_radiation_properties& r = c->radiation;
  const double rv[5] = {r.stefan_boltzmann_constant, r.ocean_albedo, r.ocean_emissivity, r.sea_ice_emissivity, r.sea_ice_albedo};
  if (!finite_all(rv, 5)) return fail(COFLUX_ERR_INVALID_ARGUMENT, "non-finite radiation property");
  if (r.sea_ice_albedo_kind != COFLUX_SEA_ICE_ALBEDO_PRESCRIBED && r.sea_ice_albedo_kind != COFLUX_SEA_ICE_ALBEDO_CCSM3)
    return fail(COFLUX_ERR_INVALID_ARGUMENT, "unknown sea_ice_albedo_kind %d", r.sea_ice_albedo_kind);
  if (r.sea_ice_albedo_kind == COFLUX_SEA_ICE_ALBEDO_CCSM3) {
    const coflux_ccsm3_albedo& a = r.ccsm3;
    const double cv[13] = {a.ice_visible, a.ice_near_infrared, a.snow_visible, a.snow_near_infrared, a.thickness_scale,
                           a.melt_temperature_range, a.ice_melt_change, a.snow_visible_melt_change, a.snow_near_infrared_melt_change,
                           a.snow_patchiness, a.ocean_albedo, a.visible_fraction, a.melting_temperature};
    if (!finite_all(cv, 13) || !(a.thickness_scale > 0) || !(a.melt_temperature_range > 0) || !(a.snow_patchiness > 0))
      return fail(COFLUX_ERR_INVALID_ARGUMENT, "CCSM3 albedo parameters must be finite (thickness_scale, melt_temperature_range, snow_patchiness > 0)");
  }
  return COFLUX_OK;
}

template <typename FT> static Visc<FT> to_dev(const coflux_air_viscosity& v) {
  Visc<FT> d;
  d.kind = v.kind; d.nu = (FT)v.nu; d.c0 = (FT)v.c0; d.c1 = (FT)v.c1; d.c2 = (FT)v.c2; d.c3 = (FT)v.c3;
  return d;
}
static bool same_viscosity(const coflux_air_viscosity& a, const coflux_air_viscosity& b) {
  return a.kind == b.kind && a.nu == b.nu && a.c0 == b.c0 && a.c1 == b.c1 && a.c2 == b.c2 && a.c3 == b.c3;
}
template <typename FT> static FluxP<FT> to_dev(const coflux_flux_params& f) {
  FluxP<FT> d;
  d.formulation = f.formulation; d.stability = f.stability_functions; d.form = f.similarity_form;
  d.velocity = f.velocity_formulation; d.stop_kind = f.stop_kind; d.maxit = f.max_iterations; d.itemp = f.interface_temperature;
  d.tol = (FT)f.tolerance; d.kappa = (FT)f.von_karman_constant; d.beta = (FT)f.gustiness_parameter;
  d.ugmin = (FT)f.minimum_gustiness; d.init = (FT)f.initial_scale; d.ly_umin = (FT)f.ly_minimum_wind;
  d.skin_max_dT = (FT)f.skin_max_delta_T;
  const coflux_momentum_roughness& m = f.momentum_roughness;
  d.mr.kind = m.kind; d.mr.waves = m.wave_formulation; d.mr.fixed = (FT)m.fixed_length; d.mr.alpha = (FT)m.gravity_wave_parameter;
  d.mr.a1 = (FT)m.wind_a1; d.mr.a2 = (FT)m.wind_a2; d.mr.umax = (FT)m.wind_umax; d.mr.amin = (FT)m.wind_alpha_min;
  d.mr.beta_s = (FT)m.smooth_wall_parameter; d.mr.lmax = (FT)m.maximum_length; d.mr.g = (FT)m.gravitational_acceleration;
  d.mr.visc = to_dev<FT>(m.viscosity);
  const coflux_scalar_roughness& t = f.temperature_roughness;
  const coflux_scalar_roughness& q = f.water_vapor_roughness;
  d.tr.kind = t.kind; d.tr.fixed = (FT)t.fixed_length; d.tr.A = (FT)t.reynolds_A; d.tr.b = (FT)t.reynolds_b; d.tr.lmax = (FT)t.maximum_length;
  d.tr.visc = to_dev<FT>(t.viscosity);
  d.qr.kind = q.kind; d.qr.fixed = (FT)q.fixed_length; d.qr.A = (FT)q.reynolds_A; d.qr.b = (FT)q.reynolds_b; d.qr.lmax = (FT)q.maximum_length;
  d.qr.visc = to_dev<FT>(q.viscosity);
  d.same_scalar = (t.kind == q.kind && t.fixed_length == q.fixed_length && t.reynolds_A == q.reynolds_A &&
                   t.reynolds_b == q.reynolds_b && t.maximum_length == q.maximum_length && same_viscosity(t.viscosity, q.viscosity))
                      ? 1 : 0;
  d.same_visc = (same_viscosity(m.viscosity, t.viscosity) && same_viscosity(m.viscosity, q.viscosity)) ? 1 : 0;
  d.skin_update = f.skin_temperature_update;
  return d;
}
template <typename FT> static DevParams<FT> make_dev_params(const coflux_config& c) {
  DevParams<FT> P;
  const coflux_thermodynamics& t = c.atmosphere.thermodynamics;
  P.th.R_d = (FT)t.gas_constant / (FT)t.dry_air_molar_mass;
  P.th.R_v = (FT)t.gas_constant / (FT)t.water_molar_mass;
  P.th.eps = (FT)t.dry_air_molar_mass / (FT)t.water_molar_mass;
  P.th.cp_d = P.th.R_d / (FT)t.dry_air_adiabatic_exponent;
  P.th.cp_v = (FT)t.water_vapor_heat_capacity; P.th.cp_l = (FT)t.liquid_water_heat_capacity; P.th.cp_i = (FT)t.ice_heat_capacity;
  P.th.LH_v0 = (FT)t.reference_vaporization_enthalpy; P.th.LH_s0 = (FT)t.reference_sublimation_enthalpy;
  P.th.T_0 = (FT)t.reference_temperature; P.th.T_tr = (FT)t.triple_point_temperature; P.th.p_tr = (FT)t.triple_point_pressure;
  P.th.T_fr = (FT)t.water_freezing_temperature; P.th.T_in = (FT)t.total_ice_nucleation_temperature;
  P.th.Rd_over_Rv = P.th.R_d / P.th.R_v;
  {   // as the device forms them (psat_generic; the library is built with -fmad=false, so the product is rounded before the
      // subtraction — volatile keeps the host compiler from fusing it either): Δcp/R_v and (LH_0 − Δcp·T_0)/R_v in FT arithmetic
    const FT dl = P.th.cp_v - P.th.cp_l, di = P.th.cp_v - P.th.cp_i;
    volatile FT pl = dl * P.th.T_0, pi = di * P.th.T_0;
    P.th.a_liq = dl / P.th.R_v; P.th.b_liq = (P.th.LH_v0 - pl) / P.th.R_v;
    P.th.a_ice = di / P.th.R_v; P.th.b_ice = (P.th.LH_s0 - pi) / P.th.R_v;
    P.th.inv_T_tr = (FT)1 / P.th.T_tr;
  }
  P.h = (FT)c.atmosphere.surface_layer_height; P.hbl = (FT)c.atmosphere.boundary_layer_height;
  P.g = (FT)c.atmosphere.gravitational_acceleration;
  P.rho0 = (FT)c.ocean.reference_density; P.c0 = (FT)c.ocean.heat_capacity; P.rhof = (FT)c.ocean.freshwater_density;
  P.rho0inv = (FT)1 / P.rho0; P.rhofinv = (FT)1 / P.rhof;
  P.Smin = (FT)c.ocean.minimum_salinity;
  FT alpha = (FT)0;
  for (int k = 0; k < 4; ++k) alpha += (FT)c.ocean.constituent_mass_fraction[k] / (FT)c.ocean.constituent_molar_mass[k];
  P.wmf_alpha = (FT)c.ocean.salt_water_molar_mass * alpha;
  P.T_offset = (c.ocean.temperature_units == COFLUX_TEMPERATURE_CELSIUS) ? (FT)273.15 : (FT)0;
  P.sigma = (FT)c.radiation.stefan_boltzmann_constant; P.alb_o = (FT)c.radiation.ocean_albedo; P.emis_o = (FT)c.radiation.ocean_emissivity;
  P.emis_i = (FT)c.radiation.sea_ice_emissivity; P.alb_i = (FT)c.radiation.sea_ice_albedo; P.sw_pen = c.radiation.shortwave_penetrates;
  P.ice_albedo_kind = c.radiation.sea_ice_albedo_kind;
  {
    const coflux_ccsm3_albedo& a = c.radiation.ccsm3;
    P.ccsm3.ice_v = (FT)a.ice_visible; P.ccsm3.ice_n = (FT)a.ice_near_infrared; P.ccsm3.snow_v = (FT)a.snow_visible;
    P.ccsm3.snow_n = (FT)a.snow_near_infrared; P.ccsm3.hmax = (FT)a.thickness_scale; P.ccsm3.dT = (FT)a.melt_temperature_range;
    P.ccsm3.d_ice = (FT)a.ice_melt_change; P.ccsm3.d_snow_v = (FT)a.snow_visible_melt_change;
    P.ccsm3.d_snow_n = (FT)a.snow_near_infrared_melt_change; P.ccsm3.hpatch = (FT)a.snow_patchiness;
    P.ccsm3.alb_o = (FT)a.ocean_albedo; P.ccsm3.fvis = (FT)a.visible_fraction; P.ccsm3.Tmelt = (FT)a.melting_temperature;
  }
  P.ao = to_dev<FT>(c.atmosphere_ocean);
  P.ai = to_dev<FT>(c.atmosphere_sea_ice);
  {
    const FluxP<FT>& F = P.ao;
    FastConsts<FT>& K = P.K;
    K.lnh = std::log(P.h);
    K.edson = (F.stability == COFLUX_STABILITY_EDSON) ? 1 : 0;
    K.gust_skip = ((F.beta >= FT(0)) && (F.ugmin >= FT(0))) ? 1 : 0;
    K.fast_q = ((F.qr.kind == COFLUX_ROUGHNESS_REYNOLDS_SCALING) && (F.qr.b > FT(0)) && (F.qr.A > FT(0)) && (F.qr.lmax > FT(0))) ? 1 : 0;
    K.fast_t = ((F.tr.kind == COFLUX_ROUGHNESS_REYNOLDS_SCALING) && (F.tr.b > FT(0)) && (F.tr.A > FT(0)) && (F.tr.lmax > FT(0))) ? 1 : 0;
    K.lnhA_q = K.fast_q ? std::log(P.h / F.qr.A) : FT(0);
    K.lnhl_q = std::log(P.h / ((F.qr.kind == COFLUX_ROUGHNESS_FIXED) ? F.qr.fixed : F.qr.lmax));
    K.lrclip_q = K.fast_q ? std::log(F.qr.A / F.qr.lmax) / F.qr.b : FT(0);
    K.lnhA_t = K.fast_t ? std::log(P.h / F.tr.A) : FT(0);
    K.lnhl_t = std::log(P.h / ((F.tr.kind == COFLUX_ROUGHNESS_FIXED) ? F.tr.fixed : F.tr.lmax));
    K.lrclip_t = K.fast_t ? std::log(F.tr.A / F.tr.lmax) / F.tr.b : FT(0);
    K.alpha_g = F.mr.alpha / F.mr.g; K.inv_g = FT(1) / F.mr.g;
    K.bnu = F.mr.beta_s * F.mr.visc.nu; K.inv_nu = FT(1) / F.mr.visc.nu;
  }
  const coflux_ice_ocean_params& io = c.ice_ocean;
  P.io.heat_flux = io.heat_flux; P.io.friction = io.friction_velocity; P.io.um_star = (FT)io.characteristic_melting_speed;
  P.io.T0 = (FT)io.liquidus_freshwater_melting_temperature; P.io.slope = (FT)io.liquidus_slope;
  P.io.alpha_h = (FT)io.heat_transfer_coefficient; P.io.alpha_s = (FT)io.salt_transfer_coefficient;
  P.io.ustar_const = (FT)io.constant_friction_velocity; P.io.ustar_min = (FT)io.minimum_friction_velocity;
  P.io.rho_i = (FT)io.ice_density; P.io.L_f = (FT)io.ice_latent_heat; P.io.Cd = (FT)io.ice_ocean_drag_coefficient;
  P.io.k_ice = (FT)io.ice_conductivity; P.io.h_c = (FT)io.ice_consolidation_thickness;
  return P;
}
template <typename FT> static const DevParams<FT>& dev_params(const coflux_ctx* c);
template <> const DevParams<double>& dev_params<double>(const coflux_ctx* c) { return c->P64; }
template <> const DevParams<float>& dev_params<float>(const coflux_ctx* c) { return c->P32; }

extern "C" int coflux_create(coflux_ctx** out, const coflux_config* cfg) {
  if (!out || !cfg) return fail(COFLUX_ERR_INVALID_ARGUMENT, "NULL argument");
  *out = nullptr;
  int rc = validate_config(cfg);
  if (rc) return rc;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(COFLUX_ERR_NO_DEVICE, "no CUDA device available (%s); coflux has no CPU fallback",
                e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  if (cfg->device < 0 || cfg->device >= ndev) return fail(COFLUX_ERR_INVALID_ARGUMENT, "device %d out of range [0,%d)", cfg->device, ndev);
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, cfg->device));
  if (prop.major < 10)
    return fail(COFLUX_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", cfg->device, prop.major, prop.minor);
  coflux_ctx* c = new (std::nothrow) coflux_ctx();
  if (!c) return fail(COFLUX_ERR_ALLOC, "out of host memory");
  c->cfg = *cfg;
  c->device = cfg->device;
  c->P64 = make_dev_params<double>(*cfg);
  c->P32 = make_dev_params<float>(*cfg);
  c->sm_count = prop.multiProcessorCount;
  *out = c;
  return COFLUX_OK;
}

static void free_stage(HostStage& s) {
  for (char*& p : s.in) { if (p) cudaFree(p); p = nullptr; }
  for (char*& p : s.xch) { if (p) cudaFree(p); p = nullptr; }
  for (char*& p : s.ao) { if (p) cudaFree(p); p = nullptr; }
  for (char*& p : s.net) { if (p) cudaFree(p); p = nullptr; }
  if (s.stream) cudaStreamDestroy(s.stream);
  if (s.stream2) cudaStreamDestroy(s.stream2);
  if (s.copy_in) cudaStreamDestroy(s.copy_in);
  if (s.copy_out) cudaStreamDestroy(s.copy_out);
  for (cudaEvent_t& ev : s.ev_in) { if (ev) cudaEventDestroy(ev); ev = nullptr; }
  for (cudaEvent_t& ev : s.ev_k) { if (ev) cudaEventDestroy(ev); ev = nullptr; }
  for (cudaEvent_t& ev : s.ev_f) { if (ev) cudaEventDestroy(ev); ev = nullptr; }
  s = HostStage();
}

extern "C" int coflux_seam_detach(coflux_ctx* c);
extern "C" int coflux_destroy(coflux_ctx* c) {
  if (!c) return COFLUX_OK;
  cudaSetDevice(c->device);
  coflux_seam_detach(c);
  if (c->seam.local) cudaFree(c->seam.local);
  free_stage(c->stage);
  if (c->salt_ws) cudaFree(c->salt_ws);
  for (auto& row : c->prof.ev)
    for (cudaEvent_t& e : row) if (e) cudaEventDestroy(e);
  delete c;
  return COFLUX_OK;
}

static int profile_drain(coflux_ctx* c) {
  Profile& p = c->prof;
  for (int k = 0; k < p.pending; ++k) {
    CUDA_TRY(cudaEventSynchronize(p.ev[k][2]));
    float a = 0, b = 0;
    CUDA_TRY(cudaEventElapsedTime(&a, p.ev[k][0], p.ev[k][1]));
    CUDA_TRY(cudaEventElapsedTime(&b, p.ev[k][1], p.ev[k][2]));
    p.flux_ms += a; p.stress_ms += b; p.calls += 1;
  }
  p.pending = 0;
  return COFLUX_OK;
}
extern "C" int coflux_profile_enable(coflux_ctx* c, int32_t enable) {
  if (!c) return fail(COFLUX_ERR_INVALID_ARGUMENT, "NULL context");
  CUDA_TRY(cudaSetDevice(c->device));
  Profile& p = c->prof;
  if (enable && !p.ev[0][0])
    for (auto& row : p.ev)
      for (cudaEvent_t& e : row) CUDA_TRY(cudaEventCreate(&e));
  if (!enable) { int rc = profile_drain(c); if (rc) return rc; }
  p.on = enable != 0;
  return COFLUX_OK;
}
extern "C" int coflux_profile_read(coflux_ctx* c, double* flux_ms, double* stress_ms, int64_t* calls) {
  if (!c) return fail(COFLUX_ERR_INVALID_ARGUMENT, "NULL context");
  CUDA_TRY(cudaSetDevice(c->device));
  int rc = profile_drain(c);
  if (rc) return rc;
  Profile& p = c->prof;
  if (flux_ms) *flux_ms = p.flux_ms;
  if (stress_ms) *stress_ms = p.stress_ms;
  if (calls) *calls = p.calls;
  p.flux_ms = p.stress_ms = 0; p.calls = 0;
  return COFLUX_OK;
}

extern "C" int coflux_launch_count(coflux_ctx* c, int64_t* n) {
  if (!c || !n) return fail(COFLUX_ERR_INVALID_ARGUMENT, "NULL argument");
  *n = c->launches;
  return COFLUX_OK;
}

// ---------------------------------------------------------------------------------------------
// time indexing (A8)
// ---------------------------------------------------------------------------------------------
extern "C" int coflux_time_indices(const double* times, int32_t Nt, int32_t mode, double period, double time, int32_t* n1,
                                   int32_t* n2, double* frac) {
  if (!times || !n1 || !n2 || !frac) return fail(COFLUX_ERR_INVALID_ARGUMENT, "NULL argument");
  if (Nt < 1) return fail(COFLUX_ERR_INVALID_ARGUMENT, "series has no time levels");
  if (mode < COFLUX_TIME_LINEAR || mode > COFLUX_TIME_CLAMP) return fail(COFLUX_ERR_INVALID_ARGUMENT, "unknown time indexing %d", mode);
  if (!std::isfinite(time)) return fail(COFLUX_ERR_INVALID_ARGUMENT, "non-finite time");
  if (Nt == 1) { *n1 = *n2 = 0; *frac = 0.0; return COFLUX_OK; }
  for (int k = 1; k < Nt; ++k)
    if (!(times[k] > times[k - 1])) return fail(COFLUX_ERR_INVALID_ARGUMENT, "series times must be strictly increasing");
  double t = time;
  if (mode == COFLUX_TIME_CYCLICAL) {
    const double T = (period > 0.0) ? period : (times[Nt - 1] - times[0] + (times[Nt - 1] - times[Nt - 2]));
    double rel = std::fmod(t - times[0], T);
    if (rel < 0.0) rel += T;
    t = times[0] + rel;
    if (t >= times[Nt - 1]) {
      *n1 = Nt - 1; *n2 = 0;
      *frac = (t - times[Nt - 1]) / (times[0] + T - times[Nt - 1]);
      return COFLUX_OK;
    }
  }
  int lo = 0, hi = Nt - 2;   // largest n in [0, Nt-2] with times[n] <= t (binary search)
  while (lo < hi) {
    const int mid = (lo + hi + 1) / 2;
    if (times[mid] <= t) lo = mid; else hi = mid - 1;
  }
  double f = (t - times[lo]) / (times[lo + 1] - times[lo]);
  if (mode == COFLUX_TIME_CLAMP) f = f < 0.0 ? 0.0 : (f > 1.0 ? 1.0 : f);
  *n1 = lo; *n2 = lo + 1; *frac = f;
  return COFLUX_OK;
}

// ---------------------------------------------------------------------------------------------
// descriptor → kernel views
// ---------------------------------------------------------------------------------------------
static inline DArr view2d(const coflux_array& a, int k, size_t esize) {
  DArr d{nullptr, 0, 0};
  if (!a.ptr) return d;
  const int64_t off = (int64_t)a.off_i * a.stride_i + (int64_t)a.off_j * a.stride_j + (int64_t)(k + a.off_k) * a.stride_k;
  d.p = static_cast<char*>(a.ptr) + off * (int64_t)esize;
  d.si = a.stride_i; d.sj = a.stride_j;
  return d;
}
static inline DArr view2d_opt(const coflux_array* a, int k, size_t esize) {
  if (!a) return DArr{nullptr, 0, 0};
  return view2d(*a, k, esize);
}
// ring_capacity > 0: logical level n lives in time slot (ring_start + n) mod ring_capacity (coflux_forcing_window)
static inline DSeries view_series(const coflux_array& a, int n1, int n2, size_t esize, int ring_start = 0, int ring_capacity = 0) {
  DSeries s{nullptr, nullptr, 0, 0};
  if (!a.ptr) return s;
  if (ring_capacity > 0) { n1 = (ring_start + n1) % ring_capacity; n2 = (ring_start + n2) % ring_capacity; }
  const int64_t off = (int64_t)a.off_i * a.stride_i + (int64_t)a.off_j * a.stride_j + (int64_t)a.off_k * a.stride_k;
  s.p1 = static_cast<char*>(a.ptr) + (off + (int64_t)n1 * a.stride_n) * (int64_t)esize;
  s.p2 = static_cast<char*>(a.ptr) + (off + (int64_t)n2 * a.stride_n) * (int64_t)esize;
  s.si = a.stride_i; s.sj = a.stride_j;
  return s;
}
static inline DCol view3d(const coflux_array& a, size_t esize) {
  DCol d{nullptr, 0, 0, 0};
  if (!a.ptr) return d;
  const int64_t off = (int64_t)a.off_i * a.stride_i + (int64_t)a.off_j * a.stride_j + (int64_t)a.off_k * a.stride_k;
  d.p = static_cast<char*>(a.ptr) + off * (int64_t)esize;
  d.si = a.stride_i; d.sj = a.stride_j; d.sk = a.stride_k;
  return d;
}

#define REQUIRE(cond, ...)                                            \
  do {                                                                \
    if (!(cond)) return fail(COFLUX_ERR_INVALID_ARGUMENT, __VA_ARGS__); \
  } while (0)

// A descriptor must cover the cells a kernel touches: `need_i` / `need_j` halo cells on the low side (and, the parents
// being symmetric, on the high side).  Guards against out-of-bounds access through a too-small Julia-side halo.
static int check_halo(const coflux_array& a, int need_i, int need_j, const char* name) {
  if (!a.ptr) return COFLUX_OK;
  if ((a.stride_i != 0 && a.off_i < need_i) || (a.stride_j != 0 && a.off_j < need_j))
    return fail(COFLUX_ERR_INVALID_ARGUMENT, "%s: halo offsets (%d, %d) smaller than the (%d, %d) cells the surface kernels touch", name,
                a.off_i, a.off_j, need_i, need_j);
  return COFLUX_OK;
}
#define HALO(arr, ni, nj, name) do { int rc__ = check_halo(arr, ni, nj, name); if (rc__) return rc__; } while (0)

static int check_launch(coflux_ctx* c, int n) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(COFLUX_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(e));
  c->launches += n;
  return COFLUX_OK;
}

template <typename FT> static void fill_geometry(const coflux_ctx* c, FluxArgs<FT>& a) {
  const coflux_grid_desc& g = c->cfg.grid;
  a.ring = g.ring; a.Nx = g.Nx; a.Ny = g.Ny;
  a.nxr = g.Nx + 2 * g.ring; a.nyr = g.Ny + 2 * g.ring;
  a.cell0 = 0;
  a.ncell = (long long)a.nxr * a.nyr;
  a.P = dev_params<FT>(c);
  a.seam_east = nullptr;
}

template <typename FT>
static int fill_interp(const coflux_ctx* c, const coflux_atmos_series* in, double time, coflux_exchange_state* out, FluxArgs<FT>& a) {
  REQUIRE(in && out, "NULL atmosphere series / exchange state");
  REQUIRE(in->u.ptr && in->v.ptr && in->T.ptr && in->q.ptr && in->p.ptr && in->Qs.ptr && in->Ql.ptr,
          "atmosphere series u, v, T, q, p, Qs, Ql are required");
  REQUIRE(in->fi.ptr && in->fj.ptr, "fractional indices fi, fj are required");
  int32_t n1, n2; double frac;
  int rc = coflux_time_indices(in->times, in->Nt, in->time_indexing, in->cycle_period, time, &n1, &n2, &frac);
  if (rc) return rc;
  const size_t es = sizeof(FT);
  {
    const int ring = c->cfg.grid.ring;
    HALO(in->fi, ring, ring, "fi"); HALO(in->fj, ring, ring, "fj"); HALO(in->cos_theta, ring, ring, "cos_theta"); HALO(in->sin_theta, ring, ring, "sin_theta");
    const coflux_array* xs[8] = {&out->u, &out->v, &out->T, &out->p, &out->q, &out->Qs, &out->Ql, &out->Mp};
    for (const coflux_array* p : xs) HALO(*p, ring, ring, "exchange state");
  }
  REQUIRE(in->ring_capacity >= 0 && in->ring_start >= 0 && (in->ring_capacity == 0 || in->Nt <= in->ring_capacity),
          "atmosphere series: bad ring_start / ring_capacity (Nt must not exceed the ring capacity)");
  const int rs = in->ring_start, rc_ = in->ring_capacity;
  a.su = view_series(in->u, n1, n2, es, rs, rc_); a.sv = view_series(in->v, n1, n2, es, rs, rc_); a.sT = view_series(in->T, n1, n2, es, rs, rc_);
  a.sq = view_series(in->q, n1, n2, es, rs, rc_); a.sp = view_series(in->p, n1, n2, es, rs, rc_); a.sQs = view_series(in->Qs, n1, n2, es, rs, rc_);
  a.sQl = view_series(in->Ql, n1, n2, es, rs, rc_); a.srain = view_series(in->rain, n1, n2, es, rs, rc_); a.ssnow = view_series(in->snow, n1, n2, es, rs, rc_);
  a.fi = view2d(in->fi, 0, es); a.fj = view2d(in->fj, 0, es);
  a.cs = view2d(in->cos_theta, 0, es); a.sn = view2d(in->sin_theta, 0, es);
  a.nfrac = (FT)frac;
  a.xu = view2d(out->u, 0, es); a.xv = view2d(out->v, 0, es); a.xT = view2d(out->T, 0, es); a.xp = view2d(out->p, 0, es);
  a.xq = view2d(out->q, 0, es); a.xQs = view2d(out->Qs, 0, es); a.xQl = view2d(out->Ql, 0, es); a.xMp = view2d(out->Mp, 0, es);
  return COFLUX_OK;
}
template <typename FT> static int fill_exchange_in(const coflux_exchange_state* x, FluxArgs<FT>& a) {
  REQUIRE(x, "NULL exchange state");
  REQUIRE(x->u.ptr && x->v.ptr && x->T.ptr && x->p.ptr && x->q.ptr && x->Qs.ptr && x->Ql.ptr,
          "exchange state u, v, T, p, q, Qs, Ql are required");
  const size_t es = sizeof(FT);
  a.xu = view2d(x->u, 0, es); a.xv = view2d(x->v, 0, es); a.xT = view2d(x->T, 0, es); a.xp = view2d(x->p, 0, es);
  a.xq = view2d(x->q, 0, es); a.xQs = view2d(x->Qs, 0, es); a.xQl = view2d(x->Ql, 0, es); a.xMp = view2d(x->Mp, 0, es);
  return COFLUX_OK;
}
// land freshwater series (rivers + icebergs): own source grid, own time axis
template <typename FT> static int fill_land(const coflux_ctx* c, const coflux_land_series* in, double time, FluxArgs<FT>& a) {
  if (!in) return COFLUX_OK;
  REQUIRE(in->rivers.ptr || in->icebergs.ptr, "land series: rivers and icebergs are both absent");
  REQUIRE(in->fi.ptr && in->fj.ptr, "land series: fractional indices fi, fj are required");
  REQUIRE(in->ring_capacity >= 0 && in->ring_start >= 0 && (in->ring_capacity == 0 || in->Nt <= in->ring_capacity),
          "land series: bad ring_start / ring_capacity");
  int32_t n1, n2; double frac;
  int rc = coflux_time_indices(in->times, in->Nt, in->time_indexing, in->cycle_period, time, &n1, &n2, &frac);
  if (rc) return rc;
  const size_t es = sizeof(FT);
  const int ring = c->cfg.grid.ring;
  HALO(in->fi, ring, ring, "land fi"); HALO(in->fj, ring, ring, "land fj");
  a.sriv = view_series(in->rivers, n1, n2, es, in->ring_start, in->ring_capacity);
  a.sicb = view_series(in->icebergs, n1, n2, es, in->ring_start, in->ring_capacity);
  a.lfi = view2d(in->fi, 0, es); a.lfj = view2d(in->fj, 0, es);
  a.lnfrac = (FT)frac;
  return COFLUX_OK;
}
template <typename FT> static void fill_avg(const coflux_ctx* c, AvgArgs<FT>& g) {
  if (!c->avg_on) { g.on = 0; return; }
  const size_t es = sizeof(FT);
  const coflux_flux_averages& v = c->avg;
  g.on = 1;
  g.JT = view2d(v.JT, 0, es); g.JS = view2d(v.JS, 0, es); g.Qc = view2d(v.Qc, 0, es); g.Qv = view2d(v.Qv, 0, es);
  g.JTao = view2d(v.JT_atmosphere_ocean, 0, es); g.JTio = view2d(v.JT_ice_ocean, 0, es); g.JSio = view2d(v.JS_ice_ocean, 0, es);
  g.T = (FT)v.previous_interval; g.dt = (FT)v.dt;
}

template <typename FT> static int fill_ocean(const coflux_ctx* c, const coflux_ocean_surface* o, FluxArgs<FT>& a) {
  REQUIRE(o, "NULL ocean surface");
  REQUIRE(o->u.ptr && o->v.ptr && o->T.ptr && o->S.ptr, "ocean u, v, T, S are required");
  const int kN = c->cfg.grid.Nz - 1;
  const size_t es = sizeof(FT);
  const int ring = c->cfg.grid.ring;
  HALO(o->u, ring + 1, ring, "ocean u");     // u is averaged to the centre: read at i + 1 of the last ring cell
  HALO(o->v, ring, ring + 1, "ocean v");
  HALO(o->T, ring, ring, "ocean T"); HALO(o->S, ring, ring, "ocean S"); HALO(o->mask, ring, ring, "ocean mask");
  a.ou = view2d(o->u, kN, es); a.ov = view2d(o->v, kN, es); a.oT = view2d(o->T, kN, es); a.oS = view2d(o->S, kN, es);
  a.mask = view2d(o->mask, 0, 1);
  return COFLUX_OK;
}
template <typename FT> static int check_interface_halos(const coflux_ctx* c, const coflux_interface_fluxes* f) {
  const int ring = c->cfg.grid.ring;
  const coflux_array* arr[10] = {&f->latent_heat, &f->sensible_heat, &f->water_vapor, &f->x_momentum, &f->y_momentum,
                                 &f->interface_temperature, &f->friction_velocity, &f->temperature_scale, &f->humidity_scale, &f->iterations};
  for (const coflux_array* p : arr) HALO(*p, ring, ring, "interface flux");
  return COFLUX_OK;
}
template <typename FT> static void fill_interface_out(coflux_interface_fluxes* f, FluxArgs<FT>& a) {
  const size_t es = sizeof(FT);
  a.Qv = view2d(f->latent_heat, 0, es); a.Qc = view2d(f->sensible_heat, 0, es); a.Fv = view2d(f->water_vapor, 0, es);
  a.rtx = view2d(f->x_momentum, 0, es); a.rty = view2d(f->y_momentum, 0, es); a.Tsout = view2d(f->interface_temperature, 0, es);
  a.ust = view2d(f->friction_velocity, 0, es); a.tst = view2d(f->temperature_scale, 0, es); a.qst = view2d(f->humidity_scale, 0, es);
  a.iters = view2d(f->iterations, 0, sizeof(int32_t));
}
template <typename FT> static void zero_args(FluxArgs<FT>& a) { memset(&a, 0, sizeof(a)); }

static inline unsigned grid_for(long long n, int block) { return (unsigned)((n + block - 1) / block); }

// The tile kernel (coflux_solve_tile.cuh) covers similarity theory over the ocean with a bulk interface
// temperature and one viscosity law; everything else runs the one-cell-per-thread kernel.
static bool force_v1() {
  static const bool f = [] { const char* e = std::getenv("COFLUX_FORCE_V1"); return e && e[0] == '1'; }();
  return f;
}
// The stand-alone solves run one cell per thread (flux_kernel).  COFLUX_REFILL=1 selects the lane-refill kernel instead
// (flux_refill_kernel): measured at 1/12° on the sea-ice solve it only pays when a cell's pass is cheap relative to
// popping a new cell — with the compact sea-ice pass, Brent cycle detection and the ψ tables: one cell per thread
// 20.9 ms, refill with batches of 8 / 16 / 32 lanes 28.4 / 25.5 / 21.7 ms (`:default`), 34.2 vs 30.2 ms (`:ncar`).
#ifndef COFLUX_REFILL_F32
#define COFLUX_REFILL_F32 0
#endif
#ifndef COFLUX_REFILL_TILE
#define COFLUX_REFILL_TILE 1024
#endif
static bool refill_v1() {
  static const bool f = [] { const char* e = std::getenv("COFLUX_REFILL"); return e && e[0] == '1'; }();
  return f;
}
// The sea-ice parameter sets that take CellSolver::pass_ice run the tile form of the solve (COFLUX_ICE_TILE=0: one cell
// per thread).  The conditions are those of CellSolver::init's `ice_fast`, plus a skin temperature and a convergence stop.
#ifndef COFLUX_ICE_TILE_CELLS
#define COFLUX_ICE_TILE_CELLS 384   /* 3 cells per lane, 4 CTAs × 51 KB per SM.  Measured at 1/12°, 92 % ice cover, after the ψ-table work
                                       of round 2 (256 / 320 / 384 cells, 4 CTAs; 448 × 3 CTAs): `:default` 11.10 / 10.89 / 11.07 / 11.60 ms,
                                       `:corrected` 8.84 / 8.27 / 8.04 / 8.56 ms.  (Before that work the L1 the larger tiles take away
                                       cost more than they gained: 16.9 ms at 256 against 17.0 at 384.) */
#endif
#ifndef COFLUX_ICE_QUEUE_CELLS
#define COFLUX_ICE_QUEUE_CELLS 8192    /* most cells of one CTA of ice_queue_kernel (2-byte queue entries in shared memory) */
#endif
template <typename FT> static bool ice_tile_eligible(const coflux_ctx* c) {
  static const bool off = [] { const char* e = std::getenv("COFLUX_ICE_TILE"); return e && e[0] == '0'; }();
  const FluxP<FT>& F = dev_params<FT>(c).ai;
  return !off && COFLUX_PSI_TABLES_V1 && F.formulation == COFLUX_FLUXES_SIMILARITY_THEORY &&
         (COFLUX_ICE_COARE || F.form == COFLUX_PROFILE_LOGARITHMIC) &&
         F.mr.kind == COFLUX_ROUGHNESS_FIXED && F.tr.kind == COFLUX_ROUGHNESS_FIXED && F.qr.kind == COFLUX_ROUGHNESS_FIXED &&
         (F.stability == COFLUX_STABILITY_SHEBA_PAULSON || F.stability == COFLUX_STABILITY_LARGE_YEAGER) && F.beta >= FT(0) &&
         F.ugmin >= FT(0) && F.mr.fixed > FT(0) && F.tr.fixed > FT(0) && F.qr.fixed > FT(0) && F.itemp == COFLUX_TEMPERATURE_SKIN &&
         F.stop_kind == COFLUX_STOP_CONVERGENCE && F.maxit >= 1;
}
// (the tile kernel parks ρ_a, c_p,m in the ρτx / ρτy output arrays between its phases: both must be present)
// Uniform layout of the tile kernel (FluxArgs::usj, ssj): every 2-D surface array it touches is contiguous in i with one
// and the same row pitch, and so are the atmosphere series among themselves — the case for Oceananigans parents living on
// one grid.  Anything else (transposed or differently padded fields) takes the one-cell-per-thread kernel, which addresses
// every array through its own strides.
template <typename FT> static bool uniform_layout(FluxArgs<FT>& a) {
  const DArr* s2[] = {&a.xu, &a.xv, &a.xT, &a.xp, &a.xq, &a.xQs, &a.xQl, &a.xMp, &a.ou, &a.ov, &a.oT, &a.oS,
                      &a.Qv, &a.Qc, &a.Fv, &a.rtx, &a.rty, &a.Tsout, &a.ust, &a.tst, &a.qst, &a.conc, &a.Qio, &a.salt_io,
                      &a.JT, &a.JS, &a.Qu, &a.Qal, &a.Qts, &a.J0,
                      &a.avg.JT, &a.avg.JS, &a.avg.Qc, &a.avg.Qv, &a.avg.JTao, &a.avg.JTio, &a.avg.JSio};
  int64_t pitch = 0;
  for (const DArr* d : s2) {
    if (!d->p) continue;
    if (d->si != 1 || d->sj <= 0) return false;
    if (!pitch) pitch = d->sj;
    if (d->sj != pitch) return false;
  }
  if (!pitch || pitch * (int64_t)(a.nyr + 2) >= (int64_t)1 << 31) return false;
  for (const DArr* d : {&a.mask, &a.iters})                  // uint8 / int32 planes: same ELEMENT layout
    if (d->p && (d->si != 1 || d->sj != pitch)) return false;
  a.usj = (int)pitch;
  int64_t fp = 0;                                            // fi, fj, cos θ, sin θ: parents of the ring-extended surface
  for (const DArr* d : {&a.fi, &a.fj, &a.cs, &a.sn}) {
    if (!d->p) continue;
    if (d->si != 1 || d->sj <= 0) return false;
    if (!fp) fp = d->sj;
    if (d->sj != fp) return false;
  }
  if (fp * (int64_t)(a.nyr + 2) >= (int64_t)1 << 31) return false;
  a.fsj = (int)fp;
  const DSeries* ss[] = {&a.su, &a.sv, &a.sT, &a.sq, &a.sp, &a.sQs, &a.sQl, &a.srain, &a.ssnow};
  int64_t sp = 0;
  for (const DSeries* d : ss) {
    if (!d->p1) continue;
    if (d->si != 1 || d->sj <= 0) return false;
    if (!sp) sp = d->sj;
    if (d->sj != sp) return false;
  }
  if (sp >= 46340) return false;                 // gather offsets j·ssj + i stay inside 32 bits for any source grid of < 46 340 rows
  a.ssj = (int)sp;
  return true;
}
template <typename FT> static bool tile_eligible(const coflux_ctx* c, FluxArgs<FT>& a) {
  const FluxP<FT>& F = dev_params<FT>(c).ao;
  return !force_v1() && a.rtx.p && a.rty.p && F.formulation == COFLUX_FLUXES_SIMILARITY_THEORY && F.itemp == COFLUX_TEMPERATURE_BULK && F.same_visc && F.maxit >= 1 &&
         uniform_layout<FT>(a);
}
// compile-time specialisation of the hot loop for the OMIP parameter sets (0 = generic)
template <typename FT> static int tile_spec(const coflux_ctx* c) {
  const DevParams<FT>& P = dev_params<FT>(c);
  const FluxP<FT>& F = P.ao;
  const bool common = P.K.edson && P.K.gust_skip && P.K.fast_q && F.same_scalar && F.mr.kind == COFLUX_ROUGHNESS_CHARNOCK;
  if (common && F.form == COFLUX_PROFILE_LOGARITHMIC && F.mr.waves == COFLUX_WAVES_CONSTANT &&
      F.mr.visc.kind == COFLUX_VISCOSITY_CONSTANT) return 1;   // SPEC 1 reads ν from the parameter block
  if (common && F.form == COFLUX_PROFILE_COARE_LOGARITHMIC && F.mr.waves == COFLUX_WAVES_WIND_DEPENDENT) return 2;
  return 0;
}
template <typename FT, bool INTERP, bool ASSEMBLE, int SPEC> static int launch_tile_spec(const FluxArgs<FT>& a, cudaStream_t st) {
  using TT = TileTraits<FT, SPEC>;
  constexpr int TILE = TT::TILE;
  auto kern = flux_tile_kernel<FT, INTERP, ASSEMBLE, TILE, SPEC>;
  const size_t smem = sizeof(TileSmem<FT, TILE, TT::VARNU, TT::LEAN>);
  static unsigned long long configured = 0;     // per instantiation, one bit per device (function attributes are per device)
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  if (!(configured >> (dev & 63) & 1ull)) {
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // the tiles and the staged ψ rows want all of the SM's shared memory; the kernel keeps nothing hot in L1 any more
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    configured |= 1ull << (dev & 63);
  }
  // Balanced tiling.  CTAs are dispatched as slots free up, so a launch lasts ceil(#CTAs / resident slots) CTA lifetimes
  // (measured: a 1/8 longitude slab of the 1/12° grid, 2.85 waves of 768-cell tiles, took exactly 3 × 168 µs).  Every CTA
  // therefore takes the same number of cells, chosen so that the grid is a whole number of waves: no ragged last wave.
  // Two refinements were measured and dropped (profiles/README.md): a staggered start of the CTAs that share an SM, and
  // tiles quantised to whole cells per lane (full tiles + one lighter last wave): both within ±2 % of this, either way.
  // COFLUX_BALANCE=0: plain tiles of TILE cells.
  static int sm_count[64] = {};
  if (!sm_count[dev & 63]) CUDA_TRY(cudaDeviceGetAttribute(&sm_count[dev & 63], cudaDevAttrMultiProcessorCount, dev));
  FluxArgs<FT> b = a;
  const long long n = a.ncell - a.cell0;
  const long long slots = (long long)sm_count[dev & 63] * TT::MIN_BLOCKS;
  long long grid = grid_for(n, TILE);
  b.tile_cells = 0;
  static const bool balance = [] { const char* e = std::getenv("COFLUX_BALANCE"); return !(e && e[0] == '0'); }();
  if (balance && n > slots * (TILE / 4)) {
    const long long waves = (n + slots * TILE - 1) / (slots * TILE);
    const long long per = (n + waves * slots - 1) / (waves * slots);           // cells per CTA, ≤ TILE
    b.tile_cells = (int)per;
    grid = (n + per - 1) / per;
  }
  kern<<<(unsigned)grid, TT::NT, smem, st>>>(b);
  return COFLUX_OK;
}
// COFLUX_KERNEL=stream selects the persistent warp-specialised kernel (coflux_solve_stream.cuh: one CTA per SM for the
// whole launch) for the OMIP parameter sets.  Measured on B200 (profiles/README.md, round 2): bit-identical results, but
// 6.4 ms against 3.3 ms at 1/12° Float64 — with 24 register-limited warps per SM the 7 service warps cannot feed the
// solver lanes (19 of 32 active) and the two instruction streams thrash the instruction cache (26 % no_inst stalls).
// The tile kernel stays the production path.
static bool use_stream_kernel() {
  static const bool f = [] { const char* e = std::getenv("COFLUX_KERNEL"); return e && !strcmp(e, "stream"); }();
  return f;
}
template <typename FT, bool INTERP, bool ASSEMBLE, int SPEC> static int launch_stream_spec(const FluxArgs<FT>& a, cudaStream_t st) {
  using TT = StreamTraits<FT, SPEC>;
  auto kern = flux_stream_kernel<FT, INTERP, ASSEMBLE, SPEC>;
  const size_t smem = sizeof(StreamSmem<FT, TT::NB, TT::VARNU>);
  static unsigned long long configured = 0;
  static int sm_count[64] = {};
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  if (!(configured >> (dev & 63) & 1ull)) {
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    CUDA_TRY(cudaDeviceGetAttribute(&sm_count[dev & 63], cudaDevAttrMultiProcessorCount, dev));
    configured |= 1ull << (dev & 63);
  }
  const long long nchunks = (a.ncell - a.cell0 + 31) / 32;
  const long long per_cta = 16;                                  // do not spread a small launch thinner than 16 chunks per CTA
  long long grid = (nchunks + per_cta - 1) / per_cta;
  if (grid > sm_count[dev & 63]) grid = sm_count[dev & 63];
  if (grid < 1) grid = 1;
  kern<<<(unsigned)grid, TT::NT, smem, st>>>(a);
  return COFLUX_OK;
}
template <typename FT, bool INTERP, bool ASSEMBLE> static int launch_tile(const coflux_ctx* c, const FluxArgs<FT>& a, cudaStream_t st) {
  const int spec = tile_spec<FT>(c);
  constexpr bool lean = (COFLUX_LEAN != 0) && (sizeof(FT) == 8 || (COFLUX_LEAN_F32 != 0));
  if constexpr (lean) {
    if (use_stream_kernel() && dev_params<FT>(c).ao.maxit < 250) {
      if (spec == 1) return launch_stream_spec<FT, INTERP, ASSEMBLE, 1>(a, st);
      if (spec == 2) return launch_stream_spec<FT, INTERP, ASSEMBLE, 2>(a, st);
    }
  }
  switch (spec) {
    case 1: return launch_tile_spec<FT, INTERP, ASSEMBLE, 1>(a, st);
    case 2: return launch_tile_spec<FT, INTERP, ASSEMBLE, 2>(a, st);
    default: return launch_tile_spec<FT, INTERP, ASSEMBLE, 0>(a, st);
  }
}

// ---------------------------------------------------------------------------------------------
// a3
// ---------------------------------------------------------------------------------------------
template <typename FT>
static int do_interpolate(coflux_ctx* c, const coflux_atmos_series* in, double time, coflux_exchange_state* out, cudaStream_t st) {
  FluxArgs<FT> a;
  zero_args(a);
  fill_geometry(c, a);
  int rc = fill_interp<FT>(c, in, time, out, a);
  if (rc) return rc;
  flux_kernel<FT, 0, true, false, false><<<grid_for(a.ncell, 128), 128, 0, st>>>(a);
  return check_launch(c, 1);
}
extern "C" int coflux_interpolate_atmosphere(coflux_ctx* c, const coflux_atmos_series* in, double time,
                                             coflux_exchange_state* out, void* stream) {
  REQUIRE(c, "NULL context");
  CUDA_TRY(cudaSetDevice(c->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return c->cfg.dtype == COFLUX_F64 ? do_interpolate<double>(c, in, time, out, st) : do_interpolate<float>(c, in, time, out, st);
}

// ---------------------------------------------------------------------------------------------
// a4–a6
// ---------------------------------------------------------------------------------------------
template <typename FT>
static int do_ao(coflux_ctx* c, const coflux_exchange_state* x, const coflux_ocean_surface* o, coflux_interface_fluxes* f, cudaStream_t st) {
  REQUIRE(f, "NULL interface fluxes");
  FluxArgs<FT> a;
  zero_args(a);
  fill_geometry(c, a);
  int rc = fill_exchange_in<FT>(x, a);
  if (rc) return rc;
  rc = fill_ocean<FT>(c, o, a);
  if (rc) return rc;
  rc = check_interface_halos<FT>(c, f);
  if (rc) return rc;
  fill_interface_out<FT>(f, a);
  if (tile_eligible<FT>(c, a)) {
    rc = launch_tile<FT, false, false>(c, a, st);
    if (rc) return rc;
  } else if (refill_v1() && (sizeof(FT) == 8 || COFLUX_REFILL_F32) && dev_params<FT>(c).ao.stop_kind != COFLUX_STOP_FIXED_ITERATIONS) {
    flux_refill_kernel<FT, 0, COFLUX_REFILL_TILE><<<grid_for(a.ncell - a.cell0, COFLUX_REFILL_TILE), 128, 0, st>>>(a);
  } else {
    flux_kernel<FT, 0, false, true, false><<<grid_for(a.ncell - a.cell0, 128), 128, 0, st>>>(a);
  }
  return check_launch(c, 1);
}
extern "C" int coflux_atmosphere_ocean_fluxes(coflux_ctx* c, const coflux_exchange_state* x, const coflux_ocean_surface* o,
                                              coflux_interface_fluxes* f, void* stream) {
  REQUIRE(c, "NULL context");
  CUDA_TRY(cudaSetDevice(c->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return c->cfg.dtype == COFLUX_F64 ? do_ao<double>(c, x, o, f, st) : do_ao<float>(c, x, o, f, st);
}

// ---------------------------------------------------------------------------------------------
// a7
// ---------------------------------------------------------------------------------------------
template <typename FT>
static int do_ai(coflux_ctx* c, const coflux_exchange_state* x, const coflux_ocean_surface* o, coflux_sea_ice_state* ice,
                 coflux_interface_fluxes* f, cudaStream_t st) {
  REQUIRE(f && ice, "NULL interface fluxes / sea ice state");
  REQUIRE(ice->u.ptr && ice->v.ptr && ice->top_temperature.ptr && ice->thickness.ptr && ice->concentration.ptr && ice->salinity.ptr,
          "sea ice u, v, top_temperature, thickness, concentration, salinity are required");
  FluxArgs<FT> a;
  zero_args(a);
  fill_geometry(c, a);
  int rc = fill_exchange_in<FT>(x, a);
  if (rc) return rc;
  const size_t es = sizeof(FT);
  a.ou = view2d(ice->u, 0, es); a.ov = view2d(ice->v, 0, es); a.oT = view2d(ice->top_temperature, 0, es);
  a.oS = DArr{nullptr, 0, 0};
  a.mask = o ? view2d(o->mask, 0, 1) : DArr{nullptr, 0, 0};
  a.ih = view2d(ice->thickness, 0, es); a.iS = view2d(ice->salinity, 0, es); a.ialb = view2d(ice->albedo, 0, es);
  a.iconc = view2d(ice->concentration, 0, es); a.ihs = view2d(ice->snow_thickness, 0, es);
  rc = check_interface_halos<FT>(c, f);
  if (rc) return rc;
  fill_interface_out<FT>(f, a);
  a.Ttop_out = view2d(ice->top_temperature, 0, es);
  // COFLUX_ICE_QUEUE=1: queue form (cells wait in global memory, thousands per tile; needs the six output arrays it parks a
  // cell's state in).  Bit-identical, but SLOWER on B200 (1/12°, 92 % ice cover: 19.9 vs 16.9 ms `:default`, 15.2 vs 13.1
  // `:corrected`, 13.3 vs 12.6 `:ncar`, Float32 13.6 vs 11.0): with ≈ one lane of a warp popping a cell per pass, the warp
  // waits for a DRAM round trip per pass, which costs more than the idle lanes of the tile form.  Kept as an A/B knob.
  static const bool queue_on = [] { const char* e = std::getenv("COFLUX_ICE_QUEUE"); return e && e[0] == '1'; }();
  if (ice_tile_eligible<FT>(c) && queue_on && a.Qv.p && a.Qc.p && a.Fv.p && a.rtx.p && a.rty.p && a.Tsout.p) {
    constexpr int TILE = COFLUX_ICE_QUEUE_CELLS;
    static int sm_count[64] = {};
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (!sm_count[dev & 63]) CUDA_TRY(cudaDeviceGetAttribute(&sm_count[dev & 63], cudaDevAttrMultiProcessorCount, dev));
    // every CTA the same number of cells: one wave when the grid fits (≥ 2 cells per lane), else a whole number of waves
    const long long n = a.ncell - a.cell0, slots = (long long)sm_count[dev & 63] * COFLUX_ICE_MIN_BLOCKS;
    long long per = std::max<long long>(256, (n + slots - 1) / slots);
    if (per > TILE) {
      const long long waves = (n + slots * TILE - 1) / (slots * TILE);
      per = (n + waves * slots - 1) / (waves * slots);
    }
    a.tile_cells = (int)per;
    ice_queue_kernel<FT, TILE><<<(unsigned)((n + per - 1) / per), 128, 0, st>>>(a);
  } else if (ice_tile_eligible<FT>(c)) {
    constexpr int TILE = COFLUX_ICE_TILE_CELLS;
    auto kern = ice_tile_kernel<FT, TILE>;
    const size_t smem = sizeof(IceTileSmem<FT, TILE>);
    static unsigned long long configured = 0;   // one bit per device
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (!(configured >> (dev & 63) & 1ull)) {
      CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      // no shared-memory carve-out preference: the driver picks the smallest carve-out that fits 4 CTAs; the pass reads its
      // log / exp / ψ tables from global memory through what is left of L1
      configured |= 1ull << (dev & 63);
    }
    kern<<<grid_for(a.ncell, TILE), 128, smem, st>>>(a);
  } else if (refill_v1() && (sizeof(FT) == 8 || COFLUX_REFILL_F32)) {
    flux_refill_kernel<FT, 1, COFLUX_REFILL_TILE><<<grid_for(a.ncell, COFLUX_REFILL_TILE), 128, 0, st>>>(a);
  } else {
    flux_kernel<FT, 1, false, true, false><<<grid_for(a.ncell, 128), 128, 0, st>>>(a);
  }
  return check_launch(c, 1);
}
extern "C" int coflux_atmosphere_sea_ice_fluxes(coflux_ctx* c, const coflux_exchange_state* x, const coflux_ocean_surface* o,
                                                coflux_sea_ice_state* ice, coflux_interface_fluxes* f, void* stream) {
  REQUIRE(c, "NULL context");
  CUDA_TRY(cudaSetDevice(c->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return c->cfg.dtype == COFLUX_F64 ? do_ai<double>(c, x, o, ice, f, st) : do_ai<float>(c, x, o, ice, f, st);
}

// ---------------------------------------------------------------------------------------------
// a8
// ---------------------------------------------------------------------------------------------
template <typename FT>
static int do_io(coflux_ctx* c, coflux_ocean_columns* oc, coflux_sea_ice_state* ice, double dt, coflux_ice_ocean_fluxes* f, cudaStream_t st) {
  REQUIRE(oc && ice && f, "NULL argument");
  REQUIRE(oc->T.ptr && oc->S.ptr && oc->dz.ptr && oc->u.ptr && oc->v.ptr, "ocean columns T, S, dz, u, v are required");
  REQUIRE(ice->thickness.ptr && ice->previous_thickness.ptr && ice->concentration.ptr && ice->salinity.ptr && ice->u.ptr && ice->v.ptr,
          "sea ice thickness, previous_thickness, concentration, salinity, u, v are required");
  REQUIRE(std::isfinite(dt) && dt > 0, "dt must be finite and > 0");
  const coflux_grid_desc& g = c->cfg.grid;
  const size_t es = sizeof(FT);
  IceOceanArgs<FT> a;
  memset(&a, 0, sizeof(a));
  a.Nx = g.Nx; a.Ny = g.Ny; a.Nz = g.Nz;
  a.T = view3d(oc->T, es); a.S = view3d(oc->S, es); a.dz = view3d(oc->dz, es);
  a.ou = view2d(oc->u, g.Nz - 1, es); a.ov = view2d(oc->v, g.Nz - 1, es);
  a.iu = view2d(ice->u, 0, es); a.iv = view2d(ice->v, 0, es); a.ih = view2d(ice->thickness, 0, es);
  a.ihm = view2d(ice->previous_thickness, 0, es); a.iconc = view2d(ice->concentration, 0, es); a.iS = view2d(ice->salinity, 0, es);
  a.Qf = view2d(f->frazil_heat, 0, es); a.Qio = view2d(f->interface_heat, 0, es); a.Js = view2d(f->salt, 0, es);
  a.tx = view2d(f->x_momentum, 0, es); a.ty = view2d(f->y_momentum, 0, es);
  a.dt = (FT)dt;
  if (c->avg_on) { a.avg_JTf = view2d(c->avg.JT_frazil, 0, es); a.avg_T = (FT)c->avg.previous_interval; a.avg_dt = (FT)c->avg.dt; }
  a.P = dev_params<FT>(c);
  // The bulk-asynchronous form (cp.async.bulk → shared memory, mbarrier pipeline) whenever the columns are contiguous in i
  // and the parents have a halo to absorb the 16-byte alignment slack: 1.66 ms = 93 % of the measured HBM peak at 1/12°,
  // Nz = 75, Float64, against 2.26 ms = 69 % for the register-staged kernel (bit-identical results; profiles/README.md).
  // COFLUX_IO_BULK=0 selects the register-staged kernel (A/B runs).
  static const bool bulk_off = [] { const char* e = std::getenv("COFLUX_IO_BULK"); return e && e[0] == '0'; }();
  const bool can_bulk = !bulk_off && a.T.si == 1 && a.S.si == 1 && oc->T.off_i >= (int)(16 / es) && oc->S.off_i >= (int)(16 / es) &&
                        ((uintptr_t)oc->T.ptr % 16 == 0) && ((uintptr_t)oc->S.ptr % 16 == 0);
  if (can_bulk) {
    constexpr int W = (sizeof(FT) == 8) ? COFLUX_IOB_W : COFLUX_IOB_W32, KB = COFLUX_IOB_KB, STAGES = COFLUX_IOB_STAGES;
    auto kern = ice_ocean_bulk_kernel<FT, W, KB, STAGES>;
    const size_t smem = sizeof(IceOceanBulkSmem<FT, W, KB, STAGES>);
    static unsigned long long configured = 0;
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (!(configured >> (dev & 63) & 1ull)) {
      CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
      configured |= 1ull << (dev & 63);
    }
    const long long tiles = (long long)((g.Nx + W - 1) / W) * g.Ny;
    kern<<<(unsigned)tiles, W, smem, st>>>(a);
  } else {
    ice_ocean_kernel<FT><<<grid_for((long long)g.Nx * g.Ny, COFLUX_IO_BLOCK), COFLUX_IO_BLOCK, 0, st>>>(a);
  }
  return check_launch(c, 1);
}
extern "C" int coflux_sea_ice_ocean_fluxes(coflux_ctx* c, coflux_ocean_columns* oc, coflux_sea_ice_state* ice, double dt,
                                           coflux_ice_ocean_fluxes* f, void* stream) {
  REQUIRE(c, "NULL context");
  CUDA_TRY(cudaSetDevice(c->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return c->cfg.dtype == COFLUX_F64 ? do_io<double>(c, oc, ice, dt, f, st) : do_io<float>(c, oc, ice, dt, f, st);
}

// ---------------------------------------------------------------------------------------------
// a9
// ---------------------------------------------------------------------------------------------
template <typename FT>
static void fill_closure(const coflux_closure_forcing& f, const coflux_net_ocean_fluxes* net, ClosureArgs<FT>& k) {
  const size_t es = sizeof(FT);
  k.on = 1;
  k.JT = view2d(net->T, 0, es); k.JS = view2d(net->S, 0, es);
  k.alpha = view2d(f.thermal_expansion, 0, es); k.beta = view2d(f.haline_contraction, 0, es);
  k.ustar = view2d(f.friction_velocity, 0, es); k.ustar2 = view2d(f.friction_velocity_squared, 0, es);
  k.tke = view2d(f.surface_tke, 0, es); k.Bo = view2d(f.buoyancy_flux, 0, es);
  k.umin = (FT)f.minimum_friction_velocity; k.emin = (FT)f.minimum_surface_tke; k.Cb = (FT)f.Cb; k.g = (FT)f.gravitational_acceleration;
}
template <typename FT>
static void fill_stress(const coflux_ctx* c, const coflux_ocean_surface* o, const coflux_interface_fluxes* ao,
                        const coflux_sea_ice_state* ice, const coflux_ice_ocean_fluxes* io, coflux_net_ocean_fluxes* out,
                        StressArgs<FT>& s) {
  const coflux_grid_desc& g = c->cfg.grid;
  const size_t es = sizeof(FT);
  s.Nx = g.Nx; s.Ny = g.Ny;
  s.wrap_x = (g.ring == 0 && g.periodic_x && !c->seam.attached) ? 1 : 0;
  s.rtx = view2d(ao->x_momentum, 0, es); s.rty = view2d(ao->y_momentum, 0, es);
  s.conc = ice ? view2d(ice->concentration, 0, es) : DArr{nullptr, 0, 0};
  s.tx_io = io ? view2d(io->x_momentum, 0, es) : DArr{nullptr, 0, 0};
  s.ty_io = io ? view2d(io->y_momentum, 0, es) : DArr{nullptr, 0, 0};
  s.mask = o ? view2d(o->mask, 0, 1) : DArr{nullptr, 0, 0};
  s.taux = view2d(out->u, 0, es); s.tauy = view2d(out->v, 0, es);
  s.seam_west = nullptr;
  s.rho0 = dev_params<FT>(c).rho0; s.rho0inv = dev_params<FT>(c).rho0inv;
  s.cell0 = 0; s.cell1 = (long long)g.Nx * g.Ny;
  if (c->closure_on) fill_closure<FT>(c->closure, out, s.closure);
}
template <typename FT>
static int do_assemble(coflux_ctx* c, const coflux_exchange_state* x, const coflux_ocean_surface* o, const coflux_interface_fluxes* ao,
                       const coflux_sea_ice_state* ice, const coflux_ice_ocean_fluxes* io, coflux_net_ocean_fluxes* out, cudaStream_t st) {
  REQUIRE(x && o && ao && out, "NULL argument");
  REQUIRE(x->Qs.ptr && x->Ql.ptr && x->Mp.ptr, "exchange state Qs, Ql, Mp are required");
  REQUIRE(ao->latent_heat.ptr && ao->sensible_heat.ptr && ao->water_vapor.ptr && ao->x_momentum.ptr && ao->y_momentum.ptr &&
              ao->interface_temperature.ptr, "atmosphere-ocean interface fluxes are required");
  REQUIRE(o->S.ptr, "ocean S is required");
  const coflux_grid_desc& g = c->cfg.grid;
  const size_t es = sizeof(FT);
  AssembleArgs<FT> a;
  memset(&a, 0, sizeof(a));
  fill_stress<FT>(c, o, ao, ice, io, out, a.s);
  a.oS = view2d(o->S, g.Nz - 1, es); a.Ts = view2d(ao->interface_temperature, 0, es);
  a.xQs = view2d(x->Qs, 0, es); a.xQl = view2d(x->Ql, 0, es); a.xMp = view2d(x->Mp, 0, es);
  a.Qc = view2d(ao->sensible_heat, 0, es); a.Qv = view2d(ao->latent_heat, 0, es); a.Fv = view2d(ao->water_vapor, 0, es);
  a.Qio = io ? view2d(io->interface_heat, 0, es) : DArr{nullptr, 0, 0};
  a.salt_io = io ? view2d(io->salt, 0, es) : DArr{nullptr, 0, 0};
  a.JT = view2d(out->T, 0, es); a.JS = view2d(out->S, 0, es); a.Qu = view2d(out->upwelling_longwave, 0, es);
  a.Qal = view2d(out->downwelling_longwave, 0, es); a.Qts = view2d(out->downwelling_shortwave, 0, es);
  a.J0 = view2d(out->penetrating_shortwave, 0, es);
  a.P = dev_params<FT>(c);
  assemble_kernel<FT><<<grid_for((long long)g.Nx * g.Ny, 256), 256, 0, st>>>(a);
  return check_launch(c, 1);
}
extern "C" int coflux_assemble_net_ocean_fluxes(coflux_ctx* c, const coflux_exchange_state* x, const coflux_ocean_surface* o,
                                                const coflux_interface_fluxes* ao, const coflux_sea_ice_state* ice,
                                                const coflux_ice_ocean_fluxes* io, coflux_net_ocean_fluxes* out, void* stream) {
  REQUIRE(c, "NULL context");
  CUDA_TRY(cudaSetDevice(c->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return c->cfg.dtype == COFLUX_F64 ? do_assemble<double>(c, x, o, ao, ice, io, out, st) : do_assemble<float>(c, x, o, ao, ice, io, out, st);
}

// ---------------------------------------------------------------------------------------------
// a2: fused update_state!
// ---------------------------------------------------------------------------------------------
template <typename FT>
static int build_update_args(coflux_ctx* c, const coflux_update_inputs* in, coflux_update_outputs* out, double time, FluxArgs<FT>& a,
                             StressArgs<FT>& s) {
  REQUIRE(in && out, "NULL argument");
  REQUIRE(in->atmosphere && in->ocean, "atmosphere series and ocean surface are required");
  REQUIRE(out->exchange && out->atmosphere_ocean && out->net_ocean, "exchange, atmosphere_ocean and net_ocean outputs are required");
  REQUIRE(out->atmosphere_ocean->x_momentum.ptr && out->atmosphere_ocean->y_momentum.ptr,
          "x_momentum / y_momentum outputs are required (the stress kernel reads them back)");
  {
    const coflux_exchange_state* x = out->exchange;
    REQUIRE(x->u.ptr && x->v.ptr && x->T.ptr && x->p.ptr && x->q.ptr && x->Qs.ptr && x->Ql.ptr && x->Mp.ptr,
            "all eight exchange-state output arrays are required");
  }
  const size_t es = sizeof(FT);
  zero_args(a);
  fill_geometry(c, a);
  int rc = fill_interp<FT>(c, in->atmosphere, time, out->exchange, a);
  if (rc) return rc;
  rc = fill_ocean<FT>(c, in->ocean, a);
  if (rc) return rc;
  rc = check_interface_halos<FT>(c, out->atmosphere_ocean);
  if (rc) return rc;
  fill_interface_out<FT>(out->atmosphere_ocean, a);
  const coflux_sea_ice_state* ice = in->sea_ice;
  const coflux_ice_ocean_fluxes* io = in->ice_ocean;
  a.conc = ice ? view2d(ice->concentration, 0, es) : DArr{nullptr, 0, 0};
  a.Qio = io ? view2d(io->interface_heat, 0, es) : DArr{nullptr, 0, 0};
  a.salt_io = io ? view2d(io->salt, 0, es) : DArr{nullptr, 0, 0};
  coflux_net_ocean_fluxes* n = out->net_ocean;
  a.JT = view2d(n->T, 0, es); a.JS = view2d(n->S, 0, es); a.Qu = view2d(n->upwelling_longwave, 0, es);
  a.Qal = view2d(n->downwelling_longwave, 0, es); a.Qts = view2d(n->downwelling_shortwave, 0, es);
  a.J0 = view2d(n->penetrating_shortwave, 0, es);
  rc = fill_land<FT>(c, in->land, time, a);
  if (rc) return rc;
  fill_avg<FT>(c, a.avg);
  memset(&s, 0, sizeof(s));
  fill_stress<FT>(c, in->ocean, out->atmosphere_ocean, ice, io, n, s);
  if (c->avg_on) {
    s.avg_tx = view2d(c->avg.tau_x, 0, es); s.avg_ty = view2d(c->avg.tau_y, 0, es);
    s.avg_T = (FT)c->avg.previous_interval; s.avg_dt = (FT)c->avg.dt;
  }
  return COFLUX_OK;
}
// flux kernel over rows jj ∈ [jj_lo, jj_hi) of the ring-extended surface
template <typename FT> static int launch_flux_rows(coflux_ctx* c, FluxArgs<FT> a, int jj_lo, int jj_hi, cudaStream_t st) {
  a.cell0 = (long long)jj_lo * a.nxr;
  a.ncell = (long long)jj_hi * a.nxr;
  if (a.ncell <= a.cell0) return COFLUX_OK;
  if (tile_eligible<FT>(c, a)) {
    int rc = launch_tile<FT, true, true>(c, a, st);
    if (rc) return rc;
  } else {
    if (uniform_layout<FT>(a)) flux_kernel<FT, 0, true, true, true, true><<<grid_for(a.ncell - a.cell0, 128), 128, 0, st>>>(a);
    else flux_kernel<FT, 0, true, true, true><<<grid_for(a.ncell - a.cell0, 128), 128, 0, st>>>(a);
  }
  return check_launch(c, 1);
}
// stress kernel over interior rows j ∈ [j_lo, j_hi)
template <typename FT> static int launch_stress_rows(coflux_ctx* c, StressArgs<FT> s, int j_lo, int j_hi, cudaStream_t st) {
  s.cell0 = (long long)j_lo * s.Nx;
  s.cell1 = (long long)j_hi * s.Nx;
  if (s.cell1 <= s.cell0) return COFLUX_OK;
  stress_kernel<FT><<<grid_for(s.cell1 - s.cell0, 256), 256, 0, st>>>(s);
  return check_launch(c, 1);
}
template <typename FT>
static int do_update(coflux_ctx* c, const coflux_update_inputs* in, coflux_update_outputs* out, double time, cudaStream_t st) {
  FluxArgs<FT> a;
  StressArgs<FT> s;
  int rc = build_update_args<FT>(c, in, out, time, a, s);
  if (rc) return rc;
  Profile& pf = c->prof;
  if (pf.on && pf.pending == Profile::RING) { rc = profile_drain(c); if (rc) return rc; }
  Seam& sm = c->seam;
  unsigned int step = 0, parity = 0;
  if (sm.attached) {
    step = ++sm.step; parity = step & 1u;
    const size_t col = (size_t)s.Ny * sizeof(FT);
    // the east neighbour must have consumed this parity's previous column (step − 2) before we overwrite it
    if (step > 2 && g_wait32(st, (unsigned long long)(uintptr_t)(sm.local + 8), step - 2, COFLUX_WAIT_GEQ) != 0)
      return fail(COFLUX_ERR_SEAM, "cuStreamWaitValue32 (ack) failed");
    a.seam_east = sm.east + SEAM_DATA_OFFSET + parity * col;
    s.seam_west = sm.local + SEAM_DATA_OFFSET + parity * col;
  }
  if (pf.on) CUDA_TRY(cudaEventRecord(pf.ev[pf.pending][0], st));
  rc = launch_flux_rows<FT>(c, a, 0, a.nyr, st);
  if (rc) return rc;
  if (sm.attached) {
    // publish: our column for `step` is in the east neighbour's buffer; then wait for the west neighbour's
    if (g_write32(st, (unsigned long long)(uintptr_t)(sm.east + 4 * parity), step, 0) != 0) return fail(COFLUX_ERR_SEAM, "cuStreamWriteValue32 failed");
    if (g_wait32(st, (unsigned long long)(uintptr_t)(sm.local + 4 * parity), step, COFLUX_WAIT_GEQ) != 0) return fail(COFLUX_ERR_SEAM, "cuStreamWaitValue32 failed");
  }
  if (pf.on) CUDA_TRY(cudaEventRecord(pf.ev[pf.pending][1], st));
  rc = launch_stress_rows<FT>(c, s, 0, s.Ny, st);
  if (rc) return rc;
  if (sm.attached && g_write32(st, (unsigned long long)(uintptr_t)(sm.west + 8), step, 0) != 0) return fail(COFLUX_ERR_SEAM, "cuStreamWriteValue32 (ack) failed");
  if (pf.on) { CUDA_TRY(cudaEventRecord(pf.ev[pf.pending][2], st)); pf.pending += 1; }
  return COFLUX_OK;
}
extern "C" int coflux_update_state(coflux_ctx* c, const coflux_update_inputs* in, coflux_update_outputs* out, double time, void* stream) {
  REQUIRE(c, "NULL context");
  CUDA_TRY(cudaSetDevice(c->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return c->cfg.dtype == COFLUX_F64 ? do_update<double>(c, in, out, time, st) : do_update<float>(c, in, out, time, st);
}

// ---------------------------------------------------------------------------------------------
// end-to-end entry with HOST buffers
// ---------------------------------------------------------------------------------------------
static coflux_array plane_desc(void* p, int Nx, int halo) {
  coflux_array a;
  memset(&a, 0, sizeof(a));
  a.ptr = p;
  a.stride_i = 1; a.stride_j = Nx + 2 * halo; a.stride_k = 0; a.stride_n = 0;
  a.off_i = halo; a.off_j = halo; a.off_k = 0;
  return a;
}
template <typename FT>
static int do_update_host(coflux_ctx* c, const coflux_atmos_series* atm, const coflux_host_step* step, double time, int64_t* h2d_bytes,
                          int64_t* d2h_bytes) {
  const coflux_grid_desc& g = c->cfg.grid;
  const size_t es = sizeof(FT);
  const int H = step->halo;
  const int ni = g.Nx + 2 * H, nj = g.Ny + 2 * H;
  const size_t row = (size_t)ni * es, plane = row * (size_t)nj;
  HostStage& s = c->stage;
  if (s.halo != H || s.plane_bytes != plane) {
    free_stage(s);
    s.halo = H; s.plane_bytes = plane;
    for (char*& p : s.in) CUDA_TRY(cudaMalloc(&p, plane));
    for (char*& p : s.xch) CUDA_TRY(cudaMalloc(&p, plane));
    for (char*& p : s.ao) CUDA_TRY(cudaMalloc(&p, plane));
    for (char*& p : s.net) CUDA_TRY(cudaMalloc(&p, plane));
    for (char* p : s.net) CUDA_TRY(cudaMemset(p, 0, plane));
    for (char* p : s.ao) CUDA_TRY(cudaMemset(p, 0, plane));
    CUDA_TRY(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&s.stream2, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&s.copy_in, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&s.copy_out, cudaStreamNonBlocking));
    for (cudaEvent_t& ev : s.ev_in) CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    for (cudaEvent_t& ev : s.ev_k) CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    for (cudaEvent_t& ev : s.ev_f) CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    CUDA_TRY(cudaDeviceSynchronize());
  }
  coflux_ocean_surface ocean;
  memset(&ocean, 0, sizeof(ocean));
  // planes are the k = Nz-1 level: present them as 3-D parents with stride_k = 0
  ocean.u = plane_desc(s.in[0], g.Nx, H); ocean.v = plane_desc(s.in[1], g.Nx, H);
  ocean.T = plane_desc(s.in[2], g.Nx, H); ocean.S = plane_desc(s.in[3], g.Nx, H);
  coflux_exchange_state xch;
  coflux_array* xa[8] = {&xch.u, &xch.v, &xch.T, &xch.p, &xch.q, &xch.Qs, &xch.Ql, &xch.Mp};
  for (int k = 0; k < 8; ++k) *xa[k] = plane_desc(s.xch[k], g.Nx, H);
  coflux_interface_fluxes ao;
  memset(&ao, 0, sizeof(ao));
  ao.latent_heat = plane_desc(s.ao[0], g.Nx, H); ao.sensible_heat = plane_desc(s.ao[1], g.Nx, H);
  ao.water_vapor = plane_desc(s.ao[2], g.Nx, H); ao.x_momentum = plane_desc(s.ao[3], g.Nx, H);
  ao.y_momentum = plane_desc(s.ao[4], g.Nx, H); ao.interface_temperature = plane_desc(s.ao[5], g.Nx, H);
  coflux_net_ocean_fluxes net;
  coflux_array* na[8] = {&net.u, &net.v, &net.T, &net.S, &net.upwelling_longwave, &net.downwelling_longwave,
                         &net.downwelling_shortwave, &net.penetrating_shortwave};
  for (int k = 0; k < 8; ++k) *na[k] = plane_desc(s.net[k], g.Nx, H);
  coflux_update_inputs in;
  memset(&in, 0, sizeof(in));
  in.atmosphere = atm; in.ocean = &ocean;
  coflux_update_outputs out;
  out.exchange = &xch; out.atmosphere_ocean = &ao; out.net_ocean = &net;
  FluxArgs<FT> a;
  StressArgs<FT> sa;
  int rc = build_update_args<FT>(c, &in, &out, time, a, sa);
  if (rc) return rc;

  // Row-chunk pipeline: H2D(c+1) ‖ kernels(c) ‖ D2H(c-1) on three streams.  Chunk c of the flux kernel
  // covers ring-extended rows [f[c], f[c+1]); it reads ocean rows up to one past its last row, i.e. the
  // first parent row of H2D chunk c+1, so it waits for that chunk.  The stress rows that became
  // computable (they need ρτy of the previous row) follow, then their D2H.
  const int ring = g.ring, nyr = a.nyr;
  // 16 equal row chunks, kernels of consecutive chunks on two alternating streams: a chunk kernel is only 1 300 CTAs
  // (1.4 waves), so back-to-back launches on ONE stream leave the SMs idle in every tail and the kernels — not PCIe —
  // bound the pipeline (measured: 8 chunks on one stream 7.8 ms per step at 1/12° Float64, while PCIe moves the 252 MB
  // each way in 5.8 ms, tools/pcie_probe.py).  With two streams the next chunk's CTAs fill the tail of the previous one.
  // Measured and NOT used (round 2): letting the kernels store the results straight into the caller's mapped pinned planes
  // instead of staging them and copying device→host — 8.1 ms against 7.5 ms per step at 1/12° Float64; the copy engine moves
  // the data faster than the SMs' stores over PCIe.  Chunk sweep (COFLUX_HOST_CHUNKS): 2 / 4 / 6 / 8 / 10–15 / 16 chunks:
  // 12.2 / 8.8 / 8.0 / 7.7 / 7.6 / 7.5 ms.
  // chunk count: ≥ 4 MB per plane copy (below that the ~10 µs per cudaMemcpyAsync / launch dominate: at 8 slabs the fixed
  // 16-chunk pipeline issued ≈ 160 tiny operations per step and end-to-end scaled 2.1× on 8 GPUs), at most MAX_CHUNKS
  int nch = (int)std::min<size_t>(HostStage::MAX_CHUNKS, std::max<size_t>(1, plane / (4u << 20)));
  static const int forced_chunks = [] { const char* e = std::getenv("COFLUX_HOST_CHUNKS"); return e ? atoi(e) : 0; }();   // A/B knob
  if (forced_chunks > 0) nch = std::min(forced_chunks, (int)HostStage::MAX_CHUNKS);
  while (nch > 1 && g.Ny < 8 * nch) nch /= 2;
  int f[HostStage::MAX_CHUNKS + 1], r[HostStage::MAX_CHUNKS + 1], sj[HostStage::MAX_CHUNKS + 1];
  for (int k = 0; k <= nch; ++k) f[k] = (int)((long long)nyr * k / nch);
  r[0] = 0; r[nch] = nj;
  for (int k = 1; k < nch; ++k) r[k] = f[k] - ring + H;            // parent row of surface row jj = f[k]
  sj[0] = 0; sj[nch] = g.Ny;
  for (int k = 1; k < nch; ++k) { int v = f[k] - ring - 1; sj[k] = v < 0 ? 0 : (v > g.Ny ? g.Ny : v); }
  const void* hin[4] = {step->ocean_u, step->ocean_v, step->ocean_T, step->ocean_S};
  void* hout[6] = {step->net_u, step->net_v, step->net_T, step->net_S, step->latent_heat, step->sensible_heat};
  char* dout[6] = {s.net[0], s.net[1], s.net[2], s.net[3], s.ao[0], s.ao[1]};
  int64_t h2d = 0, d2h = 0;
  for (int k = 0; k < nch; ++k) {
    const size_t off = (size_t)r[k] * row, len = (size_t)(r[k + 1] - r[k]) * row;
    for (int q = 0; q < 4; ++q)
      CUDA_TRY(cudaMemcpyAsync(s.in[q] + off, static_cast<const char*>(hin[q]) + off, len, cudaMemcpyHostToDevice, s.copy_in));
    h2d += 4 * (int64_t)len;
    CUDA_TRY(cudaEventRecord(s.ev_in[k], s.copy_in));
  }
  for (int k = 0; k < nch; ++k) {
    cudaStream_t st = (k & 1) ? s.stream2 : s.stream;
    CUDA_TRY(cudaStreamWaitEvent(st, s.ev_in[(k + 1 < nch) ? k + 1 : k], 0));
    rc = launch_flux_rows<FT>(c, a, f[k], f[k + 1], st);
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(s.ev_f[k], st));
    if (k > 0) CUDA_TRY(cudaStreamWaitEvent(st, s.ev_f[k - 1], 0));   // the first stress row of the chunk reads ρτy of the row before
    rc = launch_stress_rows<FT>(c, sa, sj[k], sj[k + 1], st);
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(s.ev_k[k], st));
    CUDA_TRY(cudaStreamWaitEvent(s.copy_out, s.ev_k[k], 0));
    // D2H: the interior rows just completed; the first / last chunk also carry the (untouched) halo rows
    const int p0 = (k == 0) ? 0 : sj[k] + H, p1 = (k == nch - 1) ? nj : sj[k + 1] + H;
    if (p1 > p0) {
      const size_t off = (size_t)p0 * row, len = (size_t)(p1 - p0) * row;
      for (int q = 0; q < 6; ++q) {
        if (!hout[q]) continue;
        CUDA_TRY(cudaMemcpyAsync(static_cast<char*>(hout[q]) + off, dout[q] + off, len, cudaMemcpyDeviceToHost, s.copy_out));
        d2h += (int64_t)len;
      }
    }
  }
  CUDA_TRY(cudaStreamSynchronize(s.copy_out));
  CUDA_TRY(cudaStreamSynchronize(s.stream));
  CUDA_TRY(cudaStreamSynchronize(s.stream2));
  if (h2d_bytes) *h2d_bytes = h2d;
  if (d2h_bytes) *d2h_bytes = d2h;
  return COFLUX_OK;
}
extern "C" int coflux_update_state_host(coflux_ctx* c, const coflux_atmos_series* atm, const coflux_host_step* step, double time,
                                        int64_t* h2d_bytes, int64_t* d2h_bytes) {
  REQUIRE(c && atm && step, "NULL argument");
  REQUIRE(step->ocean_u && step->ocean_v && step->ocean_T && step->ocean_S, "host ocean planes are required");
  REQUIRE(step->net_u && step->net_v && step->net_T && step->net_S, "host net-flux planes are required");
  REQUIRE(step->halo >= 2, "host planes need a halo of at least 2 cells");
  CUDA_TRY(cudaSetDevice(c->device));
  return c->cfg.dtype == COFLUX_F64 ? do_update_host<double>(c, atm, step, time, h2d_bytes, d2h_bytes)
                                    : do_update_host<float>(c, atm, step, time, h2d_bytes, d2h_bytes);
}


// ---------------------------------------------------------------------------------------------
// land freshwater, stand-alone (JRA55PrescribedLand: friver + licalvf → exchange Mp)
// ---------------------------------------------------------------------------------------------
template <typename FT>
static int do_land(coflux_ctx* c, const coflux_land_series* in, double time, coflux_exchange_state* x, cudaStream_t st) {
  FluxArgs<FT> a;
  zero_args(a);
  fill_geometry(c, a);
  REQUIRE(x->Mp.ptr, "exchange Mp is required");
  HALO(x->Mp, c->cfg.grid.ring, c->cfg.grid.ring, "exchange Mp");
  a.xMp = view2d(x->Mp, 0, sizeof(FT));
  int rc = fill_land<FT>(c, in, time, a);
  if (rc) return rc;
  land_kernel<FT><<<grid_for(a.ncell, 256), 256, 0, st>>>(a);
  return check_launch(c, 1);
}
extern "C" int coflux_interpolate_land(coflux_ctx* c, const coflux_land_series* in, double time, coflux_exchange_state* x, void* stream) {
  REQUIRE(c && in && x, "NULL argument");
  CUDA_TRY(cudaSetDevice(c->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return c->cfg.dtype == COFLUX_F64 ? do_land<double>(c, in, time, x, st) : do_land<float>(c, in, time, x, st);
}

// ---------------------------------------------------------------------------------------------
// compute_net_sea_ice_fluxes! (§8f row 1)
// ---------------------------------------------------------------------------------------------
template <typename FT>
static int do_net_ice(coflux_ctx* c, const coflux_exchange_state* x, const coflux_ocean_surface* o, const coflux_sea_ice_state* ice,
                      const coflux_interface_fluxes* ai, const coflux_ice_ocean_fluxes* io, coflux_net_sea_ice_fluxes* out, cudaStream_t st) {
  REQUIRE(x->Qs.ptr && x->Ql.ptr, "exchange state Qs, Ql are required");
  REQUIRE(ice->top_temperature.ptr && ice->concentration.ptr, "sea ice top_temperature and concentration are required");
  REQUIRE(ai->sensible_heat.ptr && ai->latent_heat.ptr, "atmosphere-sea-ice sensible and latent heat fluxes are required");
  REQUIRE(out->top_heat.ptr && out->bottom_heat.ptr, "top_heat and bottom_heat outputs are required");
  REQUIRE((!out->top_u.ptr || ai->x_momentum.ptr) && (!out->top_v.ptr || ai->y_momentum.ptr), "top stresses need the atmosphere-sea-ice momentum fluxes");
  const DevParams<FT>& P = dev_params<FT>(c);
  REQUIRE(P.ice_albedo_kind != COFLUX_SEA_ICE_ALBEDO_CCSM3 || ice->thickness.ptr, "the CCSM3 albedo needs the ice thickness");
  if (out->top_u.ptr) HALO(ai->x_momentum, 1, 0, "atmosphere-sea-ice x_momentum");
  if (out->top_v.ptr) HALO(ai->y_momentum, 0, 1, "atmosphere-sea-ice y_momentum");
  const coflux_grid_desc& g = c->cfg.grid;
  const size_t es = sizeof(FT);
  NetIceArgs<FT> a;
  memset(&a, 0, sizeof(a));
  a.Nx = g.Nx; a.Ny = g.Ny; a.wrap_x = (g.ring == 0 && g.periodic_x) ? 1 : 0;
  a.Qs = view2d(x->Qs, 0, es); a.Ql = view2d(x->Ql, 0, es);
  a.Ttop = view2d(ice->top_temperature, 0, es); a.conc = view2d(ice->concentration, 0, es); a.ih = view2d(ice->thickness, 0, es);
  a.ihs = view2d(ice->snow_thickness, 0, es); a.ialb = view2d(ice->albedo, 0, es);
  a.Qc = view2d(ai->sensible_heat, 0, es); a.Qv = view2d(ai->latent_heat, 0, es);
  a.rtx = view2d(ai->x_momentum, 0, es); a.rty = view2d(ai->y_momentum, 0, es);
  a.Qf = io ? view2d(io->frazil_heat, 0, es) : DArr{nullptr, 0, 0};
  a.Qi = io ? view2d(io->interface_heat, 0, es) : DArr{nullptr, 0, 0};
  a.mask = o ? view2d(o->mask, 0, 1) : DArr{nullptr, 0, 0};
  a.top = view2d(out->top_heat, 0, es); a.bottom = view2d(out->bottom_heat, 0, es);
  a.top_u = view2d(out->top_u, 0, es); a.top_v = view2d(out->top_v, 0, es);
  a.P = P;
  net_sea_ice_kernel<FT><<<grid_for((long long)g.Nx * g.Ny, 256), 256, 0, st>>>(a);
  return check_launch(c, 1);
}
extern "C" int coflux_assemble_net_sea_ice_fluxes(coflux_ctx* c, const coflux_exchange_state* x, const coflux_ocean_surface* o,
                                                  const coflux_sea_ice_state* ice, const coflux_interface_fluxes* ai,
                                                  const coflux_ice_ocean_fluxes* io, coflux_net_sea_ice_fluxes* out, void* stream) {
  REQUIRE(c && x && ice && ai && out, "NULL argument");
  CUDA_TRY(cudaSetDevice(c->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return c->cfg.dtype == COFLUX_F64 ? do_net_ice<double>(c, x, o, ice, ai, io, out, st) : do_net_ice<float>(c, x, o, ice, ai, io, out, st);
}

// ---------------------------------------------------------------------------------------------
// time-averaged flux diagnostics (§8f row 4; omip_diagnostics.jl:77-89, 125-158)
// ---------------------------------------------------------------------------------------------
static int check_averages(const coflux_flux_averages* v) {
  REQUIRE(std::isfinite(v->previous_interval) && std::isfinite(v->dt) && v->previous_interval >= 0 && v->dt > 0,
          "flux averages: previous_interval must be >= 0 and dt > 0");
  return COFLUX_OK;
}
extern "C" int coflux_attach_flux_averages(coflux_ctx* c, const coflux_flux_averages* v) {
  REQUIRE(c, "NULL context");
  if (!v) { c->avg_on = false; return COFLUX_OK; }
  int rc = check_averages(v);
  if (rc) return rc;
  c->avg = *v;
  c->avg_on = true;
  return COFLUX_OK;
}
template <typename FT>
static int do_averages(coflux_ctx* c, const coflux_net_ocean_fluxes* net, const coflux_interface_fluxes* ao, const coflux_sea_ice_state* ice,
                       const coflux_ice_ocean_fluxes* io, const coflux_flux_averages* v, cudaStream_t st) {
  const coflux_grid_desc& g = c->cfg.grid;
  const size_t es = sizeof(FT);
  FluxAvgArgs<FT> a;
  memset(&a, 0, sizeof(a));
  a.Nx = g.Nx; a.Ny = g.Ny;
  a.tx = view2d(net->u, 0, es); a.ty = view2d(net->v, 0, es); a.JT = view2d(net->T, 0, es); a.JS = view2d(net->S, 0, es);
  if (ao) { a.Qc = view2d(ao->sensible_heat, 0, es); a.Qv = view2d(ao->latent_heat, 0, es); }
  if (ice) a.conc = view2d(ice->concentration, 0, es);
  if (io) { a.Qio = view2d(io->interface_heat, 0, es); a.salt_io = view2d(io->salt, 0, es); a.Qf = view2d(io->frazil_heat, 0, es); }
  a.a_tx = view2d(v->tau_x, 0, es); a.a_ty = view2d(v->tau_y, 0, es); a.a_JT = view2d(v->JT, 0, es); a.a_JS = view2d(v->JS, 0, es);
  a.a_Qc = view2d(v->Qc, 0, es); a.a_Qv = view2d(v->Qv, 0, es); a.a_JTao = view2d(v->JT_atmosphere_ocean, 0, es);
  a.a_JTio = view2d(v->JT_ice_ocean, 0, es); a.a_JSio = view2d(v->JS_ice_ocean, 0, es); a.a_JTf = view2d(v->JT_frazil, 0, es);
  a.T = (FT)v->previous_interval; a.dt = (FT)v->dt;
  a.rho0 = dev_params<FT>(c).rho0; a.c0 = dev_params<FT>(c).c0;
  flux_average_kernel<FT><<<grid_for((long long)g.Nx * g.Ny, 256), 256, 0, st>>>(a);
  return check_launch(c, 1);
}
extern "C" int coflux_accumulate_flux_averages(coflux_ctx* c, const coflux_net_ocean_fluxes* net, const coflux_interface_fluxes* ao,
                                               const coflux_sea_ice_state* ice, const coflux_ice_ocean_fluxes* io,
                                               const coflux_flux_averages* v, void* stream) {
  REQUIRE(c && net && v, "NULL argument");
  int rc = check_averages(v);
  if (rc) return rc;
  CUDA_TRY(cudaSetDevice(c->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return c->cfg.dtype == COFLUX_F64 ? do_averages<double>(c, net, ao, ice, io, v, st) : do_averages<float>(c, net, ao, ice, io, v, st);
}

// ---------------------------------------------------------------------------------------------
// device forcing window (§8f row 2; time_indices_in_memory + prefetch, atmosphere.jl:22-27)
// ---------------------------------------------------------------------------------------------
struct coflux_forcing_window {
  int device = 0, n_fields = 0, capacity = 0;
  int64_t plane_elements = 0;
  size_t esize = 8;
  std::vector<char*> field;                 // per field: capacity × plane_elements elements
  std::vector<cudaEvent_t> uploaded;        // per slot: the upload of the level it holds has landed
  std::vector<cudaEvent_t> released;        // per slot: last reader enqueued so far has finished
  std::vector<char> has_release;
  std::vector<int64_t> level;               // per slot: logical level held (−1: none)
  cudaStream_t copy = nullptr;
  int64_t bytes = 0, levels = 0;
};
extern "C" int coflux_forcing_window_destroy(coflux_forcing_window* w) {
  if (!w) return COFLUX_OK;
  cudaSetDevice(w->device);
  if (w->copy) { cudaStreamSynchronize(w->copy); cudaStreamDestroy(w->copy); }
  for (char* p : w->field) if (p) cudaFree(p);
  for (cudaEvent_t e : w->uploaded) if (e) cudaEventDestroy(e);
  for (cudaEvent_t e : w->released) if (e) cudaEventDestroy(e);
  delete w;
  return COFLUX_OK;
}
extern "C" int coflux_forcing_window_create(coflux_forcing_window** out, coflux_ctx* c, int32_t n_fields, int64_t plane_elements, int32_t capacity) {
  REQUIRE(out && c, "NULL argument");
  *out = nullptr;
  REQUIRE(n_fields >= 1 && n_fields <= 64 && plane_elements >= 1 && capacity >= 2, "forcing window: need 1..64 fields, a non-empty plane and capacity >= 2 levels");
  CUDA_TRY(cudaSetDevice(c->device));
  coflux_forcing_window* w = new (std::nothrow) coflux_forcing_window();
  if (!w) return fail(COFLUX_ERR_ALLOC, "out of host memory");
  w->device = c->device; w->n_fields = n_fields; w->capacity = capacity; w->plane_elements = plane_elements;
  w->esize = (c->cfg.dtype == COFLUX_F64) ? 8 : 4;
  w->field.assign(n_fields, nullptr);
  w->uploaded.assign(capacity, nullptr); w->released.assign(capacity, nullptr);
  w->has_release.assign(capacity, 0); w->level.assign(capacity, -1);
  const size_t bytes = (size_t)capacity * (size_t)plane_elements * w->esize;
  for (int f = 0; f < n_fields; ++f) {
    if (cudaMalloc(&w->field[f], bytes) != cudaSuccess) { cudaGetLastError(); coflux_forcing_window_destroy(w); return fail(COFLUX_ERR_ALLOC, "forcing window: cudaMalloc of %zu bytes failed", bytes); }
  }
  bool ok = cudaStreamCreateWithFlags(&w->copy, cudaStreamNonBlocking) == cudaSuccess;
  for (int s = 0; s < capacity && ok; ++s)
    ok = cudaEventCreateWithFlags(&w->uploaded[s], cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&w->released[s], cudaEventDisableTiming) == cudaSuccess;
  if (!ok) { cudaGetLastError(); coflux_forcing_window_destroy(w); return fail(COFLUX_ERR_CUDA, "forcing window: stream / event creation failed"); }
  *out = w;
  return COFLUX_OK;
}
extern "C" int coflux_forcing_window_upload(coflux_forcing_window* w, int64_t level, const void* const* host_planes) {
  REQUIRE(w && host_planes && level >= 0, "forcing window upload: bad argument");
  CUDA_TRY(cudaSetDevice(w->device));
  const int slot = (int)(level % w->capacity);
  // the slot's previous level may still be read by work enqueued on a compute stream: wait for its last reader ON THE
  // COPY STREAM (the host does not block)
  if (w->has_release[slot]) CUDA_TRY(cudaStreamWaitEvent(w->copy, w->released[slot], 0));
  const size_t plane = (size_t)w->plane_elements * w->esize;
  for (int f = 0; f < w->n_fields; ++f) {
    REQUIRE(host_planes[f], "forcing window upload: NULL host plane");
    CUDA_TRY(cudaMemcpyAsync(w->field[f] + (size_t)slot * plane, host_planes[f], plane, cudaMemcpyHostToDevice, w->copy));
  }
  CUDA_TRY(cudaEventRecord(w->uploaded[slot], w->copy));
  w->level[slot] = level;
  w->has_release[slot] = 0;
  w->bytes += (int64_t)(plane * w->n_fields); w->levels += 1;
  return COFLUX_OK;
}
extern "C" int coflux_forcing_window_field(coflux_forcing_window* w, int32_t f, void** ptr) {
  REQUIRE(w && ptr && f >= 0 && f < w->n_fields, "forcing window field: bad argument");
  *ptr = w->field[f];
  return COFLUX_OK;
}
extern "C" int coflux_forcing_window_wait(coflux_forcing_window* w, int64_t first, int64_t last, void* stream) {
  REQUIRE(w && first >= 0 && last >= first && last - first < w->capacity, "forcing window wait: bad level range");
  CUDA_TRY(cudaSetDevice(w->device));
  for (int64_t l = first; l <= last; ++l) {
    const int slot = (int)(l % w->capacity);
    if (w->level[slot] != l) return fail(COFLUX_ERR_INVALID_ARGUMENT, "forcing window: level %lld is not in memory (slot %d holds level %lld)", (long long)l, slot, (long long)w->level[slot]);
    CUDA_TRY(cudaStreamWaitEvent(static_cast<cudaStream_t>(stream), w->uploaded[slot], 0));
  }
  return COFLUX_OK;
}
extern "C" int coflux_forcing_window_release(coflux_forcing_window* w, int64_t first, int64_t last, void* stream) {
  REQUIRE(w && first >= 0 && last >= first && last - first < w->capacity, "forcing window release: bad level range");
  CUDA_TRY(cudaSetDevice(w->device));
  for (int64_t l = first; l <= last; ++l) {
    const int slot = (int)(l % w->capacity);
    if (w->level[slot] != l) continue;
    CUDA_TRY(cudaEventRecord(w->released[slot], static_cast<cudaStream_t>(stream)));
    w->has_release[slot] = 1;
  }
  return COFLUX_OK;
}
extern "C" int coflux_forcing_window_stats(coflux_forcing_window* w, int64_t* bytes, int64_t* levels) {
  REQUIRE(w, "NULL window");
  if (bytes) *bytes = w->bytes;
  if (levels) *levels = w->levels;
  return COFLUX_OK;
}

// ---------------------------------------------------------------------------------------------
// closure surface-forcing front ends (KPP/kpp_surface_forcing.jl, NEMOTKE/nemo_tke_surface_forcing.jl)
// ---------------------------------------------------------------------------------------------
static int check_closure(const coflux_closure_forcing* f) {
  REQUIRE(std::isfinite(f->minimum_friction_velocity) && std::isfinite(f->minimum_surface_tke) && std::isfinite(f->Cb) &&
          std::isfinite(f->gravitational_acceleration), "closure forcing parameters must be finite");
  return COFLUX_OK;
}
extern "C" int coflux_attach_closure_forcing(coflux_ctx* c, const coflux_closure_forcing* f) {
  REQUIRE(c, "NULL context");
  if (!f) { c->closure_on = false; return COFLUX_OK; }
  int rc = check_closure(f);
  if (rc) return rc;
  c->closure = *f;
  c->closure_on = true;
  return COFLUX_OK;
}
template <typename FT>
static int do_closure(coflux_ctx* c, const coflux_net_ocean_fluxes* net, const coflux_closure_forcing* f, cudaStream_t st) {
  const coflux_grid_desc& g = c->cfg.grid;
  ClosureKernelArgs<FT> a;
  memset(&a, 0, sizeof(a));
  a.Nx = g.Nx; a.Ny = g.Ny;
  a.taux = view2d(net->u, 0, sizeof(FT)); a.tauy = view2d(net->v, 0, sizeof(FT));
  fill_closure<FT>(*f, net, a.closure);
  closure_forcing_kernel<FT><<<grid_for((long long)g.Nx * g.Ny, 256), 256, 0, st>>>(a);
  return check_launch(c, 1);
}
extern "C" int coflux_closure_surface_forcing(coflux_ctx* c, const coflux_net_ocean_fluxes* net, const coflux_closure_forcing* f, void* stream) {
  REQUIRE(c && net && f, "NULL argument");
  REQUIRE(net->u.ptr && net->v.ptr, "net momentum fluxes are required");
  int rc = check_closure(f);
  if (rc) return rc;
  CUDA_TRY(cudaSetDevice(c->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return c->cfg.dtype == COFLUX_F64 ? do_closure<double>(c, net, f, st) : do_closure<float>(c, net, f, st);
}

// ---------------------------------------------------------------------------------------------
// NormalizeSalinity (omip_simulation.jl:187-220)
// ---------------------------------------------------------------------------------------------
static int salt_workspace(coflux_ctx* c) {
  if (c->salt_ws) return COFLUX_OK;
  CUDA_TRY(cudaMalloc(&c->salt_ws, sizeof(double) * (2 * SALT_BLOCKS + 4)));
  CUDA_TRY(cudaMemset(c->salt_ws, 0, sizeof(double) * (2 * SALT_BLOCKS + 4)));     // [2·SALT_BLOCKS + 2 …]: grid-barrier words
  return COFLUX_OK;
}
template <typename FT>
static int do_salt_sums(coflux_ctx* c, const coflux_salinity_normalization* n, double* sums, cudaStream_t st) {
  const coflux_grid_desc& g = c->cfg.grid;
  SaltSumArgs<FT> a;
  memset(&a, 0, sizeof(a));
  a.Nx = g.Nx; a.Ny = g.Ny;
  a.flux = view2d(n->flux, 0, sizeof(FT)); a.add = view2d(n->additional, 0, sizeof(FT)); a.area = view2d(n->area, 0, sizeof(FT));
  a.mask = view2d(n->mask, 0, 1);
  a.partial = c->salt_ws;
  salt_sums_kernel<FT><<<SALT_BLOCKS, 256, 0, st>>>(a);
  if (!sums) return check_launch(c, 1);                 // single-slab form: the subtraction reduces the partials itself
  salt_sums_final_kernel<<<1, 32, 0, st>>>(c->salt_ws, SALT_BLOCKS, sums);
  return check_launch(c, 2);
}
template <typename FT>
static int do_subtract_mean(coflux_ctx* c, const coflux_salinity_normalization* n, const double* sums, cudaStream_t st) {
  const coflux_grid_desc& g = c->cfg.grid;
  SubMeanArgs<FT> a;
  a.p = static_cast<char*>(n->flux.ptr) + (int64_t)n->flux.off_k * n->flux.stride_k * (int64_t)sizeof(FT);
  a.si = n->flux.stride_i; a.sj = n->flux.stride_j;
  a.ni = g.Nx + 2 * n->flux.off_i; a.nj = g.Ny + 2 * n->flux.off_j;
  a.sums = sums ? sums : c->salt_ws;
  a.nblocks = sums ? 0 : SALT_BLOCKS;
  const long long cells = (long long)a.ni * a.nj;
  const unsigned grid = (unsigned)std::min<long long>((cells + 255) / 256, 148LL * 8);
  subtract_mean_kernel<FT><<<grid, 256, 0, st>>>(a);
  return check_launch(c, 1);
}
// one cooperative launch (sums → grid barrier → subtraction); COFLUX_NORM_FUSED=0 keeps the two-launch form
template <typename FT>
static int do_normalize_fused(coflux_ctx* c, const coflux_salinity_normalization* n, cudaStream_t st, bool* done) {
  static const bool fused = [] { const char* e = std::getenv("COFLUX_NORM_FUSED"); return !(e && e[0] == '0'); }();
  *done = false;
  if (!fused) return COFLUX_OK;
  int per_sm = 0, sms = 0;
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, normalize_salinity_kernel<FT>, 256, 0));
  CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
  int grid = SALT_BLOCKS;                         // largest divisor of SALT_BLOCKS that is co-resident
  while (grid > per_sm * sms || SALT_BLOCKS % grid) --grid;
  if (grid < sms) return COFLUX_OK;               // not this device: two launches
  const coflux_grid_desc& g = c->cfg.grid;
  NormalizeArgs<FT> a;
  memset(&a, 0, sizeof(a));
  a.s.Nx = g.Nx; a.s.Ny = g.Ny;
  a.s.flux = view2d(n->flux, 0, sizeof(FT)); a.s.add = view2d(n->additional, 0, sizeof(FT)); a.s.area = view2d(n->area, 0, sizeof(FT));
  a.s.mask = view2d(n->mask, 0, 1);
  a.s.partial = c->salt_ws;
  a.m.p = static_cast<char*>(n->flux.ptr) + (int64_t)n->flux.off_k * n->flux.stride_k * (int64_t)sizeof(FT);
  a.m.si = n->flux.stride_i; a.m.sj = n->flux.stride_j;
  a.m.ni = g.Nx + 2 * n->flux.off_i; a.m.nj = g.Ny + 2 * n->flux.off_j;
  a.vblocks = SALT_BLOCKS;
  a.bar = reinterpret_cast<unsigned*>(c->salt_ws + 2 * SALT_BLOCKS + 2);
  void* params[] = {&a};
  CUDA_TRY(cudaLaunchCooperativeKernel((void*)normalize_salinity_kernel<FT>, dim3(grid), dim3(256), params, 0, st));
  *done = true;
  return check_launch(c, 1);
}
static int check_norm(coflux_ctx* c, const coflux_salinity_normalization* n) {
  REQUIRE(c && n, "NULL argument");
  REQUIRE(n->flux.ptr && n->area.ptr, "salinity normalization needs the flux field and the cell areas");
  REQUIRE(n->flux.off_i >= 0 && n->flux.off_j >= 0, "negative halo offsets");
  return COFLUX_OK;
}
extern "C" int coflux_salinity_flux_sums(coflux_ctx* c, const coflux_salinity_normalization* n, double* device_sums, void* stream) {
  int rc = check_norm(c, n);
  if (rc) return rc;
  REQUIRE(device_sums, "device_sums is NULL");
  CUDA_TRY(cudaSetDevice(c->device));
  rc = salt_workspace(c);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return c->cfg.dtype == COFLUX_F64 ? do_salt_sums<double>(c, n, device_sums, st) : do_salt_sums<float>(c, n, device_sums, st);
}
extern "C" int coflux_subtract_mean_flux(coflux_ctx* c, const coflux_salinity_normalization* n, const double* device_sums, void* stream) {
  int rc = check_norm(c, n);
  if (rc) return rc;
  REQUIRE(device_sums, "device_sums is NULL");
  CUDA_TRY(cudaSetDevice(c->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return c->cfg.dtype == COFLUX_F64 ? do_subtract_mean<double>(c, n, device_sums, st) : do_subtract_mean<float>(c, n, device_sums, st);
}
extern "C" int coflux_normalize_salinity_flux(coflux_ctx* c, const coflux_salinity_normalization* n, void* stream) {
  int rc = check_norm(c, n);
  if (rc) return rc;
  CUDA_TRY(cudaSetDevice(c->device));
  rc = salt_workspace(c);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  bool done = false;
  rc = c->cfg.dtype == COFLUX_F64 ? do_normalize_fused<double>(c, n, st, &done) : do_normalize_fused<float>(c, n, st, &done);
  if (rc || done) return rc;
  rc = c->cfg.dtype == COFLUX_F64 ? do_salt_sums<double>(c, n, nullptr, st) : do_salt_sums<float>(c, n, nullptr, st);
  if (rc) return rc;
  return c->cfg.dtype == COFLUX_F64 ? do_subtract_mean<double>(c, n, nullptr, st) : do_subtract_mean<float>(c, n, nullptr, st);
}

// ---------------------------------------------------------------------------------------------
// multi-GPU seam, mode B: fused compute + exchange over NVLink peer memory.
// The flux kernel stores the last interior column of ρτx straight into the east neighbour's seam
// buffer (peer store), a stream write-value publishes it, the neighbour's stream wait-value orders
// its stress kernel behind it.  No host round trip, no separate message, no NCCL call per step.
// ---------------------------------------------------------------------------------------------
static int load_stream_memops() {
  if (g_write32 && g_wait32) return COFLUX_OK;
  cudaDriverEntryPointQueryResult qr;
  void* fw = nullptr; void* fq = nullptr;
  CUDA_TRY(cudaGetDriverEntryPoint("cuStreamWriteValue32", &fw, cudaEnableDefault, &qr));
  CUDA_TRY(cudaGetDriverEntryPoint("cuStreamWaitValue32", &fq, cudaEnableDefault, &qr));
  if (!fw || !fq) return fail(COFLUX_ERR_SEAM, "driver does not export cuStreamWriteValue32 / cuStreamWaitValue32");
  g_write32 = reinterpret_cast<cuStreamMemOp32_t>(fw);
  g_wait32 = reinterpret_cast<cuStreamMemOp32_t>(fq);
  return COFLUX_OK;
}
static int ensure_seam_buffer(coflux_ctx* c) {
  if (c->seam.local) return COFLUX_OK;
  const size_t es = (c->cfg.dtype == COFLUX_F64) ? 8 : 4;
  c->seam.bytes = SEAM_DATA_OFFSET + 2 * (size_t)c->cfg.grid.Ny * es;
  CUDA_TRY(cudaMalloc(&c->seam.local, c->seam.bytes));
  CUDA_TRY(cudaMemset(c->seam.local, 0, c->seam.bytes));
  CUDA_TRY(cudaDeviceSynchronize());
  return COFLUX_OK;
}
struct SeamHandleWire {          // what travels between ranks (≤ COFLUX_SEAM_HANDLE_BYTES)
  cudaIpcMemHandle_t ipc;        // 64 bytes
  int32_t Ny, dtype, device, pid;
  uint64_t local_ptr;            // for the same-process case (world == 1 or several contexts per process)
};
static_assert(sizeof(SeamHandleWire) <= COFLUX_SEAM_HANDLE_BYTES, "seam handle too large");

extern "C" int coflux_seam_export(coflux_ctx* c, void* handle_out) {
  REQUIRE(c && handle_out, "NULL argument");
  CUDA_TRY(cudaSetDevice(c->device));
  int rc = ensure_seam_buffer(c);
  if (rc) return rc;
  SeamHandleWire w;
  memset(&w, 0, sizeof(w));
  CUDA_TRY(cudaIpcGetMemHandle(&w.ipc, c->seam.local));
  w.Ny = c->cfg.grid.Ny; w.dtype = c->cfg.dtype; w.device = c->device; w.pid = (int32_t)getpid();
  w.local_ptr = (uint64_t)(uintptr_t)c->seam.local;
  memset(handle_out, 0, COFLUX_SEAM_HANDLE_BYTES);
  memcpy(handle_out, &w, sizeof(w));
  return COFLUX_OK;
}
static int open_peer(coflux_ctx* c, const SeamHandleWire& w, char** out, bool* is_ipc) {
  if (w.Ny != c->cfg.grid.Ny || w.dtype != c->cfg.dtype)
    return fail(COFLUX_ERR_SEAM, "neighbour slab has Ny=%d dtype=%d, this slab Ny=%d dtype=%d", w.Ny, w.dtype, c->cfg.grid.Ny, c->cfg.dtype);
  if (w.pid == (int32_t)getpid()) {            // same process: the pointer is directly usable (peer access enabled below)
    *out = reinterpret_cast<char*>((uintptr_t)w.local_ptr);
    *is_ipc = false;
    if (w.device != c->device) {
      int can = 0;
      CUDA_TRY(cudaDeviceCanAccessPeer(&can, c->device, w.device));
      if (!can) return fail(COFLUX_ERR_SEAM, "device %d cannot access peer device %d", c->device, w.device);
      cudaError_t e = cudaDeviceEnablePeerAccess(w.device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(COFLUX_ERR_CUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
      cudaGetLastError();
    }
    return COFLUX_OK;
  }
  void* p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, w.ipc, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) return fail(COFLUX_ERR_SEAM, "cudaIpcOpenMemHandle failed: %s", cudaGetErrorString(e));
  *out = static_cast<char*>(p);
  *is_ipc = true;
  return COFLUX_OK;
}
extern "C" int coflux_seam_attach(coflux_ctx* c, const void* west, const void* east, int32_t rank, int32_t world) {
  REQUIRE(c && west && east, "NULL argument");
  REQUIRE(world >= 1 && rank >= 0 && rank < world, "bad rank/world");
  REQUIRE(c->cfg.grid.ring == 0, "seam mode needs grid.ring == 0 (ring == 1 is the zero-message mode)");
  CUDA_TRY(cudaSetDevice(c->device));
  int rc = load_stream_memops();
  if (rc) return rc;
  rc = ensure_seam_buffer(c);
  if (rc) return rc;
  coflux_seam_detach(c);
  SeamHandleWire ww, we;
  memcpy(&ww, west, sizeof(ww));
  memcpy(&we, east, sizeof(we));
  Seam& s = c->seam;
  rc = open_peer(c, we, &s.east, &s.east_is_ipc);
  if (rc) return rc;
  if (memcmp(&ww, &we, sizeof(ww)) == 0) { s.west = s.east; s.west_is_ipc = false; }   // two slabs: the same neighbour on both sides
  else { rc = open_peer(c, ww, &s.west, &s.west_is_ipc); if (rc) return rc; }
  // The flag / ack words are NOT reset here: the buffer was zeroed before it was exported (ensure_seam_buffer), and a
  // neighbour that attached earlier may already have published its first column into it.  The step counter is not
  // reset either (flags are monotone step ids), so a detach / re-attach of the same ring continues the sequence.
  s.rank = rank; s.world = world; s.attached = true;
  return COFLUX_OK;
}
extern "C" int coflux_seam_detach(coflux_ctx* c) {
  if (!c) return COFLUX_OK;
  Seam& s = c->seam;
  if (s.attached) {
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    if (s.west_is_ipc && s.west && s.west != s.east) cudaIpcCloseMemHandle(s.west);
    if (s.east_is_ipc && s.east) cudaIpcCloseMemHandle(s.east);
  }
  s.east = s.west = nullptr; s.east_is_ipc = s.west_is_ipc = false; s.attached = false;
  return COFLUX_OK;
}

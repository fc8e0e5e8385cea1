/* coflux.h — C ABI of the Blackwell-native surface-flux engine.
 *
 * This is the drop-in boundary for ONE hot path of CliMA/ClimaOcean.jl: the per-coupling-step
 * surface-flux computation behind `update_state!(::OceanSeaIceModel)`.  The reference has no FFI
 * on this path (it is Julia multiple dispatch); each entry point below names the
 * reference-side generic function whose method a maintainer would override with a `ccall`
 * (see INTEGRATION.md and julia/CoFluxExt/).  Reference citations are relative to /root/reference.
 *
 *   coflux_interpolate_atmosphere     <- interpolate_atmosphere_state!      (SURVEY §8 row a3;
 *                                        inputs src/OMIPConfigurations/atmosphere.jl:22-46,
 *                                        jra55_data_staging.jl:8)
 *   coflux_atmosphere_ocean_fluxes    <- compute_atmosphere_ocean_fluxes!   (rows a4-a6; parameter
 *                                        surface src/OMIPConfigurations/omip_simulation.jl:40-113,
 *                                        outputs omip_diagnostics.jl:81-82)
 *   coflux_atmosphere_sea_ice_fluxes  <- compute_atmosphere_sea_ice_fluxes! (row a7;
 *                                        omip_simulation.jl:52-69,91-113)
 *   coflux_sea_ice_ocean_fluxes       <- compute_sea_ice_ocean_fluxes!      (row a8;
 *                                        omip_simulation.jl:71-77, omip_diagnostics.jl:84-89)
 *   coflux_assemble_net_ocean_fluxes  <- compute_net_ocean_fluxes!          (row a9;
 *                                        omip_diagnostics.jl:77-80, visualize/cache.jl:359-386)
 *   coflux_update_state               <- update_state!(::OceanSeaIceModel)  (row a2; fused path)
 *   coflux_update_state_host          <- same, HOST buffers (end-to-end measurement entry)
 *   coflux_create / coflux_destroy    <- ComponentInterfaces(...) construction (row a10;
 *                                        omip_simulation.jl:128-158)
 *   coflux_interpolate_land           <- the land / auxiliary-freshwater part of interpolate_atmosphere_state!
 *                                        (JRA55PrescribedLand: friver + licalvf, atmosphere.jl:46, jra55_data_staging.jl:8)
 *   coflux_assemble_net_sea_ice_fluxes <- compute_net_sea_ice_fluxes!        
 *   coflux_forcing_window_*           <- FieldTimeSeries InMemory window with prefetch (time_indices_in_memory,
 *                                        prefetch = true: atmosphere.jl:22-27; §8f row 2)
 *   coflux_attach_flux_averages       <- AveragedTimeInterval output of the flux fields (omip_diagnostics.jl:77-89,
 *                                        125-158; §8f row 4)
 *
 * Conventions
 *   - every function returns 0 (COFLUX_OK) or a negative coflux_status; nothing throws or aborts;
 *     coflux_last_error() returns a thread-local, library-owned message.
 *   - ownership: every data array belongs to the caller (Julia GC / CUDA.jl pool).  The library
 *     allocates only its context workspace.  Array descriptors are never retained past the call.
 *   - asynchrony: all device work is enqueued on the caller's CUstream (`cu_stream`, may be NULL
 *     for the legacy default stream) and the call returns immediately; no implicit sync.
 *   - layout: Oceananigans "parent" arrays — halo-padded, column-major (i fastest).  A descriptor
 *     carries the base pointer of the parent, ELEMENT strides, and the halo offsets, so that the
 *     zero-based interior cell (i,j,k) lives at ptr[(i+off_i)*stride_i + (j+off_j)*stride_j +
 *     (k+off_k)*stride_k + n*stride_n].  Nothing is assumed compact.  Negative interior indices
 *     (halo cells) are legal wherever off_* allows.  (Fast path: when all 2-D surface arrays of a call have
 *     stride_i == 1 and one common stride_j — Oceananigans parents of one grid do — and the atmosphere series
 *     likewise among themselves, the kernels share one element offset per cell; any other layout is served by
 *     a kernel that addresses every array through its own strides, with the same results.)
 *   - element type of all floating-point arrays = the context dtype (COFLUX_F32 / COFLUX_F64).
 *   - no CPU fallback exists anywhere in this library.
 */
#ifndef COFLUX_H
#define COFLUX_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define COFLUX_ABI_VERSION 2   /* 2: land freshwater series, forcing ring, CCSM3 sea-ice albedo, net sea-ice fluxes, flux averages */

typedef enum {
  COFLUX_OK = 0,
  COFLUX_ERR_INVALID_ARGUMENT = -1,  /* NULL where required, bad enum, non-finite parameter ...   */
  COFLUX_ERR_CUDA = -2,              /* a CUDA runtime call failed; message has the CUDA string     */
  COFLUX_ERR_UNSUPPORTED = -3,       /* valid but not implemented combination                        */
  COFLUX_ERR_NO_DEVICE = -4,         /* no usable CUDA device (the library never falls back to CPU)  */
  COFLUX_ERR_ALLOC = -5,
  COFLUX_ERR_SEAM = -6               /* multi-GPU seam not attached / peer access unavailable        */
} coflux_status;

typedef enum { COFLUX_F32 = 32, COFLUX_F64 = 64 } coflux_dtype;

/* ------------------------------------------------------------------------------------------------
 * Array descriptor (device memory unless stated otherwise).  ptr == NULL means "field absent".
 * ---------------------------------------------------------------------------------------------- */
typedef struct coflux_array {
  void*   ptr;
  int64_t stride_i, stride_j, stride_k, stride_n;  /* element strides (n = time index of a series) */
  int32_t off_i, off_j, off_k;                     /* halo offsets                                  */
  int32_t reserved;
} coflux_array;

/* ------------------------------------------------------------------------------------------------
 * Parameter PODs — mirrors of the Julia parameter structs (omip_simulation.jl:40-113,
 * OMIPConfigurations.jl:15-33).  All scalars are double in the ABI and converted once to the
 * context dtype.  coflux_default_*() fill the reference's defaults.
 * ---------------------------------------------------------------------------------------------- */
typedef enum { COFLUX_VISCOSITY_CONSTANT = 0, COFLUX_VISCOSITY_TEMPERATURE_POLY = 1 } coflux_viscosity_kind;
typedef struct coflux_air_viscosity {     /* TemperatureDependentAirViscosity (omip_simulation.jl:41) */
  int32_t kind, reserved;
  double  nu;                             /* CONSTANT: ν [m²/s]                                        */
  double  c0, c1, c2, c3;                 /* POLY: ν = c0 + c1 T' + c2 T'^2 + c3 T'^3, T' = T - 273.15 */
} coflux_air_viscosity;

typedef enum { COFLUX_ROUGHNESS_FIXED = 0,           /* plain number (omip_simulation.jl:67-69)      */
               COFLUX_ROUGHNESS_CHARNOCK = 1,        /* MomentumRoughnessLength                       */
               COFLUX_ROUGHNESS_REYNOLDS_SCALING = 2 /* ScalarRoughnessLength                         */
} coflux_roughness_kind;
typedef enum { COFLUX_WAVES_CONSTANT = 0,            /* constant Charnock parameter                   */
               COFLUX_WAVES_WIND_DEPENDENT = 1       /* WindDependentWaveFormulation (Edson 2013 e13) */
} coflux_wave_formulation;

typedef struct coflux_momentum_roughness {
  int32_t kind;                    /* FIXED | CHARNOCK                                               */
  int32_t wave_formulation;        /* CONSTANT | WIND_DEPENDENT                                      */
  double  fixed_length;            /* FIXED: ℓu [m]                                                  */
  double  gravity_wave_parameter;  /* Charnock α (constant formulation)                              */
  double  wind_a1, wind_a2;        /* α = a1·min(U,Umax) + a2, floored at wind_alpha_min             */
  double  wind_umax, wind_alpha_min;
  double  smooth_wall_parameter;   /* β_s in β_s ν/u★ (0.11)                                         */
  double  maximum_length;          /* ℓ_max                                                          */
  double  gravitational_acceleration;
  coflux_air_viscosity viscosity;
} coflux_momentum_roughness;

typedef struct coflux_scalar_roughness {
  int32_t kind, reserved;          /* FIXED | REYNOLDS_SCALING                                       */
  double  fixed_length;
  double  reynolds_A, reynolds_b;  /* ℓ = A · R★^(-b)  (5.85e-5, 0.72)                               */
  double  maximum_length;          /* 1.6e-4                                                         */
  coflux_air_viscosity viscosity;
} coflux_scalar_roughness;

typedef enum { COFLUX_FLUXES_SIMILARITY_THEORY = 0,  /* SimilarityTheoryFluxes                        */
               COFLUX_FLUXES_COEFFICIENT_LARGE_YEAGER = 1 /* CoefficientBasedFluxes + LargeYeager     */
} coflux_flux_formulation;
typedef enum { COFLUX_STABILITY_EDSON = 0,           /* atmosphere_ocean_stability_functions          */
               COFLUX_STABILITY_SHEBA_PAULSON = 1,   /* atmosphere_sea_ice_stability_functions        */
               COFLUX_STABILITY_LARGE_YEAGER = 2,    /* large_yeager_stability_functions              */
               COFLUX_STABILITY_NEUTRAL = 3          /* ψ ≡ 0 (known-answer tests)                    */
} coflux_stability_functions;
typedef enum { COFLUX_PROFILE_LOGARITHMIC = 0,       /* ln(h/ℓ) − ψ(h/L) + ψ(ℓ/L)                     */
               COFLUX_PROFILE_COARE_LOGARITHMIC = 1  /* COARELogarithmicSimilarityProfile             */
} coflux_similarity_form;
typedef enum { COFLUX_VELOCITY_RELATIVE = 0, COFLUX_VELOCITY_WIND = 1 } coflux_velocity_formulation;
typedef enum { COFLUX_STOP_CONVERGENCE = 0, COFLUX_STOP_FIXED_ITERATIONS = 1 } coflux_stop_kind;
typedef enum { COFLUX_TEMPERATURE_BULK = 0, COFLUX_TEMPERATURE_SKIN = 1 } coflux_interface_temperature;
/* How the SKIN temperature is advanced inside the similarity iteration (row a7).  With Q_a = Q_d + Q_u(T_s) + Q_c + Q_v the
 * conductive balance through the slab is  T★ = T_b − Q_a h / k.
 *   CLAMPED_EXPLICIT      T_s ← min(T_s + clamp(T★(T_s⁻) − T_s⁻, ±ΔT_max), T_melt): every flux at the previous iterate — the form the
 *                         reference's SkinTemperature runs.  Its slope (h/k)·∂Q_a/∂T_s exceeds one for h ≳ 0.1 m: not a
 *                         contraction; one ice cell in nine ends on a limit cycle at maxiter (tests/test_a7_conditioning.py).
 *   LINEARIZED_LONGWAVE   the emitted long wave is taken implicitly, Q_u ≈ σ ε T_s⁻³ · T_s⁺:
 *                         T★ = (T_b − (Q_d + Q_c + Q_v) h / k) / (1 + σ ε T_s⁻³ h / k), then the same clamp and melting cap.
 *                         (The alternative upstream keeps commented out next to the explicit form.)  Removes the radiative
 *                         part (4σεT³ ≈ 5 W m⁻² K⁻¹) of the slope.  Measured (tests/test_a7_conditioning.py): same fixed points,
 *                         but the limit cycles barely recede (940 → 937 of 10 571 ice cells) — they are driven by the
 *                         turbulent fluxes' own dependence on T_s through Δθ and Δq (ρ c_p C_h U ≈ 15 W m⁻² K⁻¹ and the
 *                         latent analogue), which both forms lag.  Offered as a parameter; NOT the reference default.           */
typedef enum { COFLUX_SKIN_CLAMPED_EXPLICIT = 0, COFLUX_SKIN_LINEARIZED_LONGWAVE = 1 } coflux_skin_temperature_update;

typedef struct coflux_flux_params {
  int32_t formulation;              /* coflux_flux_formulation                                       */
  int32_t stability_functions;      /* coflux_stability_functions                                    */
  int32_t similarity_form;          /* coflux_similarity_form                                        */
  int32_t velocity_formulation;     /* coflux_velocity_formulation (omip_simulation.jl:135-137)      */
  int32_t stop_kind;                /* coflux_stop_kind                                              */
  int32_t max_iterations;           /* maxiter (CONVERGENCE) or n (FixedIterations(n))               */
  int32_t interface_temperature;    /* BULK (ocean) | SKIN (sea ice, row a7)                         */
  int32_t skin_temperature_update;  /* coflux_skin_temperature_update (SKIN only)                    */
  double  tolerance;                /* Σ|Δ(u★,θ★,q★)| < tolerance                                   */
  double  von_karman_constant;
  double  turbulent_prandtl_number; /* must be 1 (kept for struct parity)                            */
  double  gustiness_parameter;      /* β                                                             */
  double  minimum_gustiness;        /* floor on Uᴳ (omip_simulation.jl:44,66,110)                    */
  double  initial_scale;            /* u★=θ★=q★ initial guess (1e-4)                                 */
  double  ly_minimum_wind;          /* Large–Yeager wind floor (0.5)                                 */
  double  skin_max_delta_T;         /* SKIN: cap on |ΔT_s| per iteration                             */
  coflux_momentum_roughness momentum_roughness;
  coflux_scalar_roughness   temperature_roughness;
  coflux_scalar_roughness   water_vapor_roughness;
} coflux_flux_params;

typedef struct coflux_thermodynamics {   /* AtmosphereThermodynamicsParameters                        */
  double gas_constant, dry_air_molar_mass, water_molar_mass;
  double dry_air_adiabatic_exponent;     /* κ_d = 2/7                                                 */
  double water_vapor_heat_capacity, liquid_water_heat_capacity, ice_heat_capacity;
  double reference_vaporization_enthalpy, reference_sublimation_enthalpy;
  double reference_temperature, triple_point_temperature, triple_point_pressure;
  double water_freezing_temperature, total_ice_nucleation_temperature;
} coflux_thermodynamics;

typedef struct coflux_atmosphere_properties {
  coflux_thermodynamics thermodynamics;
  double surface_layer_height;       /* h   = 10 m                                                    */
  double boundary_layer_height;      /* h_bℓ = 512 m                                                  */
  double gravitational_acceleration; /* g                                                             */
} coflux_atmosphere_properties;

typedef enum { COFLUX_TEMPERATURE_CELSIUS = 0, COFLUX_TEMPERATURE_KELVIN = 1 } coflux_temperature_units;

typedef struct coflux_ocean_properties {
  double  reference_density;     /* ρ₀  (1026;  visualize/common.jl:17)                               */
  double  heat_capacity;         /* c₀  (3991.86795711963; visualize/common.jl:18)                    */
  double  freshwater_density;    /* ρ_f (1000)                                                        */
  double  minimum_salinity;      /* ocean_minimum_salinity (omip_simulation.jl:125; launch.sh:74-78)  */
  int32_t temperature_units;     /* coflux_temperature_units                                          */
  int32_t reserved;
  /* Raoult water-mole-fraction constituents (chloride, sodium, sulfate, magnesium)                  */
  double  salt_water_molar_mass;               /* 18.02                                               */
  double  constituent_molar_mass[4];
  double  constituent_mass_fraction[4];
} coflux_ocean_properties;

typedef enum { COFLUX_SEA_ICE_ALBEDO_PRESCRIBED = 0,  /* coflux_sea_ice_state.albedo plane, else the constant sea_ice_albedo */
               COFLUX_SEA_ICE_ALBEDO_CCSM3 = 1        /* SeaIceAlbedo(h_i, h_s, T_s) from the LIVE sea-ice fields (atmosphere.jl:31-44) */
} coflux_sea_ice_albedo_kind;
/* CCSM3 sea-ice albedo (Briegleb et al. 2004, the "ccsm3" shortwave option of CICE): thickness, snow-depth and surface
 * temperature dependent, two spectral bands combined with a fixed visible fraction.
 *   f_h = min(atan(4 h_i) / atan(4 h_max), 1)             thin ice fades into the ocean albedo
 *   f_T = clamp(1 − (T_melt − T_s)/ΔT_melt, 0, 1)          0 below T_melt − ΔT_melt, 1 at the melting point
 *   α_ice,b  = α_ice,b⁰ f_h + α_ocean (1 − f_h) − Δα_ice f_T          (b = visible, near infrared)
 *   α_snow,b = α_snow,b⁰ − Δα_snow,b f_T
 *   f_s = h_s / (h_s + h_patch);   α_b = (1 − f_s) α_ice,b + f_s α_snow,b;   α = f_vis α_vis + (1 − f_vis) α_nir            */
typedef struct coflux_ccsm3_albedo {
  double ice_visible, ice_near_infrared;            /* 0.78, 0.36                                            */
  double snow_visible, snow_near_infrared;          /* 0.98, 0.70                                            */
  double thickness_scale;                           /* h_max = 0.3 m                                         */
  double melt_temperature_range;                    /* ΔT_melt = 1.5 K                                       */
  double ice_melt_change;                           /* Δα_ice = 0.075                                        */
  double snow_visible_melt_change, snow_near_infrared_melt_change;   /* 0.10, 0.15                          */
  double snow_patchiness;                           /* h_patch = 0.02 m                                      */
  double ocean_albedo;                              /* 0.06                                                  */
  double visible_fraction;                          /* f_vis = 0.52                                          */
  double melting_temperature;                       /* T_melt = 273.15 K                                     */
} coflux_ccsm3_albedo;

typedef struct coflux_radiation_properties {  /* SurfaceRadiationProperties (atmosphere.jl:42-46)     */
  double stefan_boltzmann_constant;
  double ocean_albedo, ocean_emissivity;
  double sea_ice_emissivity;
  double sea_ice_albedo;         /* PRESCRIBED kind: used when coflux_sea_ice_state.albedo is absent  */
  int32_t shortwave_penetrates;  /* 1: transmitted SW goes to the penetrating-radiation surface flux */
  int32_t sea_ice_albedo_kind;   /* coflux_sea_ice_albedo_kind                                       */
  coflux_ccsm3_albedo ccsm3;
} coflux_radiation_properties;

typedef enum { COFLUX_ICE_OCEAN_ICE_BATH = 0,        /* bulk: ρ₀c₀ u_m★ (T − T_m) ℵ                   */
               COFLUX_ICE_OCEAN_THREE_EQUATION = 1   /* ThreeEquationHeatFlux (omip_simulation.jl:77) */
} coflux_ice_ocean_heat_flux;
typedef enum { COFLUX_FRICTION_VELOCITY_CONSTANT = 0,
               COFLUX_FRICTION_VELOCITY_MOMENTUM_BASED = 1 /* MomentumBasedFrictionVelocity           */
} coflux_friction_velocity;

typedef struct coflux_ice_ocean_params {
  int32_t heat_flux;                     /* coflux_ice_ocean_heat_flux                                */
  int32_t friction_velocity;             /* coflux_friction_velocity                                  */
  double  characteristic_melting_speed;  /* u_m★ (ICE_BATH)                                           */
  double  liquidus_freshwater_melting_temperature;  /* T₀ in ocean temperature units                  */
  double  liquidus_slope;                /* T_m(S) = T₀ − slope·S                                     */
  double  heat_transfer_coefficient;     /* α_h : γ_T = α_h u★                                        */
  double  salt_transfer_coefficient;     /* α_s : γ_S = α_s u★                                        */
  double  constant_friction_velocity;    /* u★ when CONSTANT                                          */
  double  minimum_friction_velocity;     /* floor for MOMENTUM_BASED                                  */
  double  ice_density, ice_latent_heat;  /* ρ_i, ℒ_f                                                  */
  double  ice_ocean_drag_coefficient;    /* Cᴰ of the quadratic ice–ocean stress                      */
  double  ice_conductivity;              /* k of the conductive flux (SKIN temperature, row a7)       */
  double  ice_consolidation_thickness;   /* h_c                                                       */
} coflux_ice_ocean_params;

typedef struct coflux_grid_desc {
  int32_t Nx, Ny, Nz;      /* interior size of the (local) ocean grid                                */
  int32_t ring;            /* surface kernels run over i∈[-ring, Nx+ring), j∈[-ring, Ny+ring)         */
                           /* (the reference computes into one halo ring so that the centre→face     */
                           /* stress averaging needs no halo exchange)                  */
  int32_t periodic_x;      /* 1: when ring==0, stress averaging wraps i=-1 → Nx-1                     */
  int32_t reserved;
} coflux_grid_desc;

typedef struct coflux_config {
  int32_t abi_version;     /* must be COFLUX_ABI_VERSION                                              */
  int32_t dtype;           /* coflux_dtype                                                            */
  int32_t device;          /* CUDA device ordinal                                                     */
  int32_t reserved;
  coflux_grid_desc             grid;
  coflux_flux_params           atmosphere_ocean;
  coflux_flux_params           atmosphere_sea_ice;
  coflux_ice_ocean_params      ice_ocean;
  coflux_atmosphere_properties atmosphere;
  coflux_ocean_properties      ocean;
  coflux_radiation_properties  radiation;
} coflux_config;

/* Fill every field of cfg with the reference defaults: Edson ψ, constant
 * Charnock, convergence 1e-8 / 100 iterations, relative velocity, ice-bath heat flux ...        */
int coflux_default_config(coflux_config* cfg, int32_t Nx, int32_t Ny, int32_t Nz, int32_t dtype);
/* Presets named after build_coupled_model's flux_configuration (omip_simulation.jl:127-160):
 * "default", "corrected", "ncar".  velocity: coflux_velocity_formulation.                         */
int coflux_apply_flux_configuration(coflux_config* cfg, const char* name, int32_t velocity);

/* ------------------------------------------------------------------------------------------------
 * Data bundles
 * ---------------------------------------------------------------------------------------------- */
typedef enum { COFLUX_TIME_LINEAR = 0, COFLUX_TIME_CYCLICAL = 1, COFLUX_TIME_CLAMP = 2 } coflux_time_indexing;

/* PrescribedAtmosphere / PrescribedRadiation FieldTimeSeries windows currently in memory.
 * Each array is (Nλ+2H, Nφ+2H, 1, Nt) with stride_n between time levels.  `times` is a HOST
 * pointer to Nt doubles (seconds).  Source-grid halos in λ must be filled (periodic).           */
typedef struct coflux_atmos_series {
  coflux_array u, v, T, q, p;        /* uas vas tas huss psl                                         */
  coflux_array Qs, Ql;               /* rsds rlds                                                    */
  coflux_array rain, snow;           /* prra prsn (summed AFTER interpolation into Mp)               */
  const double* times;               /* host, length Nt                                              */
  int32_t Nt;
  int32_t time_indexing;             /* coflux_time_indexing                                         */
  double  cycle_period;              /* CYCLICAL: period; <=0 → times[Nt-1]-times[0]+Δt              */
  /* fractional zero-based source indices of every ocean cell of the ring-extended surface:        */
  coflux_array fi, fj;               /* (Nx+2ring.., Ny+2ring..) dtype; see coflux_grid_desc.ring    */
  /* optional rotation of (u,v) into the grid frame (curvilinear grids): u' = c·u + s·v, v' = −s·u + c·v */
  coflux_array cos_theta, sin_theta;
  /* device ring buffer (coflux_forcing_window): logical level n of `times` is stored at time slot
   * (ring_start + n) mod ring_capacity of every series array.  ring_capacity == 0: plain layout (slot n).    */
  int32_t ring_start, ring_capacity;
} coflux_atmos_series;

/* PrescribedLand — JRA55PrescribedLand (atmosphere.jl:46): river runoff `friver` and iceberg calving `licalvf`
 * (jra55_data_staging.jl:8), freshwater mass fluxes [kg m⁻² s⁻¹] on their OWN source grid and time axis.  They are
 * interpolated like the atmosphere (bilinear × linear in time) and ADDED to the exchange freshwater flux Mp, so that they
 * enter the net salinity flux like rain and snow.                                                              */
typedef struct coflux_land_series {
  coflux_array rivers, icebergs;     /* friver, licalvf; either may be absent                                */
  const double* times;               /* host, length Nt                                                      */
  int32_t Nt;
  int32_t time_indexing;             /* coflux_time_indexing                                                 */
  double  cycle_period;
  coflux_array fi, fj;               /* fractional source indices of every ocean cell (ring-extended surface) */
  int32_t ring_start, ring_capacity; /* as in coflux_atmos_series                                            */
} coflux_land_series;

/* The 2-D "exchange" atmosphere state on the ocean grid (outputs of a3, inputs of a4/a9).         */
typedef struct coflux_exchange_state {
  coflux_array u, v, T, p, q, Qs, Ql, Mp;
} coflux_exchange_state;

/* Ocean surface: 3-D parents of which only the k = Nz-1 plane is read; u at (Face,Center),
 * v at (Center,Face).  mask: optional uint8 2-D array, 1 = active (wet) surface cell.             */
typedef struct coflux_ocean_surface {
  coflux_array u, v, T, S;
  coflux_array mask;                 /* uint8 elements; strides in bytes-as-elements                 */
} coflux_ocean_surface;

/* interfaces.atmosphere_ocean_interface.fluxes.* (omip_diagnostics.jl:81-82) + interface T.       */
typedef struct coflux_interface_fluxes {
  coflux_array latent_heat, sensible_heat, water_vapor, x_momentum, y_momentum;
  coflux_array interface_temperature;
  coflux_array friction_velocity, temperature_scale, humidity_scale;  /* optional u★ θ★ q★          */
  coflux_array iterations;           /* optional diagnostic: int32 iteration count per cell         */
} coflux_interface_fluxes;

/* Sea-ice state seen by the flux path.                                                             */
typedef struct coflux_sea_ice_state {
  coflux_array thickness, previous_thickness;   /* h, h⁻ (h⁻ is UPDATED by the ice–ocean kernel)     */
  coflux_array concentration;                   /* ℵ                                                 */
  coflux_array salinity;                        /* S_i                                               */
  coflux_array u, v;                            /* ice velocities at (F,C) / (C,F)                   */
  coflux_array top_temperature;                 /* T_top (in/out of the skin-temperature solve)      */
  coflux_array snow_thickness;                  /* optional                                          */
  coflux_array albedo;                          /* optional 2-D albedo (PRESCRIBED kind); else radiation default */
} coflux_sea_ice_state;

/* Ocean columns for the frazil sweep: T is READ AND CONDITIONALLY WRITTEN.                         */
typedef struct coflux_ocean_columns {
  coflux_array T, S;
  coflux_array dz;                   /* Δz: 1-D in k (stride_k only) or full 3-D                     */
  coflux_array u, v;                 /* for the ice–ocean stress (k = Nz-1 plane)                    */
} coflux_ocean_columns;

/* interfaces.sea_ice_ocean_interface.fluxes.* (omip_diagnostics.jl:84-89)                          */
typedef struct coflux_ice_ocean_fluxes {
  coflux_array frazil_heat, interface_heat, salt, x_momentum, y_momentum;
} coflux_ice_ocean_fluxes;

/* interfaces.net_fluxes.ocean.{u,v,T,S} (= the ocean's top boundary-condition arrays,
 * omip_diagnostics.jl:77-80) + radiative diagnostics + penetrating shortwave surface flux.         */
typedef struct coflux_net_ocean_fluxes {
  coflux_array u, v, T, S;
  coflux_array upwelling_longwave, downwelling_longwave, downwelling_shortwave;
  coflux_array penetrating_shortwave;
} coflux_net_ocean_fluxes;

/* interfaces.net_fluxes.sea_ice.{top, bottom} — what compute_net_sea_ice_fluxes! hands to the sea-ice model
 * (SURVEY §3.2): the heat flux into the ice from above (radiation + turbulent, where ice is present) and from below
 * (frazil + interface heat), W m⁻², positive upward; optionally the atmosphere–ice stress moved to the velocity
 * points of the ice model (N m⁻²).                                                                              */
typedef struct coflux_net_sea_ice_fluxes {
  coflux_array top_heat, bottom_heat;
  coflux_array top_u, top_v;                     /* optional: ρτx at (Face,Center), ρτy at (Center,Face)          */
} coflux_net_sea_ice_fluxes;

typedef struct coflux_update_inputs {
  const coflux_atmos_series*  atmosphere;
  const coflux_ocean_surface* ocean;
  const coflux_sea_ice_state* sea_ice;           /* NULL: ocean-only model                           */
  const coflux_ice_ocean_fluxes* ice_ocean;      /* NULL: no ice–ocean contribution in the assembly  */
  const coflux_land_series*   land;              /* NULL: no land freshwater                         */
} coflux_update_inputs;

typedef struct coflux_update_outputs {
  coflux_exchange_state*   exchange;
  coflux_interface_fluxes* atmosphere_ocean;
  coflux_net_ocean_fluxes* net_ocean;
} coflux_update_outputs;

/* ------------------------------------------------------------------------------------------------
 * Entry points
 * ---------------------------------------------------------------------------------------------- */
typedef struct coflux_ctx coflux_ctx;

int         coflux_abi_version(void);
const char* coflux_last_error(void);
const char* coflux_build_info(void);      /* "sm_100a nvcc 12.9 ..."                                  */
/* sizeof(struct coflux_<name>) as compiled into the library, e.g. coflux_sizeof("config") — lets an
 * FFI binding (Julia struct mirror, ctypes) verify its layout at load time.  -1 if unknown.        */
int         coflux_sizeof(const char* name);

int coflux_create(coflux_ctx** ctx, const coflux_config* cfg);
int coflux_destroy(coflux_ctx* ctx);

/* Resolve (n1, n2, ñ) for `time` in a series window — host-side helper, also used internally.      */
int coflux_time_indices(const double* times, int32_t Nt, int32_t time_indexing, double cycle_period,
                        double time, int32_t* n1, int32_t* n2, double* frac);

int coflux_interpolate_atmosphere(coflux_ctx*, const coflux_atmos_series* in, double time,
                                  coflux_exchange_state* out, void* cu_stream);

/* Exchange Mp += interpolated (rivers + icebergs).  Run AFTER coflux_interpolate_atmosphere (which sets Mp); the fused
 * coflux_update_state does both when inputs->land is given.                                                       */
int coflux_interpolate_land(coflux_ctx*, const coflux_land_series* in, double time, coflux_exchange_state* inout, void* cu_stream);

int coflux_atmosphere_ocean_fluxes(coflux_ctx*, const coflux_exchange_state* atmos,
                                   const coflux_ocean_surface* ocean,
                                   coflux_interface_fluxes* out, void* cu_stream);

int coflux_atmosphere_sea_ice_fluxes(coflux_ctx*, const coflux_exchange_state* atmos,
                                     const coflux_ocean_surface* ocean,
                                     coflux_sea_ice_state* ice /* top_temperature updated */,
                                     coflux_interface_fluxes* out, void* cu_stream);

int coflux_sea_ice_ocean_fluxes(coflux_ctx*, coflux_ocean_columns* ocean_inout,
                                coflux_sea_ice_state* ice /* previous_thickness updated */,
                                double dt, coflux_ice_ocean_fluxes* out, void* cu_stream);

int coflux_assemble_net_ocean_fluxes(coflux_ctx*, const coflux_exchange_state* atmos,
                                     const coflux_ocean_surface* ocean,
                                     const coflux_interface_fluxes* atmosphere_ocean,
                                     const coflux_sea_ice_state* ice /* may be NULL */,
                                     const coflux_ice_ocean_fluxes* ice_ocean /* may be NULL */,
                                     coflux_net_ocean_fluxes* out, void* cu_stream);

/* compute_net_sea_ice_fluxes!: top = (Q_d + Q_u + Q_c + Q_v)·[ℵ > 0], bottom = Q_frazil + Q_interface, land cells 0.
 * Q_u = ε σ T_s⁴ with the ice top temperature, Q_d = −(1 − α) Q_s − ε Q_ℓ with the sea-ice albedo (prescribed or CCSM3). */
int coflux_assemble_net_sea_ice_fluxes(coflux_ctx*, const coflux_exchange_state* atmos, const coflux_ocean_surface* ocean /* mask; may be NULL */,
                                       const coflux_sea_ice_state* ice, const coflux_interface_fluxes* atmosphere_sea_ice,
                                       const coflux_ice_ocean_fluxes* ice_ocean, coflux_net_sea_ice_fluxes* out, void* cu_stream);

/* Fused interpolate + similarity solve + net-ocean assembly (2 launches: flux kernel, then the
 * centre→face stress kernel).  Writes every contract output exactly once.                          */
int coflux_update_state(coflux_ctx*, const coflux_update_inputs* in, coflux_update_outputs* out,
                        double time, void* cu_stream);

/* End-to-end entry: identical semantics, but the per-step arrays (ocean surface planes in, net
 * ocean fluxes + turbulent fluxes out) are HOST (ideally pinned) buffers.  The call stages them
 * through context-owned device buffers in column slabs on internal streams (H2D / compute / D2H
 * overlapped) and RETURNS AFTER the outputs are in host memory.  The atmosphere series and fi/fj
 * stay device-resident (they change once per forcing window, not per step).                        */
typedef struct coflux_host_step {
  const void* ocean_u; const void* ocean_v; const void* ocean_T; const void* ocean_S; /* (Nxh,Nyh) surface planes, compact, halo-padded like the device parents: dims (Nx+2H, Ny+2H) */
  void* net_u; void* net_v; void* net_T; void* net_S;                                 /* (Nx+2H, Ny+2H) compact                                                        */
  void* latent_heat; void* sensible_heat;                                             /* optional                                                                    */
  int32_t halo;                                                                       /* H of the host planes                                                         */
  int32_t reserved;
} coflux_host_step;
int coflux_update_state_host(coflux_ctx*, const coflux_atmos_series* atmosphere_device,
                             const coflux_host_step* step, double time,
                             int64_t* h2d_bytes, int64_t* d2h_bytes);

/* ------------------------------------------------------------------------------------------------
 * NormalizeSalinity — the per-step post-processing of the net salinity flux, SURVEY §8f row 4.
 * Replaces /root/reference/src/OMIPConfigurations/omip_simulation.jl:187-220 (`NormalizeSalinity`,
 * `salinity_normalizer`): `compute!(Field(Average(flux_field [+ additional_buffer], dims=(1,2))))`
 * followed by `parent(flux_field) .-= mean_total` — the area-weighted global mean of the combined surface
 * salinity flux is removed from the bulk-flux field over its WHOLE parent (halos included), so that the
 * global salt budget integrates to zero.  Immersed (masked) cells do not enter the average.
 * Multi-GPU: every slab computes its partial sums, the host all-reduces the two doubles (NCCL through
 * torch.distributed / MPI — the one real collective of this path), every slab subtracts the same mean.
 * The reduction is a fixed-order tree: bit-reproducible from run to run.
 * ---------------------------------------------------------------------------------------------- */
typedef struct coflux_salinity_normalization {
  coflux_array flux;          /* bulk salinity flux field (= net_fluxes.ocean.S), corrected in place          */
  coflux_array additional;    /* materialised additional flux (`additional_buffer`); ptr NULL: none           */
  coflux_array area;          /* horizontal cell areas Az [m²]; stride_i = 0 on a latitude–longitude grid     */
  coflux_array mask;          /* uint8, 1 = wet; ptr NULL: every cell is averaged                             */
} coflux_salinity_normalization;
/* device_sums[0] = Σ (flux + additional)·Az, device_sums[1] = Σ Az over this slab's wet interior cells       */
int coflux_salinity_flux_sums(coflux_ctx*, const coflux_salinity_normalization*, double* device_sums, void* cu_stream);
/* parent(flux) .-= device_sums[0] / device_sums[1]                                                           */
int coflux_subtract_mean_flux(coflux_ctx*, const coflux_salinity_normalization*, const double* device_sums, void* cu_stream);
/* single slab: both of the above in 2 launches (the subtraction reduces the CTA partials itself)              */
int coflux_normalize_salinity_flux(coflux_ctx*, const coflux_salinity_normalization*, void* cu_stream);

/* ------------------------------------------------------------------------------------------------
 * Surface-forcing front ends of the ocean vertical-mixing closures — the immediate consumers of the net
 * fluxes inside the ocean step.  Each is a by-product of net_fluxes.ocean.{u,v,T,S} at the
 * SAME (i, j) (the reference indexes the face-located stresses with the cell's own indices, no averaging):
 *   KPP       u★  = max(√√(τx² + τy²), u★_min)              /root/reference/src/OMIPConfigurations/KPP/kpp_surface_forcing.jl:18-22
 *             Bo  = −g (α Jᵀ − β Jˢ)   (stabilising positive)  …/KPP/kpp_surface_forcing.jl:28-29 (−top_buoyancy_flux)
 *   NEMO-TKE  u★² = √(τx² + τy²),  e_surf = max(e_min0, Cᵇ u★²)  …/NEMOTKE/nemo_tke_surface_forcing.jl:14-22,
 *                                                                 …/NEMOTKE/nemo_tke_compute_closure_fields.jl:82-90
 * α, β (thermal expansion, haline contraction of the surface cell) are INPUTS: the equation of state belongs to
 * the host ocean model.  Any output / α / β pointer may be NULL (that by-product is skipped).
 * coflux_attach_closure_forcing makes every following coflux_update_state emit the by-products from the
 * centre→face stress kernel itself (τx, τy are in registers there; no extra launch, no re-read of the stresses);
 * coflux_closure_surface_forcing is the stand-alone form (one launch, reads the net fluxes back).
 * ---------------------------------------------------------------------------------------------- */
typedef struct coflux_closure_forcing {
  coflux_array thermal_expansion, haline_contraction;   /* α [1/K], β [kg/g] at k = Nz-1 (2-D)                */
  coflux_array friction_velocity;                       /* KPP u★                                              */
  coflux_array friction_velocity_squared;               /* NEMO-TKE u★²                                        */
  coflux_array surface_tke;                             /* NEMO-TKE e_surf                                     */
  coflux_array buoyancy_flux;                           /* KPP Bo                                              */
  double minimum_friction_velocity;                     /* 1e-6   kpp_parameters.jl:98                         */
  double minimum_surface_tke;                           /* 1e-4   nemo_tke_parameters.jl:54 (rn_emin0)         */
  double Cb;                                            /* 3.75   nemo_tke_parameters.jl:45 (rn_ebb)           */
  double gravitational_acceleration;                    /* buoyancy.formulation.gravitational_acceleration     */
} coflux_closure_forcing;
int coflux_closure_surface_forcing(coflux_ctx*, const coflux_net_ocean_fluxes* net, const coflux_closure_forcing*, void* cu_stream);
int coflux_attach_closure_forcing(coflux_ctx*, const coflux_closure_forcing* forcing_or_null);

/* ------------------------------------------------------------------------------------------------
 * Time-averaged flux diagnostics — SURVEY §8f row 4, second half.  The reference writes the flux fields through
 * `JLD2Writer(...; schedule = AveragedTimeInterval(...))` (omip_diagnostics.jl:125-158), i.e. Oceananigans'
 * WindowedTimeAverage: at every iteration inside the window  result ← (result·T + field·Δt) / (T + Δt), T the time
 * already accumulated.  Attached to a context, the running averages are updated by coflux_update_state itself — in the
 * epilogue of the flux kernel (tracer fluxes, turbulent heat fluxes) and of the stress kernel (τx, τy), while the values
 * are in registers: no flux field is read back.  Any array may be absent.
 *   tau_x tau_y JT JS Qc Qv          tauuo tauvo hfds wfo hfss hfls   (omip_diagnostics.jl:77-82, 125-130)
 *   JT_atmosphere_ocean               JTao = (1 − ℵ) ΣQ / (ρ₀ c₀)
 *   JT_ice_ocean, JS_ice_ocean        JTio = Q_io / (ρ₀ c₀),  JSio = ℵ · J^S_io
 *   JT_frazil                         JTf  = Q_frazil / (ρ₀ c₀)  (accumulated by coflux_sea_ice_ocean_fluxes)
 * (JTn, JSn of omip_diagnostics.jl:85-88 are JT, JS.)
 * ---------------------------------------------------------------------------------------------- */
typedef struct coflux_flux_averages {
  coflux_array tau_x, tau_y, JT, JS, Qc, Qv;
  coflux_array JT_atmosphere_ocean, JT_ice_ocean, JS_ice_ocean, JT_frazil;
  double previous_interval;      /* T: seconds already accumulated in the current window (0 starts a new window) */
  double dt;                     /* Δt of this collection                                                     */
} coflux_flux_averages;
/* every following coflux_update_state / coflux_sea_ice_ocean_fluxes accumulates; call again with the new
 * (previous_interval, dt) before each step; NULL detaches                                                       */
int coflux_attach_flux_averages(coflux_ctx*, const coflux_flux_averages* averages_or_null);
/* stand-alone form: one launch, reads the flux fields back                                                       */
int coflux_accumulate_flux_averages(coflux_ctx*, const coflux_net_ocean_fluxes* net, const coflux_interface_fluxes* atmosphere_ocean,
                                    const coflux_sea_ice_state* ice /* may be NULL */, const coflux_ice_ocean_fluxes* ice_ocean /* may be NULL */,
                                    const coflux_flux_averages* averages, void* cu_stream);

/* ------------------------------------------------------------------------------------------------
 * Device forcing window — SURVEY §8f row 2.  The reference keeps `time_indices_in_memory` levels of every JRA55
 * series on the device and prefetches the next ones (`prefetch = true`, atmosphere.jl:22-27; launch.sh:86-87).  Here:
 * a ring of `capacity` time slots per field in device memory.  coflux_forcing_window_upload copies ONE time level of
 * all fields from (pinned) host memory into the slot  level mod capacity  on the window's own copy stream and returns
 * at once; coflux_forcing_window_wait orders a compute stream behind the uploads of the levels it is about to read
 * (event wait, no host sync); coflux_forcing_window_release records that the work enqueued on a compute stream so far
 * is the last reader of the given levels, so a later upload into one of their slots waits for it (on the copy stream).  The series descriptors
 * handed to the kernels point into the ring with (ring_start, ring_capacity).
 * ---------------------------------------------------------------------------------------------- */
typedef struct coflux_forcing_window coflux_forcing_window;
int coflux_forcing_window_create(coflux_forcing_window** w, coflux_ctx* ctx, int32_t n_fields, int64_t plane_elements, int32_t capacity);
int coflux_forcing_window_destroy(coflux_forcing_window* w);
int coflux_forcing_window_upload(coflux_forcing_window* w, int64_t level, const void* const* host_planes /* n_fields */);
int coflux_forcing_window_field(coflux_forcing_window* w, int32_t field, void** device_ptr /* capacity × plane_elements elements */);
int coflux_forcing_window_wait(coflux_forcing_window* w, int64_t first_level, int64_t last_level, void* cu_stream);
int coflux_forcing_window_release(coflux_forcing_window* w, int64_t first_level, int64_t last_level, void* cu_stream);
/* bytes uploaded so far, and the number of uploads that had to wait for a reader (diagnostics)                 */
int coflux_forcing_window_stats(coflux_forcing_window* w, int64_t* bytes_uploaded, int64_t* levels_uploaded);

/* Diagnostics */
int coflux_launch_count(coflux_ctx*, int64_t* launches);   /* kernels launched through this context  */
/* Per-kernel device timing of coflux_update_state: when enabled, CUDA events are recorded on the
 * caller's stream around the flux kernel and the stress kernel of every call.  coflux_profile_read
 * synchronises on the recorded events, returns the accumulated milliseconds and call count since
 * the last read, and resets the accumulators.                                                      */
int coflux_profile_enable(coflux_ctx*, int32_t enable);
int coflux_profile_read(coflux_ctx*, double* flux_kernel_ms, double* stress_kernel_ms, int64_t* calls);

/* ------------------------------------------------------------------------------------------------
 * Multi-GPU longitude slabs.  One context per GPU/process.  The only cross-slab
 * datum produced by this path is the last column of ρτx needed by the east neighbour's
 * centre→face stress average.  Mode A (grid.ring = 1) needs no exchange at all.  Mode B
 * (grid.ring = 0): the flux kernel stores its seam column straight into the east neighbour's
 * context-owned seam buffer over NVLink (peer mapping via CUDA IPC), published / awaited with
 * stream write-value / wait-value operations on the caller's stream — no host round trip, no
 * separate message.  Protocol: every rank calls coflux_seam_export, the handles are exchanged by
 * the host (torch.distributed / MPI all-gather), every rank calls coflux_seam_attach with its west
 * and east neighbours' handles (periodic ring); afterwards all ranks must call coflux_update_state
 * the same number of times.  With world == 1 a context may attach to itself (periodic single slab).
 * Ordering required of the host: a barrier between the LAST coflux_update_state of the ring and any
 * coflux_seam_detach / coflux_destroy (a neighbour may still be storing into this context's buffer); none is needed
 * between attach and the first update (attach does not reset the flag words).
 * ---------------------------------------------------------------------------------------------- */
#define COFLUX_SEAM_HANDLE_BYTES 128
int coflux_seam_export(coflux_ctx*, void* handle_out /* COFLUX_SEAM_HANDLE_BYTES */);
int coflux_seam_attach(coflux_ctx*, const void* west_handle, const void* east_handle,
                       int32_t rank, int32_t world_size);
int coflux_seam_detach(coflux_ctx*);

#ifdef __cplusplus
}
#endif
#endif /* COFLUX_H */

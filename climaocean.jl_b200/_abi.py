"""ctypes mirror of include/coflux.h and the loader of libcoflux.so.

The struct classes here are plain layout mirrors (field for field) of the C header; a Julia
binding would declare the same `struct`s (see julia/CoFluxExt/src/abi.jl).  `load_library()`
fails loudly when the CUDA library has not been built — there is no fallback of any kind.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("COFLUX_LIB") or os.path.join(_HERE, "lib", "libcoflux.so")   # COFLUX_LIB: tuning variants

ABI_VERSION = 2
F32, F64 = 32, 64

# status codes
OK, ERR_INVALID_ARGUMENT, ERR_CUDA, ERR_UNSUPPORTED, ERR_NO_DEVICE, ERR_ALLOC, ERR_SEAM = 0, -1, -2, -3, -4, -5, -6

# enums (values as in coflux.h)
VISCOSITY_CONSTANT, VISCOSITY_TEMPERATURE_POLY = 0, 1
ROUGHNESS_FIXED, ROUGHNESS_CHARNOCK, ROUGHNESS_REYNOLDS_SCALING = 0, 1, 2
WAVES_CONSTANT, WAVES_WIND_DEPENDENT = 0, 1
FLUXES_SIMILARITY_THEORY, FLUXES_COEFFICIENT_LARGE_YEAGER = 0, 1
STABILITY_EDSON, STABILITY_SHEBA_PAULSON, STABILITY_LARGE_YEAGER, STABILITY_NEUTRAL = 0, 1, 2, 3
PROFILE_LOGARITHMIC, PROFILE_COARE_LOGARITHMIC = 0, 1
VELOCITY_RELATIVE, VELOCITY_WIND = 0, 1
STOP_CONVERGENCE, STOP_FIXED_ITERATIONS = 0, 1
TEMPERATURE_BULK, TEMPERATURE_SKIN = 0, 1
SKIN_CLAMPED_EXPLICIT, SKIN_LINEARIZED_LONGWAVE = 0, 1
TEMPERATURE_CELSIUS, TEMPERATURE_KELVIN = 0, 1
ICE_OCEAN_ICE_BATH, ICE_OCEAN_THREE_EQUATION = 0, 1
FRICTION_VELOCITY_CONSTANT, FRICTION_VELOCITY_MOMENTUM_BASED = 0, 1
TIME_LINEAR, TIME_CYCLICAL, TIME_CLAMP = 0, 1, 2
SEA_ICE_ALBEDO_PRESCRIBED, SEA_ICE_ALBEDO_CCSM3 = 0, 1
SEAM_HANDLE_BYTES = 128

i32, i64, f64 = C.c_int32, C.c_int64, C.c_double


class Array(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("stride_i", i64), ("stride_j", i64), ("stride_k", i64), ("stride_n", i64),
                ("off_i", i32), ("off_j", i32), ("off_k", i32), ("reserved", i32)]


class AirViscosity(C.Structure):
    _fields_ = [("kind", i32), ("reserved", i32), ("nu", f64), ("c0", f64), ("c1", f64), ("c2", f64), ("c3", f64)]


class MomentumRoughness(C.Structure):
    _fields_ = [("kind", i32), ("wave_formulation", i32), ("fixed_length", f64), ("gravity_wave_parameter", f64),
                ("wind_a1", f64), ("wind_a2", f64), ("wind_umax", f64), ("wind_alpha_min", f64),
                ("smooth_wall_parameter", f64), ("maximum_length", f64), ("gravitational_acceleration", f64),
                ("viscosity", AirViscosity)]


class ScalarRoughness(C.Structure):
    _fields_ = [("kind", i32), ("reserved", i32), ("fixed_length", f64), ("reynolds_A", f64), ("reynolds_b", f64),
                ("maximum_length", f64), ("viscosity", AirViscosity)]


class FluxParams(C.Structure):
    _fields_ = [("formulation", i32), ("stability_functions", i32), ("similarity_form", i32),
                ("velocity_formulation", i32), ("stop_kind", i32), ("max_iterations", i32),
                ("interface_temperature", i32), ("skin_temperature_update", i32),
                ("tolerance", f64), ("von_karman_constant", f64), ("turbulent_prandtl_number", f64),
                ("gustiness_parameter", f64), ("minimum_gustiness", f64), ("initial_scale", f64),
                ("ly_minimum_wind", f64), ("skin_max_delta_T", f64),
                ("momentum_roughness", MomentumRoughness), ("temperature_roughness", ScalarRoughness),
                ("water_vapor_roughness", ScalarRoughness)]


class Thermodynamics(C.Structure):
    _fields_ = [(n, f64) for n in (
        "gas_constant", "dry_air_molar_mass", "water_molar_mass", "dry_air_adiabatic_exponent",
        "water_vapor_heat_capacity", "liquid_water_heat_capacity", "ice_heat_capacity",
        "reference_vaporization_enthalpy", "reference_sublimation_enthalpy", "reference_temperature",
        "triple_point_temperature", "triple_point_pressure", "water_freezing_temperature",
        "total_ice_nucleation_temperature")]


class AtmosphereProperties(C.Structure):
    _fields_ = [("thermodynamics", Thermodynamics), ("surface_layer_height", f64), ("boundary_layer_height", f64),
                ("gravitational_acceleration", f64)]


class OceanProperties(C.Structure):
    _fields_ = [("reference_density", f64), ("heat_capacity", f64), ("freshwater_density", f64),
                ("minimum_salinity", f64), ("temperature_units", i32), ("reserved", i32),
                ("salt_water_molar_mass", f64), ("constituent_molar_mass", f64 * 4),
                ("constituent_mass_fraction", f64 * 4)]


class Ccsm3Albedo(C.Structure):
    _fields_ = [(n, f64) for n in (
        "ice_visible", "ice_near_infrared", "snow_visible", "snow_near_infrared", "thickness_scale",
        "melt_temperature_range", "ice_melt_change", "snow_visible_melt_change", "snow_near_infrared_melt_change",
        "snow_patchiness", "ocean_albedo", "visible_fraction", "melting_temperature")]


class RadiationProperties(C.Structure):
    _fields_ = [("stefan_boltzmann_constant", f64), ("ocean_albedo", f64), ("ocean_emissivity", f64),
                ("sea_ice_emissivity", f64), ("sea_ice_albedo", f64), ("shortwave_penetrates", i32),
                ("sea_ice_albedo_kind", i32), ("ccsm3", Ccsm3Albedo)]


class IceOceanParams(C.Structure):
    _fields_ = [("heat_flux", i32), ("friction_velocity", i32)] + [(n, f64) for n in (
        "characteristic_melting_speed", "liquidus_freshwater_melting_temperature", "liquidus_slope",
        "heat_transfer_coefficient", "salt_transfer_coefficient", "constant_friction_velocity",
        "minimum_friction_velocity", "ice_density", "ice_latent_heat", "ice_ocean_drag_coefficient",
        "ice_conductivity", "ice_consolidation_thickness")]


class GridDesc(C.Structure):
    _fields_ = [("Nx", i32), ("Ny", i32), ("Nz", i32), ("ring", i32), ("periodic_x", i32), ("reserved", i32)]


class Config(C.Structure):
    _fields_ = [("abi_version", i32), ("dtype", i32), ("device", i32), ("reserved", i32), ("grid", GridDesc),
                ("atmosphere_ocean", FluxParams), ("atmosphere_sea_ice", FluxParams), ("ice_ocean", IceOceanParams),
                ("atmosphere", AtmosphereProperties), ("ocean", OceanProperties), ("radiation", RadiationProperties)]


class AtmosSeries(C.Structure):
    _fields_ = [(n, Array) for n in ("u", "v", "T", "q", "p", "Qs", "Ql", "rain", "snow")] + [
        ("times", C.POINTER(f64)), ("Nt", i32), ("time_indexing", i32), ("cycle_period", f64),
        ("fi", Array), ("fj", Array), ("cos_theta", Array), ("sin_theta", Array),
        ("ring_start", i32), ("ring_capacity", i32)]


class LandSeries(C.Structure):
    _fields_ = [("rivers", Array), ("icebergs", Array), ("times", C.POINTER(f64)), ("Nt", i32), ("time_indexing", i32),
                ("cycle_period", f64), ("fi", Array), ("fj", Array), ("ring_start", i32), ("ring_capacity", i32)]


class ExchangeState(C.Structure):
    _fields_ = [(n, Array) for n in ("u", "v", "T", "p", "q", "Qs", "Ql", "Mp")]


class OceanSurface(C.Structure):
    _fields_ = [(n, Array) for n in ("u", "v", "T", "S", "mask")]


class InterfaceFluxes(C.Structure):
    _fields_ = [(n, Array) for n in ("latent_heat", "sensible_heat", "water_vapor", "x_momentum", "y_momentum",
                                     "interface_temperature", "friction_velocity", "temperature_scale",
                                     "humidity_scale", "iterations")]


class SeaIceState(C.Structure):
    _fields_ = [(n, Array) for n in ("thickness", "previous_thickness", "concentration", "salinity", "u", "v",
                                     "top_temperature", "snow_thickness", "albedo")]


class OceanColumns(C.Structure):
    _fields_ = [(n, Array) for n in ("T", "S", "dz", "u", "v")]


class IceOceanFluxes(C.Structure):
    _fields_ = [(n, Array) for n in ("frazil_heat", "interface_heat", "salt", "x_momentum", "y_momentum")]


class NetOceanFluxes(C.Structure):
    _fields_ = [(n, Array) for n in ("u", "v", "T", "S", "upwelling_longwave", "downwelling_longwave",
                                     "downwelling_shortwave", "penetrating_shortwave")]


class NetSeaIceFluxes(C.Structure):
    _fields_ = [(n, Array) for n in ("top_heat", "bottom_heat", "top_u", "top_v")]


class FluxAverages(C.Structure):
    _fields_ = [(n, Array) for n in ("tau_x", "tau_y", "JT", "JS", "Qc", "Qv", "JT_atmosphere_ocean", "JT_ice_ocean",
                                     "JS_ice_ocean", "JT_frazil")] + [("previous_interval", f64), ("dt", f64)]


class UpdateInputs(C.Structure):
    _fields_ = [("atmosphere", C.POINTER(AtmosSeries)), ("ocean", C.POINTER(OceanSurface)),
                ("sea_ice", C.POINTER(SeaIceState)), ("ice_ocean", C.POINTER(IceOceanFluxes)),
                ("land", C.POINTER(LandSeries))]


class UpdateOutputs(C.Structure):
    _fields_ = [("exchange", C.POINTER(ExchangeState)), ("atmosphere_ocean", C.POINTER(InterfaceFluxes)),
                ("net_ocean", C.POINTER(NetOceanFluxes))]


class HostStep(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("ocean_u", "ocean_v", "ocean_T", "ocean_S", "net_u", "net_v", "net_T",
                                          "net_S", "latent_heat", "sensible_heat")] + [("halo", i32), ("reserved", i32)]


class ClosureForcing(C.Structure):
    _fields_ = [("thermal_expansion", Array), ("haline_contraction", Array), ("friction_velocity", Array),
                ("friction_velocity_squared", Array), ("surface_tke", Array), ("buoyancy_flux", Array),
                ("minimum_friction_velocity", f64), ("minimum_surface_tke", f64), ("Cb", f64), ("gravitational_acceleration", f64)]


class SalinityNormalization(C.Structure):
    _fields_ = [("flux", Array), ("additional", Array), ("area", Array), ("mask", Array)]


STRUCTS = {"array": Array, "air_viscosity": AirViscosity, "momentum_roughness": MomentumRoughness,
           "scalar_roughness": ScalarRoughness, "flux_params": FluxParams, "thermodynamics": Thermodynamics,
           "atmosphere_properties": AtmosphereProperties, "ocean_properties": OceanProperties,
           "radiation_properties": RadiationProperties, "ice_ocean_params": IceOceanParams, "grid_desc": GridDesc,
           "config": Config, "atmos_series": AtmosSeries, "exchange_state": ExchangeState,
           "ocean_surface": OceanSurface, "interface_fluxes": InterfaceFluxes, "sea_ice_state": SeaIceState,
           "ocean_columns": OceanColumns, "ice_ocean_fluxes": IceOceanFluxes, "net_ocean_fluxes": NetOceanFluxes,
           "update_inputs": UpdateInputs, "update_outputs": UpdateOutputs, "host_step": HostStep,
           "salinity_normalization": SalinityNormalization, "closure_forcing": ClosureForcing,
           "land_series": LandSeries, "ccsm3_albedo": Ccsm3Albedo, "net_sea_ice_fluxes": NetSeaIceFluxes,
           "flux_averages": FluxAverages}

# every symbol include/coflux.h declares
EXPORTS = ("coflux_abi_version", "coflux_last_error", "coflux_build_info", "coflux_sizeof", "coflux_default_config",
           "coflux_apply_flux_configuration", "coflux_create", "coflux_destroy", "coflux_time_indices",
           "coflux_interpolate_atmosphere", "coflux_atmosphere_ocean_fluxes", "coflux_atmosphere_sea_ice_fluxes",
           "coflux_sea_ice_ocean_fluxes", "coflux_assemble_net_ocean_fluxes", "coflux_update_state",
           "coflux_update_state_host", "coflux_launch_count", "coflux_profile_enable", "coflux_profile_read", "coflux_seam_export", "coflux_seam_attach",
           "coflux_seam_detach", "coflux_salinity_flux_sums", "coflux_subtract_mean_flux", "coflux_normalize_salinity_flux",
           "coflux_closure_surface_forcing", "coflux_attach_closure_forcing",
           "coflux_interpolate_land", "coflux_assemble_net_sea_ice_fluxes", "coflux_attach_flux_averages",
           "coflux_accumulate_flux_averages", "coflux_forcing_window_create", "coflux_forcing_window_destroy",
           "coflux_forcing_window_upload", "coflux_forcing_window_field", "coflux_forcing_window_wait",
           "coflux_forcing_window_release", "coflux_forcing_window_stats")


class CofluxError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"coflux status {status}: {message}")
        self.status = status
        self.message = message


_lib = None


def load_library(path=None):
    """dlopen libcoflux.so (built by __graft_entry__.build()).  Raises if it is missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise ImportError(f"{p} not found: build the CUDA library first (python -c 'import __graft_entry__ as g; "
                          f"g.build()').  coflux has no CPU fallback.")
    lib = C.CDLL(p)
    P = C.POINTER
    vp = C.c_void_p
    lib.coflux_abi_version.restype = C.c_int
    lib.coflux_last_error.restype = C.c_char_p
    lib.coflux_build_info.restype = C.c_char_p
    lib.coflux_sizeof.argtypes = [C.c_char_p]
    lib.coflux_default_config.argtypes = [P(Config), i32, i32, i32, i32]
    lib.coflux_apply_flux_configuration.argtypes = [P(Config), C.c_char_p, i32]
    lib.coflux_create.argtypes = [P(vp), P(Config)]
    lib.coflux_destroy.argtypes = [vp]
    lib.coflux_time_indices.argtypes = [P(f64), i32, i32, f64, f64, P(i32), P(i32), P(f64)]
    lib.coflux_interpolate_atmosphere.argtypes = [vp, P(AtmosSeries), f64, P(ExchangeState), vp]
    lib.coflux_atmosphere_ocean_fluxes.argtypes = [vp, P(ExchangeState), P(OceanSurface), P(InterfaceFluxes), vp]
    lib.coflux_atmosphere_sea_ice_fluxes.argtypes = [vp, P(ExchangeState), P(OceanSurface), P(SeaIceState),
                                                     P(InterfaceFluxes), vp]
    lib.coflux_sea_ice_ocean_fluxes.argtypes = [vp, P(OceanColumns), P(SeaIceState), f64, P(IceOceanFluxes), vp]
    lib.coflux_assemble_net_ocean_fluxes.argtypes = [vp, P(ExchangeState), P(OceanSurface), P(InterfaceFluxes),
                                                     P(SeaIceState), P(IceOceanFluxes), P(NetOceanFluxes), vp]
    lib.coflux_update_state.argtypes = [vp, P(UpdateInputs), P(UpdateOutputs), f64, vp]
    lib.coflux_update_state_host.argtypes = [vp, P(AtmosSeries), P(HostStep), f64, P(i64), P(i64)]
    lib.coflux_launch_count.argtypes = [vp, P(i64)]
    lib.coflux_profile_enable.argtypes = [vp, i32]
    lib.coflux_profile_read.argtypes = [vp, P(f64), P(f64), P(i64)]
    lib.coflux_seam_export.argtypes = [vp, vp]
    lib.coflux_seam_attach.argtypes = [vp, vp, vp, i32, i32]
    lib.coflux_seam_detach.argtypes = [vp]
    lib.coflux_salinity_flux_sums.argtypes = [vp, P(SalinityNormalization), vp, vp]
    lib.coflux_subtract_mean_flux.argtypes = [vp, P(SalinityNormalization), vp, vp]
    lib.coflux_normalize_salinity_flux.argtypes = [vp, P(SalinityNormalization), vp]
    lib.coflux_closure_surface_forcing.argtypes = [vp, P(NetOceanFluxes), P(ClosureForcing), vp]
    lib.coflux_attach_closure_forcing.argtypes = [vp, P(ClosureForcing)]
    lib.coflux_interpolate_land.argtypes = [vp, P(LandSeries), f64, P(ExchangeState), vp]
    lib.coflux_assemble_net_sea_ice_fluxes.argtypes = [vp, P(ExchangeState), P(OceanSurface), P(SeaIceState), P(InterfaceFluxes),
                                                       P(IceOceanFluxes), P(NetSeaIceFluxes), vp]
    lib.coflux_attach_flux_averages.argtypes = [vp, P(FluxAverages)]
    lib.coflux_accumulate_flux_averages.argtypes = [vp, P(NetOceanFluxes), P(InterfaceFluxes), P(SeaIceState), P(IceOceanFluxes),
                                                    P(FluxAverages), vp]
    lib.coflux_forcing_window_create.argtypes = [P(vp), vp, i32, i64, i32]
    lib.coflux_forcing_window_destroy.argtypes = [vp]
    lib.coflux_forcing_window_upload.argtypes = [vp, i64, P(vp)]
    lib.coflux_forcing_window_field.argtypes = [vp, i32, P(vp)]
    lib.coflux_forcing_window_wait.argtypes = [vp, i64, i64, vp]
    lib.coflux_forcing_window_release.argtypes = [vp, i64, i64, vp]
    lib.coflux_forcing_window_stats.argtypes = [vp, P(i64), P(i64)]
    # layout check: a binding whose struct mirrors drifted from the header must not run (it would corrupt memory)
    if lib.coflux_abi_version() != ABI_VERSION:
        raise ImportError(f"{p}: library ABI version {lib.coflux_abi_version()} != binding {ABI_VERSION}; rebuild (python __graft_entry__.py)")
    for sname, cls in STRUCTS.items():
        n = lib.coflux_sizeof(sname.encode())
        if n != C.sizeof(cls):
            raise ImportError(f"{p}: sizeof(coflux_{sname}) is {n} in the library but {C.sizeof(cls)} in the ctypes mirror; rebuild the library")
    for name in EXPORTS:
        fn = getattr(lib, name)
        if fn.restype is C.c_int and name not in ("coflux_abi_version", "coflux_sizeof"):
            fn.restype = C.c_int
    if path is None:
        _lib = lib
    return lib


def check(status, lib=None):
    if status != OK:
        lib = lib or load_library()
        raise CofluxError(status, lib.coflux_last_error().decode())
    return status


def default_config(Nx, Ny, Nz, dtype=F64, flux_configuration="default", velocity_formulation="relative", lib=None):
    """coflux_default_config + coflux_apply_flux_configuration (build_coupled_model's options,
    /root/reference/src/OMIPConfigurations/omip_simulation.jl:123-164)."""
    lib = lib or load_library()
    cfg = Config()
    check(lib.coflux_default_config(C.byref(cfg), Nx, Ny, Nz, dtype), lib)
    vel = {"relative": VELOCITY_RELATIVE, "wind": VELOCITY_WIND}.get(velocity_formulation)
    if flux_configuration == "default":
        vel = VELOCITY_RELATIVE    # `:default` returns before velocity_formulation is looked at (omip_simulation.jl:127-133)
    if vel is None:
        raise ValueError(f"Unknown velocity_formulation: {velocity_formulation}. Options: :relative, :wind")
    check(lib.coflux_apply_flux_configuration(C.byref(cfg), flux_configuration.encode(), vel), lib)
    return cfg

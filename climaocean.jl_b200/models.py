"""Host-side mirror of the reference's interface for the surface-flux path.

The reference reaches this path through Julia constructors and generic functions
(/root/reference/src/OMIPConfigurations/omip_simulation.jl:40-164, README.md:74-75,
examples/one_degree_tripolar_ocean_sea_ice.jl:17-42): `SimilarityTheoryFluxes`,
`CoefficientBasedFluxes`, `ComponentInterfaces`, `OceanSeaIceModel`, `time_step!`, `update_state!`.
No Julia toolchain exists in this environment, so the same names, keyword arguments, option
strings and error messages are mirrored here in Python on top of the C ABI; `julia/CoFluxExt`
carries the (unexecuted) Julia binding.  This module only assembles parameters and array
descriptors — all arithmetic runs in the CUDA library.
"""
import ctypes as C
import math

import numpy as np

from . import _abi
from .engine import Engine
from .fields import Field, LatitudeLongitudeGrid
from .state import SurfaceFluxData

# ------------------------------------------------------------------------------------------------
# parameter objects (OMIPConfigurations.jl:15-33 import list)
# ------------------------------------------------------------------------------------------------


class TemperatureDependentAirViscosity:
    def __init__(self, FT=np.float64, c0=1.326e-5, c1=None, c2=None, c3=None):
        self.c0 = c0
        self.c1 = c0 * 6.542e-3 if c1 is None else c1
        self.c2 = c0 * 8.301e-6 if c2 is None else c2
        self.c3 = -c0 * 4.84e-9 if c3 is None else c3

    def fill(self, v):
        v.kind = _abi.VISCOSITY_TEMPERATURE_POLY
        v.c0, v.c1, v.c2, v.c3 = self.c0, self.c1, self.c2, self.c3


class ConstantAirViscosity:
    def __init__(self, nu=1.5e-5):
        self.nu = nu

    def fill(self, v):
        v.kind = _abi.VISCOSITY_CONSTANT
        v.nu = self.nu


def _fill_viscosity(v, visc):
    if isinstance(visc, (int, float)):
        visc = ConstantAirViscosity(float(visc))
    visc.fill(v)


class WindDependentWaveFormulation:
    """Edson et al. (2013) eq. 13 Charnock parameter α = a1·min(U, Umax) + a2."""

    def __init__(self, FT=np.float64, a1=0.0017, a2=-0.005, Umax=19.0, alpha_min=0.0):
        self.a1, self.a2, self.Umax, self.alpha_min = a1, a2, Umax, alpha_min


class MomentumRoughnessLength:
    def __init__(self, FT=np.float64, wave_formulation=0.02, air_kinematic_viscosity=1.5e-5, smooth_wall_parameter=0.11,
                 maximum_roughness_length=1.0, gravitational_acceleration=9.81):
        self.wave_formulation = wave_formulation
        self.air_kinematic_viscosity = air_kinematic_viscosity
        self.smooth_wall_parameter = smooth_wall_parameter
        self.maximum_roughness_length = maximum_roughness_length
        self.gravitational_acceleration = gravitational_acceleration

    def fill(self, m):
        m.kind = _abi.ROUGHNESS_CHARNOCK
        if isinstance(self.wave_formulation, WindDependentWaveFormulation):
            w = self.wave_formulation
            m.wave_formulation = _abi.WAVES_WIND_DEPENDENT
            m.wind_a1, m.wind_a2, m.wind_umax, m.wind_alpha_min = w.a1, w.a2, w.Umax, w.alpha_min
        else:
            m.wave_formulation = _abi.WAVES_CONSTANT
            m.gravity_wave_parameter = float(self.wave_formulation)
        m.smooth_wall_parameter = self.smooth_wall_parameter
        m.maximum_length = self.maximum_roughness_length
        m.gravitational_acceleration = self.gravitational_acceleration
        _fill_viscosity(m.viscosity, self.air_kinematic_viscosity)


class ScalarRoughnessLength:
    def __init__(self, FT=np.float64, air_kinematic_viscosity=1.5e-5, reynolds_A=5.85e-5, reynolds_b=0.72,
                 maximum_roughness_length=1.6e-4):
        self.air_kinematic_viscosity = air_kinematic_viscosity
        self.reynolds_A, self.reynolds_b = reynolds_A, reynolds_b
        self.maximum_roughness_length = maximum_roughness_length

    def fill(self, s):
        s.kind = _abi.ROUGHNESS_REYNOLDS_SCALING
        s.reynolds_A, s.reynolds_b = self.reynolds_A, self.reynolds_b
        s.maximum_length = self.maximum_roughness_length
        _fill_viscosity(s.viscosity, self.air_kinematic_viscosity)


def _fill_roughness(dst, value, momentum):
    if isinstance(value, (int, float)):           # "roughness either a struct or a plain number"
        dst.kind = _abi.ROUGHNESS_FIXED
        dst.fixed_length = float(value)
    else:
        value.fill(dst)


class LogarithmicSimilarityProfile:
    code = _abi.PROFILE_LOGARITHMIC


class COARELogarithmicSimilarityProfile:
    code = _abi.PROFILE_COARE_LOGARITHMIC


class _Stability:
    def __init__(self, code):
        self.code = code


def atmosphere_ocean_stability_functions(FT=np.float64):
    return _Stability(_abi.STABILITY_EDSON)


def atmosphere_sea_ice_stability_functions(FT=np.float64):
    return _Stability(_abi.STABILITY_SHEBA_PAULSON)


def large_yeager_stability_functions(FT=np.float64):
    return _Stability(_abi.STABILITY_LARGE_YEAGER)


class ConvergenceStopCriteria:
    def __init__(self, tolerance=1e-8, maxiter=100):
        self.tolerance, self.maxiter = tolerance, maxiter


class FixedIterations:
    def __init__(self, iterations=5):
        self.iterations = int(iterations)


class RelativeVelocity:
    code = _abi.VELOCITY_RELATIVE


class WindVelocity:
    code = _abi.VELOCITY_WIND


class LargeYeagerTransferCoefficients:
    def __init__(self, FT=np.float64, minimum_wind=0.5):
        self.minimum_wind = minimum_wind


def _base_flux_params(stability):
    lib = _abi.load_library()
    cfg = _abi.Config()
    _abi.check(lib.coflux_default_config(C.byref(cfg), 1, 1, 1, _abi.F64), lib)
    p = _abi.FluxParams.from_buffer_copy(cfg.atmosphere_ocean)
    p.stability_functions = stability
    return p


class SimilarityTheoryFluxes:
    """Monin–Obukhov similarity-theory fluxes (omip_simulation.jl:42-49,63-69,106-113 kwargs)."""

    def __init__(self, FT=np.float64, stability_functions=None, similarity_form=None, gustiness_parameter=1.0,
                 minimum_gustiness=0.0, momentum_roughness_length=None, temperature_roughness_length=None,
                 water_vapor_roughness_length=None, von_karman_constant=0.4, solver_stop_criteria=None,
                 solver_tolerance=1e-8, solver_maxiter=100):
        self.stability_functions = stability_functions or atmosphere_ocean_stability_functions(FT)
        self.similarity_form = similarity_form or LogarithmicSimilarityProfile()
        self.gustiness_parameter = gustiness_parameter
        self.minimum_gustiness = minimum_gustiness
        self.momentum_roughness_length = MomentumRoughnessLength(FT) if momentum_roughness_length is None else momentum_roughness_length
        self.temperature_roughness_length = ScalarRoughnessLength(FT) if temperature_roughness_length is None else temperature_roughness_length
        self.water_vapor_roughness_length = ScalarRoughnessLength(FT) if water_vapor_roughness_length is None else water_vapor_roughness_length
        self.von_karman_constant = von_karman_constant
        self.solver_stop_criteria = solver_stop_criteria or ConvergenceStopCriteria(solver_tolerance, solver_maxiter)

    def to_params(self):
        p = _base_flux_params(self.stability_functions.code)
        p.formulation = _abi.FLUXES_SIMILARITY_THEORY
        p.similarity_form = self.similarity_form.code
        p.gustiness_parameter = float(self.gustiness_parameter)
        p.minimum_gustiness = float(self.minimum_gustiness)
        p.von_karman_constant = float(self.von_karman_constant)
        _fill_roughness(p.momentum_roughness, self.momentum_roughness_length, True)
        _fill_roughness(p.temperature_roughness, self.temperature_roughness_length, False)
        _fill_roughness(p.water_vapor_roughness, self.water_vapor_roughness_length, False)
        _fill_stop(p, self.solver_stop_criteria)
        return p


class CoefficientBasedFluxes:
    """Large & Yeager coefficient-based fluxes (omip_simulation.jl:86-89)."""

    def __init__(self, FT=np.float64, transfer_coefficients=None, solver_stop_criteria=None):
        self.transfer_coefficients = transfer_coefficients or LargeYeagerTransferCoefficients(FT)
        self.solver_stop_criteria = solver_stop_criteria or FixedIterations(5)

    def to_params(self):
        p = _base_flux_params(_abi.STABILITY_LARGE_YEAGER)
        p.formulation = _abi.FLUXES_COEFFICIENT_LARGE_YEAGER
        p.ly_minimum_wind = float(self.transfer_coefficients.minimum_wind)
        _fill_stop(p, self.solver_stop_criteria)
        return p


def _fill_stop(p, crit):
    if isinstance(crit, FixedIterations):
        p.stop_kind, p.max_iterations = _abi.STOP_FIXED_ITERATIONS, crit.iterations
    else:
        p.stop_kind, p.max_iterations, p.tolerance = _abi.STOP_CONVERGENCE, int(crit.maxiter), float(crit.tolerance)


class MomentumBasedFrictionVelocity:
    pass


class ThreeEquationHeatFlux:
    def __init__(self, friction_velocity=0.002):
        self.friction_velocity = friction_velocity


class IceBathHeatFlux:
    def __init__(self, characteristic_melting_speed=1e-5):
        self.characteristic_melting_speed = characteristic_melting_speed


# --- the OMIP presets, same names as omip_simulation.jl:40-113 ---
def corrected_atmosphere_ocean_fluxes(FT=np.float64, minimum_gustiness=0.5):
    visc = TemperatureDependentAirViscosity(FT)
    return SimilarityTheoryFluxes(FT, similarity_form=COARELogarithmicSimilarityProfile(), minimum_gustiness=minimum_gustiness,
                                  momentum_roughness_length=MomentumRoughnessLength(FT, wave_formulation=WindDependentWaveFormulation(FT),
                                                                                    air_kinematic_viscosity=TemperatureDependentAirViscosity(FT)),
                                  temperature_roughness_length=ScalarRoughnessLength(FT, air_kinematic_viscosity=visc),
                                  water_vapor_roughness_length=ScalarRoughnessLength(FT, air_kinematic_viscosity=visc))


def corrected_atmosphere_sea_ice_fluxes(FT=np.float64):
    return SimilarityTheoryFluxes(FT, stability_functions=atmosphere_sea_ice_stability_functions(FT),
                                  similarity_form=COARELogarithmicSimilarityProfile(), minimum_gustiness=0.2,
                                  momentum_roughness_length=5e-4, temperature_roughness_length=5e-5,
                                  water_vapor_roughness_length=5e-5)


def corrected_ice_ocean_heat_flux():
    return ThreeEquationHeatFlux(friction_velocity=MomentumBasedFrictionVelocity())


def ncar_atmosphere_ocean_fluxes(FT=np.float64):
    return CoefficientBasedFluxes(FT, transfer_coefficients=LargeYeagerTransferCoefficients(FT),
                                  solver_stop_criteria=FixedIterations(5))


def ncar_atmosphere_sea_ice_fluxes(FT=np.float64):
    return SimilarityTheoryFluxes(FT, stability_functions=large_yeager_stability_functions(FT),
                                  similarity_form=COARELogarithmicSimilarityProfile(), gustiness_parameter=0.0,
                                  minimum_gustiness=0.5, momentum_roughness_length=5e-4,
                                  temperature_roughness_length=5e-4, water_vapor_roughness_length=5e-4)


# ------------------------------------------------------------------------------------------------
# components
# ------------------------------------------------------------------------------------------------
class Clock:
    def __init__(self, time=0.0):
        self.time = float(time)
        self.iteration = 0


class _Named:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class OceanSimulation:
    """ocean_simulation(grid): holds the ocean prognostic Fields the flux path reads (u, v, T, S) and
    a `step` callable standing in for the Oceananigans HydrostaticFreeSurfaceModel time step
    (out of scope: "Oceananigans ocean stencil unchanged", BASELINE.json config 3)."""

    def __init__(self, grid, data, step=None):
        self.grid = grid
        self.model = _Named(grid=grid, velocities=_Named(u=data.ocean["u"], v=data.ocean["v"]),
                            tracers=_Named(T=data.ocean["T"], S=data.ocean["S"]), clock=Clock())
        self._step = step

    def time_step(self, dt):
        if self._step is not None:
            self._step(self, dt)
        self.model.clock.time += dt


def ocean_simulation(grid, data, step=None):
    return OceanSimulation(grid, data, step)


class SeaIceSimulation:
    def __init__(self, grid, data, step=None):
        ice = data.ice
        self.model = _Named(grid=grid, ice_thickness=ice["thickness"], ice_concentration=ice["concentration"],
                            ice_salinity=ice["salinity"], velocities=_Named(u=ice["u"], v=ice["v"]),
                            top_surface_temperature=ice["top_temperature"], clock=Clock())
        self._step = step

    def time_step(self, dt):
        if self._step is not None:
            self._step(self, dt)
        self.model.clock.time += dt


def sea_ice_simulation(grid, data, step=None):
    return SeaIceSimulation(grid, data, step)


class PrescribedAtmosphere:
    """Prescribed atmosphere FieldTimeSeries window (u, v, T, q, p, Qs, Qℓ, rain, snow) living in `data`."""

    def __init__(self, data, surface_layer_height=10.0, boundary_layer_height=512.0):
        self.data = data
        self.times = data.times
        self.surface_layer_height = surface_layer_height
        self.boundary_layer_height = boundary_layer_height


class PrescribedLand:
    """JRA55PrescribedLand (atmosphere.jl:46): river runoff + iceberg calving series living in `data.land`; their sum is
    added to the exchange freshwater flux (enters J^S like rain and snow)."""

    def __init__(self, data):
        if data.land is None:
            raise ValueError("PrescribedLand needs land freshwater series (SurfaceFluxData.land)")
        self.data = data
        self.times = data.land_times


class SeaIceAlbedo:
    """SeaIceAlbedo(hi, hs, Ts) — CCSM3 thickness / snow / temperature dependent albedo evaluated from the LIVE sea-ice
    fields (atmosphere.jl:31-44).  Keyword arguments override the CCSM3 constants (coflux_ccsm3_albedo)."""

    def __init__(self, ice_thickness=None, snow_thickness=None, surface_temperature=None, **params):
        self.ice_thickness, self.snow_thickness, self.surface_temperature = ice_thickness, snow_thickness, surface_temperature
        self.params = params


class SurfaceRadiationProperties:
    def __init__(self, albedo, emissivity):
        self.albedo, self.emissivity = albedo, emissivity


class Radiation:
    def __init__(self, ocean_surface=None, sea_ice_surface=None, stefan_boltzmann_constant=5.67e-8):
        self.ocean_surface = ocean_surface or SurfaceRadiationProperties(0.06, 1.0)      # atmosphere.jl:43
        self.sea_ice_surface = sea_ice_surface or SurfaceRadiationProperties(0.7, 1.0)
        self.stefan_boltzmann_constant = stefan_boltzmann_constant


class ComponentInterfaces:
    """ComponentInterfaces(atmosphere, ocean, sea_ice; radiation, atmosphere_ocean_fluxes, ...)
    (omip_simulation.jl:128-158).  Allocates nothing new: the 2-D flux Fields already live in the
    SurfaceFluxData; builds the coflux context from the parameter objects."""

    def __init__(self, atmosphere, ocean, sea_ice=None, radiation=None, land=None, atmosphere_ocean_fluxes=None,
                 atmosphere_sea_ice_fluxes=None, sea_ice_ocean_heat_flux=None,
                 atmosphere_ocean_velocity_difference=None, atmosphere_sea_ice_velocity_difference=None,
                 ocean_minimum_salinity=1.0, device_index=0, ring=1):
        data = atmosphere.data
        grid = ocean.grid
        dtype = _abi.F64 if np.dtype(grid.dtype) == np.float64 else _abi.F32
        cfg = _abi.default_config(grid.Nx, grid.Ny, grid.Nz, dtype)
        cfg.device = device_index
        cfg.grid.ring = ring
        if atmosphere_ocean_fluxes is not None:
            cfg.atmosphere_ocean = atmosphere_ocean_fluxes.to_params()
        if atmosphere_sea_ice_fluxes is not None:
            p = atmosphere_sea_ice_fluxes.to_params()
            p.interface_temperature = _abi.TEMPERATURE_SKIN
            cfg.atmosphere_sea_ice = p
        if atmosphere_ocean_velocity_difference is not None:
            cfg.atmosphere_ocean.velocity_formulation = atmosphere_ocean_velocity_difference.code
        if atmosphere_sea_ice_velocity_difference is not None:
            cfg.atmosphere_sea_ice.velocity_formulation = atmosphere_sea_ice_velocity_difference.code
        if isinstance(sea_ice_ocean_heat_flux, ThreeEquationHeatFlux):
            cfg.ice_ocean.heat_flux = _abi.ICE_OCEAN_THREE_EQUATION
            if isinstance(sea_ice_ocean_heat_flux.friction_velocity, MomentumBasedFrictionVelocity):
                cfg.ice_ocean.friction_velocity = _abi.FRICTION_VELOCITY_MOMENTUM_BASED
            else:
                cfg.ice_ocean.friction_velocity = _abi.FRICTION_VELOCITY_CONSTANT
                cfg.ice_ocean.constant_friction_velocity = float(sea_ice_ocean_heat_flux.friction_velocity)
        elif isinstance(sea_ice_ocean_heat_flux, IceBathHeatFlux):
            cfg.ice_ocean.heat_flux = _abi.ICE_OCEAN_ICE_BATH
            cfg.ice_ocean.characteristic_melting_speed = sea_ice_ocean_heat_flux.characteristic_melting_speed
        radiation = radiation or Radiation()
        cfg.radiation.stefan_boltzmann_constant = radiation.stefan_boltzmann_constant
        cfg.radiation.ocean_albedo = radiation.ocean_surface.albedo
        cfg.radiation.ocean_emissivity = radiation.ocean_surface.emissivity
        if isinstance(radiation.sea_ice_surface.albedo, SeaIceAlbedo):       # atmosphere.jl:39-44
            cfg.radiation.sea_ice_albedo_kind = _abi.SEA_ICE_ALBEDO_CCSM3
            for k, v in radiation.sea_ice_surface.albedo.params.items():
                setattr(cfg.radiation.ccsm3, k, float(v))
        else:
            cfg.radiation.sea_ice_albedo = radiation.sea_ice_surface.albedo
        cfg.radiation.sea_ice_emissivity = radiation.sea_ice_surface.emissivity
        if land is not None and not isinstance(land, PrescribedLand):
            raise NotImplementedError("land must be a PrescribedLand (river runoff + iceberg calving series) or None")
        self.land = land
        cfg.ocean.minimum_salinity = float(ocean_minimum_salinity)
        cfg.atmosphere.surface_layer_height = atmosphere.surface_layer_height
        cfg.atmosphere.boundary_layer_height = atmosphere.boundary_layer_height
        self.cfg = cfg
        self.data = data
        self.engine = Engine(cfg)
        self.has_sea_ice = sea_ice is not None
        # the Field handles the reference exposes (omip_diagnostics.jl:77-89)
        self.net_fluxes = _Named(ocean=_Named(u=data.net["u"], v=data.net["v"], T=data.net["T"], S=data.net["S"]))
        self.atmosphere_ocean_interface = _Named(fluxes=_Named(**{k: data.ao[k] for k in (
            "latent_heat", "sensible_heat", "water_vapor", "x_momentum", "y_momentum")}),
            temperature=data.ao["interface_temperature"])
        if self.has_sea_ice:
            self.net_fluxes.sea_ice = _Named(top=_Named(heat=data.net_ice["top_heat"], u=data.net_ice["top_u"], v=data.net_ice["top_v"]),
                                             bottom=_Named(heat=data.net_ice["bottom_heat"]))
            self.sea_ice_ocean_interface = _Named(fluxes=_Named(**data.io))
            self.atmosphere_sea_ice_interface = _Named(fluxes=_Named(**{k: data.ai[k] for k in (
                "latent_heat", "sensible_heat", "water_vapor", "x_momentum", "y_momentum")}),
                temperature=data.ai["interface_temperature"])
        self.exchange_atmosphere_state = _Named(**data.exchange)


def build_coupled_model(ocean, sea_ice, atmosphere, radiation, land, flux_configuration,
                        velocity_formulation="relative", ocean_minimum_salinity=1, **kw):
    """build_coupled_model (omip_simulation.jl:123-164): same options, same error strings."""
    if flux_configuration == "default":
        interfaces = ComponentInterfaces(atmosphere, ocean, sea_ice, radiation=radiation, land=land,
                                         ocean_minimum_salinity=ocean_minimum_salinity, **kw)
        return OceanSeaIceModel(ocean, sea_ice, atmosphere=atmosphere, land=land, interfaces=interfaces)
    if velocity_formulation == "relative":
        vd = RelativeVelocity()
    elif velocity_formulation == "wind":
        vd = WindVelocity()
    else:
        raise ValueError(f"Unknown velocity_formulation: {velocity_formulation}. Options: :relative, :wind")
    if flux_configuration == "corrected":
        ao, ai = corrected_atmosphere_ocean_fluxes(), corrected_atmosphere_sea_ice_fluxes()
    elif flux_configuration == "ncar":
        ao, ai = ncar_atmosphere_ocean_fluxes(), ncar_atmosphere_sea_ice_fluxes()
    else:
        raise ValueError(f"Unknown flux_configuration: {flux_configuration}. Options: :default, :corrected, :ncar")
    interfaces = ComponentInterfaces(atmosphere, ocean, sea_ice, radiation=radiation, land=land,
                                     atmosphere_ocean_fluxes=ao, atmosphere_sea_ice_fluxes=ai,
                                     sea_ice_ocean_heat_flux=corrected_ice_ocean_heat_flux(),
                                     atmosphere_ocean_velocity_difference=vd, atmosphere_sea_ice_velocity_difference=vd,
                                     ocean_minimum_salinity=ocean_minimum_salinity, **kw)
    return OceanSeaIceModel(ocean, sea_ice, atmosphere=atmosphere, land=land, interfaces=interfaces)


class OceanSeaIceModel:
    """OceanSeaIceModel(ocean, sea_ice; atmosphere, radiation | interfaces) — fields .ocean .sea_ice
    .atmosphere .interfaces .clock (src/ClimaOcean.jl:56-57).  The constructor ends with an initial
    update_state! like the reference."""

    def __init__(self, ocean, sea_ice=None, atmosphere=None, radiation=None, land=None, interfaces=None, clock=None,
                 stream=None):
        if atmosphere is None:
            raise ValueError("OceanSeaIceModel requires a prescribed atmosphere")
        self.ocean, self.sea_ice, self.atmosphere, self.land = ocean, sea_ice, atmosphere, land
        self.interfaces = interfaces or ComponentInterfaces(atmosphere, ocean, sea_ice, radiation=radiation, land=land)
        self.clock = clock or Clock()
        self.stream = stream
        self.last_dt = 1.0
        update_state(self)


def OceanOnlyModel(ocean, atmosphere=None, **kw):
    return OceanSeaIceModel(ocean, None, atmosphere=atmosphere, **kw)


def update_state(model):
    """update_state!(model): recompute every interface flux from the current component states
    (SURVEY §3.2).  Ocean-only: the fused interpolate+solve+assemble path (2 launches).  With sea
    ice: interpolate → atmosphere–ocean → atmosphere–sea-ice → sea-ice–ocean → net ocean assembly."""
    itf = model.interfaces
    d, eng, t = itf.data, itf.engine, model.clock.time
    if not itf.has_sea_ice:
        inp, out = d.update_bundles(with_land=itf.land is not None)
        eng.update_state(inp, out, t, model.stream)
        return
    series, xch, ocean = d.atmos_series(), d.exchange_state(), d.ocean_surface()
    ice, io = d.sea_ice_state(), d.ice_ocean_fluxes()
    ao, ai, net = d.interface_fluxes("ao"), d.interface_fluxes("ai"), d.net_ocean_fluxes()
    eng.interpolate_atmosphere_state(series, t, xch, model.stream)
    if itf.land is not None:
        eng.interpolate_land(d.land_series(), t, xch, model.stream)
    eng.compute_atmosphere_ocean_fluxes(xch, ocean, ao, model.stream)
    eng.compute_atmosphere_sea_ice_fluxes(xch, ocean, ice, ai, model.stream)
    eng.compute_sea_ice_ocean_fluxes(d.ocean_columns(), ice, model.last_dt, io, model.stream)
    eng.compute_net_ocean_fluxes(xch, ocean, ao, ice, io, net, model.stream)
    eng.compute_net_sea_ice_fluxes(xch, ocean, ice, ai, io, d.net_sea_ice_fluxes(), model.stream)


def time_step(model, dt):
    """time_step!(model, Δt): step sea ice → step ocean → tick clock → update_state!."""
    if model.sea_ice is not None:
        model.sea_ice.time_step(dt)
    model.ocean.time_step(dt)
    model.clock.time += dt
    model.clock.iteration += 1
    model.last_dt = dt
    update_state(model)


class NormalizeSalinity:
    """NormalizeSalinity(flux_field, additional_fluxes, additional_buffer, mean_total) — the callback of
    /root/reference/src/OMIPConfigurations/omip_simulation.jl:187-220: at each call subtract the global (area-weighted)
    mean of the combined surface salinity flux from the bulk-flux Field, over its whole parent, so that the global salt
    budget integrates to zero.  `additional_fluxes`, when given, is a callable `(buffer_field, sim)` that materialises the
    additional flux (e.g. a surface restoring) into `additional_buffer` — the reference's `_materialize_top_flux!` launch.
    Multi-GPU: pass `dist`/`world`; the slabs' partial sums are all-reduced (climaocean.jl_b200/slabs.py)."""

    def __init__(self, data, engine, additional_fluxes=None, dist=None, world=1):
        self.data, self.engine = data, engine
        self.flux_field = data.net["S"]
        self.additional_fluxes = additional_fluxes
        self.additional_buffer = None
        if additional_fluxes is not None:
            self.additional_buffer = self.flux_field.clone()
            self.additional_buffer.data.zero_() if hasattr(self.additional_buffer.data, "zero_") else self.additional_buffer.data.fill(0)
        self.dist, self.world = dist, world
        self.mean_total = None          # device tensor (Σ f·Az, Σ Az) of the last call

    def __call__(self, sim=None, stream=None):
        from . import slabs
        if self.additional_fluxes is not None:
            self.additional_fluxes(self.additional_buffer, sim)
        norm = self.data.salinity_normalization(self.additional_buffer)
        self.mean_total = slabs.normalize_salinity_flux(self.engine, norm, self.dist, self.world, stream)
        return None


def salinity_normalizer(model, additional_fluxes=None, dist=None, world=1):
    """salinity_normalizer(bc) of the reference (omip_simulation.jl:193-206), for a coupled model built here."""
    itf = model.interfaces
    return NormalizeSalinity(itf.data, itf.engine, additional_fluxes, dist, world)

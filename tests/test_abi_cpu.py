"""The C-ABI library loads without a GPU, exports every symbol include/coflux.h declares, agrees
with the ctypes mirror on every struct size, and validates parameters like the reference's
constructors do (same option names / error texts).  No compute calls: CPU only."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import climaocean.jl_b200 as cj
from climaocean.jl_b200 import _abi
from oracle import pyoracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = _abi.load_library()


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "coflux.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(coflux_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_abi.EXPORTS), declared ^ set(_abi.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.coflux_abi_version() == _abi.ABI_VERSION
    assert b"sm_100a" in lib.coflux_build_info()


def test_struct_sizes_match_the_compiled_header():
    for name, cls in _abi.STRUCTS.items():
        assert lib.coflux_sizeof(name.encode()) == C.sizeof(cls), name
    assert lib.coflux_sizeof(b"nonsense") == -1


def test_default_config_values():
    cfg = cj.default_config(64, 32, 8)
    ao = cfg.atmosphere_ocean
    assert (cfg.grid.Nx, cfg.grid.Ny, cfg.grid.Nz, cfg.grid.ring) == (64, 32, 8, 1)
    assert ao.stability_functions == _abi.STABILITY_EDSON and ao.similarity_form == _abi.PROFILE_LOGARITHMIC
    assert ao.tolerance == 1e-8 and ao.max_iterations == 100 and ao.von_karman_constant == 0.4
    assert ao.momentum_roughness.gravity_wave_parameter == 0.02          # omip_simulation.jl:263
    assert cfg.ocean.reference_density == 1026.0 and cfg.ocean.heat_capacity == 3991.86795711963   # visualize/common.jl:17-18
    assert cfg.radiation.ocean_albedo == 0.06 and cfg.radiation.ocean_emissivity == 1.0             # atmosphere.jl:43
    assert cfg.ocean.minimum_salinity == 1.0                                                        # omip_simulation.jl:125


def test_flux_configurations_follow_omip_simulation():
    c = cj.default_config(8, 8, 2, flux_configuration="corrected")
    ao, ai = c.atmosphere_ocean, c.atmosphere_sea_ice
    assert ao.similarity_form == _abi.PROFILE_COARE_LOGARITHMIC and ao.minimum_gustiness == 0.5       # :43-44
    assert ao.momentum_roughness.wave_formulation == _abi.WAVES_WIND_DEPENDENT                          # :46
    assert ao.momentum_roughness.viscosity.kind == _abi.VISCOSITY_TEMPERATURE_POLY                       # :47
    assert ai.stability_functions == _abi.STABILITY_SHEBA_PAULSON and ai.minimum_gustiness == 0.2      # :64-66
    assert (ai.momentum_roughness.fixed_length, ai.temperature_roughness.fixed_length) == (5e-4, 5e-5)  # :67-68
    assert c.ice_ocean.heat_flux == _abi.ICE_OCEAN_THREE_EQUATION                                        # :77
    assert c.ice_ocean.friction_velocity == _abi.FRICTION_VELOCITY_MOMENTUM_BASED
    n = cj.default_config(8, 8, 2, flux_configuration="ncar", velocity_formulation="wind")
    assert n.atmosphere_ocean.formulation == _abi.FLUXES_COEFFICIENT_LARGE_YEAGER                        # :86-89
    assert n.atmosphere_ocean.stop_kind == _abi.STOP_FIXED_ITERATIONS and n.atmosphere_ocean.max_iterations == 5
    assert n.atmosphere_sea_ice.gustiness_parameter == 0.0 and n.atmosphere_sea_ice.minimum_gustiness == 0.5   # :109-110
    assert n.atmosphere_sea_ice.water_vapor_roughness.fixed_length == 5e-4                                     # :113
    assert n.atmosphere_ocean.velocity_formulation == _abi.VELOCITY_WIND                                       # :135-137
    with pytest.raises(cj.CofluxError, match="Unknown flux_configuration: shear_aware. Options: default, corrected, ncar"):
        cj.default_config(8, 8, 2, flux_configuration="shear_aware")                                           # :159-160
    with pytest.raises(ValueError, match="Unknown velocity_formulation"):
        cj.default_config(8, 8, 2, flux_configuration="corrected", velocity_formulation="absolute")
    # `:default` returns before velocity_formulation is looked at (omip_simulation.jl:127-133): never validated, never applied
    d = cj.default_config(8, 8, 2, flux_configuration="default", velocity_formulation="wind")
    assert d.atmosphere_ocean.velocity_formulation == _abi.VELOCITY_RELATIVE
    assert cj.default_config(8, 8, 2, velocity_formulation="absolute").atmosphere_sea_ice.velocity_formulation == _abi.VELOCITY_RELATIVE
    lib = cj.load_library()
    import ctypes
    assert lib.coflux_apply_flux_configuration(ctypes.byref(d), b"default", 7) == 0      # the C entry agrees
    assert d.atmosphere_ocean.velocity_formulation == _abi.VELOCITY_RELATIVE
    assert lib.coflux_apply_flux_configuration(ctypes.byref(d), b"ncar", 7) == _abi.ERR_INVALID_ARGUMENT


def test_python_parameter_objects_reproduce_the_presets():
    """The mirrored constructors (SimilarityTheoryFluxes(...) etc.) and the C presets must agree byte for byte."""
    from climaocean.jl_b200 import models as m
    c = cj.default_config(8, 8, 2, flux_configuration="corrected")
    assert bytes(m.corrected_atmosphere_ocean_fluxes().to_params()) == bytes(c.atmosphere_ocean)
    ai = m.corrected_atmosphere_sea_ice_fluxes().to_params()
    ai.interface_temperature = _abi.TEMPERATURE_SKIN
    assert bytes(ai) == bytes(c.atmosphere_sea_ice)
    n = cj.default_config(8, 8, 2, flux_configuration="ncar")
    assert bytes(m.ncar_atmosphere_ocean_fluxes().to_params()) == bytes(n.atmosphere_ocean)
    ai = m.ncar_atmosphere_sea_ice_fluxes().to_params()
    ai.interface_temperature = _abi.TEMPERATURE_SKIN
    assert bytes(ai) == bytes(n.atmosphere_sea_ice)
    assert bytes(m.SimilarityTheoryFluxes(momentum_roughness_length=m.MomentumRoughnessLength(wave_formulation=0.02)).to_params()) == \
        bytes(cj.default_config(8, 8, 2).atmosphere_ocean)


@pytest.mark.parametrize("mutate,status,fragment", [
    (lambda c: setattr(c, "abi_version", 99), _abi.ERR_INVALID_ARGUMENT, "abi_version"),
    (lambda c: setattr(c, "dtype", 16), _abi.ERR_INVALID_ARGUMENT, "dtype"),
    (lambda c: setattr(c.grid, "Nx", 0), _abi.ERR_INVALID_ARGUMENT, "grid size"),
    (lambda c: setattr(c.grid, "ring", 3), _abi.ERR_INVALID_ARGUMENT, "ring"),
    (lambda c: setattr(c.atmosphere_ocean, "stability_functions", 17), _abi.ERR_INVALID_ARGUMENT, "stability"),
    (lambda c: setattr(c.atmosphere_ocean, "tolerance", float("nan")), _abi.ERR_INVALID_ARGUMENT, "non-finite"),
    (lambda c: setattr(c.atmosphere_ocean, "turbulent_prandtl_number", 0.9), _abi.ERR_UNSUPPORTED, "prandtl"),
    (lambda c: setattr(c.atmosphere_ocean, "interface_temperature", _abi.TEMPERATURE_SKIN), _abi.ERR_UNSUPPORTED, "bulk"),
    (lambda c: setattr(c.atmosphere_sea_ice.momentum_roughness, "fixed_length", 0.0), _abi.ERR_INVALID_ARGUMENT, "roughness"),
    (lambda c: setattr(c.ice_ocean, "heat_flux", 5), _abi.ERR_INVALID_ARGUMENT, "heat_flux"),
    (lambda c: setattr(c.ocean, "reference_density", -1.0), _abi.ERR_INVALID_ARGUMENT, "ocean properties"),
    (lambda c: setattr(c.atmosphere.thermodynamics, "gas_constant", float("inf")), _abi.ERR_INVALID_ARGUMENT, "thermodynamics"),
])
def test_create_rejects_bad_parameters_before_touching_the_device(mutate, status, fragment):
    cfg = cj.default_config(8, 8, 2)
    mutate(cfg)
    ctx = C.c_void_p()
    rc = lib.coflux_create(C.byref(ctx), C.byref(cfg))
    assert rc == status and not ctx.value
    assert fragment.lower() in lib.coflux_last_error().decode().lower()


def test_create_without_a_device_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    ctx = C.c_void_p()
    cfg = cj.default_config(8, 8, 2)
    assert lib.coflux_create(C.byref(ctx), C.byref(cfg)) == _abi.ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.coflux_last_error()
    with pytest.raises(cj.CofluxError):
        cj.Engine(cfg)


def test_null_arguments_are_errors_not_crashes():
    assert lib.coflux_create(None, None) == _abi.ERR_INVALID_ARGUMENT
    assert lib.coflux_update_state(None, None, None, 0.0, None) == _abi.ERR_INVALID_ARGUMENT
    assert lib.coflux_interpolate_atmosphere(None, None, 0.0, None, None) == _abi.ERR_INVALID_ARGUMENT
    assert lib.coflux_destroy(None) == _abi.OK


def test_time_indices_agree_with_the_oracle_and_validate():
    ora = pyoracle.load()
    rng = np.random.default_rng(7)
    times = np.cumsum(rng.uniform(100.0, 20000.0, size=17))
    tp = times.ctypes.data_as(C.POINTER(C.c_double))
    a = [C.c_int32(), C.c_int32(), C.c_double()]
    b = [C.c_int32(), C.c_int32(), C.c_double()]
    for mode in (_abi.TIME_LINEAR, _abi.TIME_CYCLICAL, _abi.TIME_CLAMP):
        for t in np.concatenate([rng.uniform(times[0] - 5e4, times[-1] + 5e4, 200), times]):
            assert lib.coflux_time_indices(tp, 17, mode, 0.0, float(t), *[C.byref(x) for x in a]) == 0
            assert ora.oracle_time_indices(tp, 17, mode, 0.0, float(t), *[C.byref(x) for x in b]) == 0
            assert (a[0].value, a[1].value) == (b[0].value, b[1].value) and a[2].value == b[2].value
    bad = np.array([0.0, 2.0, 1.0])
    assert lib.coflux_time_indices(bad.ctypes.data_as(C.POINTER(C.c_double)), 3, 0, 0.0, 0.5, *[C.byref(x) for x in a]) == _abi.ERR_INVALID_ARGUMENT
    assert lib.coflux_time_indices(tp, 17, 9, 0.0, 0.5, *[C.byref(x) for x in a]) == _abi.ERR_INVALID_ARGUMENT


def test_field_descriptors_describe_oceananigans_parents():
    grid = cj.LatitudeLongitudeGrid((10, 6, 4), halo=(3, 2, 1))
    f = cj.Field.zeros(grid.size, grid.halo, np.float64)
    a = f.array()
    assert (a.stride_i, a.stride_j, a.stride_k) == (1, 16, 16 * 10) and (a.off_i, a.off_j, a.off_k) == (3, 2, 1)
    f.interior[...] = 1.0
    assert f.data.sum() == 10 * 6 * 4
    FI, FJ = cj.fractional_indices(grid, 640, 320, ring=1)
    assert FI.shape == (1, 8, 12) and FI.min() >= 0.0 and FI.max() < 640.0


def test_oracle_side_defaults_equal_the_library_defaults_byte_for_byte():
    """bench.py's CPU arm builds its configuration on the oracle side (it must not load the product library); the two
    statements of the reference defaults (omip_simulation.jl:40-164) must never drift apart."""
    import ctypes
    from oracle import pyoracle
    for name in ("default", "corrected", "ncar"):
        for vel in (_abi.VELOCITY_RELATIVE, _abi.VELOCITY_WIND):
            for dtype in (_abi.F64, _abi.F32):
                a = cj.default_config(12, 7, 3, dtype, name, "relative" if vel == _abi.VELOCITY_RELATIVE else "wind")
                b = pyoracle.default_config(_abi.Config(), 12, 7, 3, dtype, name, vel)
                assert bytes(a) == bytes(b), (name, vel, dtype)
    assert pyoracle.load().oracle_apply_flux_configuration(ctypes.byref(_abi.Config()), b"shear_aware", 0) != 0

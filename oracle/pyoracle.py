"""ctypes loader for the CPU oracle (oracle/libcoflux_oracle.so).

TEST INFRASTRUCTURE — PARITY UNPINNED (see oracle/coflux_oracle.c).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this module.
The oracle consumes the same ctypes bundles as the CUDA library, pointing at HOST (numpy) arrays.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcoflux_oracle.so")
_lib = None


def use_fast_build():
    """Switch this process to the CPU-BASELINE build of the oracle (-O3 -march=native, built here and now for this host's
    cores: oracle/Makefile `fast`).  For bench.py's timed CPU legs only; the parity tests keep the -O2 -ffp-contract=off build."""
    global LIB_PATH, _lib
    subprocess.run(["make", "-C", _HERE, "-B", "fast"], check=True, stdout=subprocess.DEVNULL)
    LIB_PATH = os.path.join(_HERE, "libcoflux_oracle_fast.so")
    _lib = None
    return load(build_if_missing=False)


def default_config(cfg, Nx, Ny, Nz, dtype, flux_configuration="default", velocity=0):
    """Fill a coflux_config with the reference defaults + a build_coupled_model flux configuration — on the oracle side, so
    that the CPU arm of bench.py does not load the product library."""
    lib = load()
    assert lib.oracle_default_config(C.byref(cfg), int(Nx), int(Ny), int(Nz), int(dtype)) == 0
    assert lib.oracle_apply_flux_configuration(C.byref(cfg), flux_configuration.encode(), int(velocity)) == 0
    return cfg


def load(build_if_missing=True):
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH) and build_if_missing:
        subprocess.run(["make", "-C", _HERE], check=True, stdout=subprocess.DEVNULL)
    _lib = C.CDLL(LIB_PATH)
    _lib.oracle_time_indices.argtypes = [C.POINTER(C.c_double), C.c_int32, C.c_int32, C.c_double, C.c_double,
                                         C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_double)]
    return _lib


def _suffix(cfg):
    return "_f64" if cfg.dtype == 64 else "_f32"


def fn(name, cfg):
    return getattr(load(), name + _suffix(cfg))


def interpolate_atmosphere(cfg, series, time, exchange):
    f = fn("oracle_interpolate_atmosphere", cfg)
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
    assert f(C.byref(cfg), C.byref(series), float(time), C.byref(exchange)) == 0


def atmosphere_ocean_fluxes(cfg, exchange, ocean, fluxes):
    f = fn("oracle_atmosphere_ocean_fluxes", cfg)
    assert f(C.byref(cfg), C.byref(exchange), C.byref(ocean), C.byref(fluxes)) == 0


def atmosphere_sea_ice_fluxes(cfg, exchange, ocean, ice, fluxes):
    f = fn("oracle_atmosphere_sea_ice_fluxes", cfg)
    assert f(C.byref(cfg), C.byref(exchange), C.byref(ocean), C.byref(ice), C.byref(fluxes)) == 0


def sea_ice_ocean_fluxes(cfg, columns, ice, dt, fluxes):
    f = fn("oracle_sea_ice_ocean_fluxes", cfg)
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
    assert f(C.byref(cfg), C.byref(columns), C.byref(ice), float(dt), C.byref(fluxes)) == 0


def assemble_net_ocean_fluxes(cfg, exchange, ocean, ao, ice, io, net):
    f = fn("oracle_assemble_net_ocean_fluxes", cfg)
    assert f(C.byref(cfg), C.byref(exchange), C.byref(ocean), C.byref(ao), C.byref(ice) if ice is not None else None,
             C.byref(io) if io is not None else None, C.byref(net)) == 0


def update_state(cfg, inputs, outputs, time):
    f = fn("oracle_update_state", cfg)
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double]
    assert f(C.byref(cfg), C.byref(inputs), C.byref(outputs), float(time)) == 0


def closure_surface_forcing(cfg, net, forcing):
    """KPP / NEMO-TKE surface-forcing front ends (kpp_surface_forcing.jl:18-29, nemo_tke_surface_forcing.jl:14-22)."""
    f = fn("oracle_closure_surface_forcing", cfg)
    assert f(C.byref(cfg), C.byref(net), C.byref(forcing)) == 0


def normalize_salinity_flux(cfg, norm, mean=None):
    """NormalizeSalinity (omip_simulation.jl:187-220) on host arrays; returns (Σ f·Az, Σ Az).  `mean`: subtract this
    value instead of the local mean (multi-slab)."""
    f = fn("oracle_normalize_salinity_flux", cfg)
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    sums = (C.c_double * 2)()
    m = C.c_double(mean) if mean is not None else None
    assert f(C.byref(cfg), C.byref(norm), sums, C.byref(m) if m is not None else None) == 0
    return sums[0], sums[1]


def interpolate_land(cfg, land, time, exchange):
    f = fn("oracle_interpolate_land", cfg)
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
    assert f(C.byref(cfg), C.byref(land), float(time), C.byref(exchange)) == 0


def assemble_net_sea_ice_fluxes(cfg, exchange, ocean, ice, ai, io, out):
    f = fn("oracle_assemble_net_sea_ice_fluxes", cfg)
    assert f(C.byref(cfg), C.byref(exchange), C.byref(ocean) if ocean is not None else None, C.byref(ice), C.byref(ai),
             C.byref(io) if io is not None else None, C.byref(out)) == 0


def accumulate_flux_averages(cfg, net, ao, ice, io, averages):
    f = fn("oracle_accumulate_flux_averages", cfg)
    assert f(C.byref(cfg), C.byref(net), C.byref(ao) if ao is not None else None, C.byref(ice) if ice is not None else None,
             C.byref(io) if io is not None else None, C.byref(averages)) == 0


def set_threads(n):
    """Limit / set OpenMP threads of the oracle through libgomp's omp_set_num_threads."""
    gomp = C.CDLL("libgomp.so.1")
    gomp.omp_set_num_threads(int(n))


def max_threads():
    gomp = C.CDLL("libgomp.so.1")
    return int(gomp.omp_get_max_threads())

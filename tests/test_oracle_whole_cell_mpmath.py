"""The C oracle against a 50-digit mpmath evaluation of the WHOLE per-cell fixed point (tests/mp_cell.py) — every formula
of SURVEY Appendix A1–A7 chained through the iteration with its stop rule — for the three OMIP flux configurations over the
ocean and the three sea-ice parameter sets with the skin temperature (/root/reference/src/OMIPConfigurations/
omip_simulation.jl:40-113).  Pins "oracle = Appendix A evaluated exactly" to 1e-13; it does not (cannot, here) pin the
appendix to NumericalEarth.  CPU only; 3 × 96 ocean cells + 3 × 64 ice cells keep the run under two minutes."""
import ctypes as C

import numpy as np
import pytest

import climaocean.jl_b200 as cj
from climaocean.jl_b200 import _abi
from oracle import pyoracle
from tests import mp_cell

lib = pyoracle.load()
# Bars (relative, denominator floored at 1e-3·max|field|).  u★ and θ★ sit at 1e-13 and better.  q★ = χ_q·(q_a − q_s) inherits the
# conditioning of Δq: q_a and q_s agree to two or three digits in many cells and p_sat(T) itself cancels three digits in
# 1/T_tr − 1/T, so a Float64 evaluation of the SAME formulas carries up to ≈ 1e3 ulp ≈ 1.4e-13 there (measured worst case on
# these cells: 1.35e-13); it is held to 5e-13.  Iteration counts must agree exactly.
RTOL = {0: 1e-13, 1: 1e-13, 2: 5e-13, 3: 1e-13}


def _cells(n, seed, ice=False):
    r = np.random.default_rng(seed)
    c = dict(ua=r.uniform(-25, 25, n), va=r.uniform(-25, 25, n), Ta=r.uniform(250, 305, n), pa=r.uniform(9.6e4, 1.04e5, n),
             qa=r.uniform(1e-4, 2e-2, n), us=r.uniform(-1, 1, n), vs=r.uniform(-1, 1, n))
    if ice:
        c.update(Ts=r.uniform(243.15, 272.65, n), Qs=r.uniform(0, 1000, n), Ql=r.uniform(100, 450, n), h_ice=r.uniform(0.02, 3, n),
                 S_ice=r.uniform(2, 8, n), albedo=r.uniform(0.5, 0.85, n), So=np.zeros(n))
    else:
        c.update(Ts=r.uniform(271.35, 303.15, n), So=r.uniform(30, 38, n))
    # a few special cells: calm, neutral, light wind / strongly unstable
    c["ua"][0], c["va"][0], c["us"][0], c["vs"][0] = 0.0, 0.0, 0.0, 0.0
    c["ua"][1], c["va"][1] = 0.3, -0.1
    c["Ta"][1], c["Ts"][1] = (262.0, 268.0) if ice else (283.0, 301.0)
    return c


def _close(a, b, scale, j, floor=1e-3):
    return abs(a - b) <= RTOL[j] * max(abs(b), floor * scale)


@pytest.mark.parametrize("flux_configuration", ["default", "corrected", "ncar"])
def test_atmosphere_ocean_cell_solve_matches_50_digit_evaluation(flux_configuration):
    cfg = cj.default_config(4, 4, 2, 64, flux_configuration)
    n = 96
    cells = _cells(n, 11)
    out = np.zeros((n, 4))
    for k in range(n):
        a = np.array([cells[x][k] for x in ("ua", "va", "Ta", "pa", "qa", "us", "vs", "Ts", "So")])
        lib.oracle_probe_solve_f64(C.byref(cfg), a.ctypes.data_as(C.c_void_p), out[k].ctypes.data_as(C.c_void_p))
    scale = np.abs(out[:, :3]).max(axis=0)
    worst = 0.0
    for k in range(n):
        us, ts, qs, _, it = mp_cell.solve_cell(cfg, cfg.atmosphere_ocean, {x: cells[x][k] for x in cells}, 0)
        assert it == int(out[k, 3]), (k, it, out[k, 3])                   # same iteration path
        for j, ref in enumerate((us, ts, qs)):
            assert _close(out[k, j], float(ref), scale[j], j), (flux_configuration, k, j, out[k, j], float(ref))
            worst = max(worst, abs(out[k, j] - float(ref)) / max(abs(float(ref)), 1e-3 * scale[j]))
    print(flux_configuration, "worst relative deviation from the 50-digit fixed point:", worst)


@pytest.mark.parametrize("flux_configuration", ["default", "corrected", "ncar", "default+linearized_longwave"])
def test_atmosphere_sea_ice_cell_solve_matches_50_digit_evaluation(flux_configuration):
    linearized = flux_configuration.endswith("+linearized_longwave")
    cfg = cj.default_config(4, 4, 2, 64, flux_configuration.split("+")[0])
    if linearized:
        cfg.atmosphere_sea_ice.skin_temperature_update = _abi.SKIN_LINEARIZED_LONGWAVE
    n = 64
    cells = _cells(n, 23, ice=True)
    out = np.zeros((n, 5))
    for k in range(n):
        a = np.array([cells[x][k] for x in ("ua", "va", "Ta", "pa", "qa", "us", "vs", "Ts", "Qs", "Ql", "h_ice", "S_ice", "albedo")])
        lib.oracle_probe_solve_ice_f64(C.byref(cfg), a.ctypes.data_as(C.c_void_p), out[k].ctypes.data_as(C.c_void_p))
    scale = np.abs(out[:, :4]).max(axis=0)
    maxit = cfg.atmosphere_sea_ice.max_iterations
    checked = cycles = 0
    for k in range(n):
        us, ts, qs, Ts, it = mp_cell.solve_cell(cfg, cfg.atmosphere_sea_ice, {x: cells[x][k] for x in cells}, 1)
        if it >= maxit or int(out[k, 4]) >= maxit:
            # the clamped skin-temperature update is not a contraction for thick ice: such a cell never meets the stop rule and
            # ends on a limit cycle, where 1e-16 differences decide the phase — both evaluations must at least agree on THAT
            assert it >= maxit and int(out[k, 4]) >= maxit, (k, it, out[k, 4])
            cycles += 1
            continue
        assert it == int(out[k, 4]), (k, it, out[k, 4])
        for j, ref in enumerate((us, ts, qs, Ts)):
            # over ice θ★ = χ_θ (θ_a − T_s) with an ITERATED T_s ≈ 270 K carried to ulp(T_s) = 5.7e-14 K: a cell whose skin
            # temperature has relaxed to the air temperature has θ★ → 0 with an absolute error floor of χ·ulp(T_s) ≈ 2e-15,
            # hence the denominator floor of 1e-2·max|field| here (1e-3 over the ocean, where T_s is an input)
            assert _close(out[k, j], float(ref), scale[j], j, floor=1e-2), (flux_configuration, k, j, out[k, j], float(ref))
        checked += 1
    print(flux_configuration, f"{checked} converged ice cells agree to {RTOL}; {cycles} of {n} run to maxiter (limit cycle) in both")
    assert checked >= n // 2

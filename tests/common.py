"""Shared helpers of the parity tests: run the CPU oracle (the checker) and the CUDA library (the
product, through the C ABI) on the same seeded synthetic inputs and compare every output field."""
import numpy as np

import climaocean.jl_b200 as cj
from climaocean.jl_b200 import _abi
from oracle import pyoracle

QUERY_TIME = 1.37 * 3 * 3600.0   # exercises the time weights

# Tolerances of BASELINE.json:north_star: ≤1e-12 relative in Float64, ≤1e-5 in Float32.
# Relative error is |a−b| / max(|b|, FLOOR·max|b|).  Every flux is proportional to a difference of
# nearly equal inputs (Δθ = θ_a − T_s, Δq = q_a − q_s, signed sums in the assembly), so a result near
# zero carries an ABSOLUTE rounding error of a few ulp of the operands — i.e. a few ulp of the field's
# own scale — no matter who computes it (measured: ≤ 8e-16·max|b| in Float64, ≤ 5e-7·max|b| in Float32
# between oracle and CUDA, tests/diag/parity_report.py).  The denominator is therefore floored at a
# fraction of the field's largest magnitude: 1e-3 in Float64 (absolute error ≤ 1e-15·max|b| ≈ 9 ulp)
# and 1e-1 in Float32 (≤ 1e-6·max|b| ≈ 17 ulp).  Stated once here, used by every parity test.
RTOL = {64: 1e-12, 32: 1e-5}
FLOOR = {64: 1e-3, 32: 1e-1}


def rel_err(a, b, bits=64):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if np.isnan(a).any() or np.isnan(b).any():
        return float("nan")
    scale = np.max(np.abs(b)) if b.size else 0.0
    if scale == 0.0:
        return float(np.max(np.abs(a))) if a.size else 0.0
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), FLOOR[bits] * scale)))


def np_dtype(bits):
    return np.float64 if bits == 64 else np.float32


def make_case(Nx, Ny, Nz, bits=64, latitude=(-60.0, 60.0), flux_configuration="default", velocity="relative",
              halo=(7, 7, 7), ring=1, **synth_kw):
    grid = cj.LatitudeLongitudeGrid((Nx, Ny, Nz), latitude=latitude, halo=halo, dtype=np_dtype(bits))
    host = cj.SurfaceFluxData.synthetic(grid, ring=ring, **synth_kw)
    cfg = cj.default_config(Nx, Ny, Nz, bits, flux_configuration, velocity)
    if flux_configuration == "default" and velocity == "wind":
        # build_coupled_model ignores velocity_formulation for `:default` (omip_simulation.jl:127-133); a user gets wind
        # velocities there through ComponentInterfaces(...; atmosphere_ocean_velocity_difference = WindVelocity())
        cfg.atmosphere_ocean.velocity_formulation = _abi.VELOCITY_WIND
        cfg.atmosphere_sea_ice.velocity_formulation = _abi.VELOCITY_WIND
    cfg.grid.ring = ring
    return grid, host, cfg


def oracle_update(host, cfg, time=QUERY_TIME, with_ice_terms=False):
    inp, out = host.update_bundles(with_ice_terms)
    pyoracle.update_state(cfg, inp, out, time)
    return host.outputs()


def gpu_update(host, cfg, time=QUERY_TIME, with_ice_terms=False, device="cuda:0"):
    import torch
    dev = host.to(device)
    eng = cj.Engine(cfg)
    inp, out = dev.update_bundles(with_ice_terms)
    eng.update_state(inp, out, time)
    torch.cuda.synchronize()
    res = dev.outputs()
    res["_iterations"] = dev.iterations.numpy()
    res["_launches"] = eng.launches
    eng.close()
    return res, dev


def compare(gpu, ref, bits, keys=None, rtol=None):
    rtol = rtol or RTOL[bits]
    worst = {}
    for k, v in ref.items():
        if keys is not None and k not in keys:
            continue
        if k not in gpu:
            continue
        worst[k] = rel_err(gpu[k], v, bits)
    bad = {k: e for k, e in worst.items() if not (e <= rtol)}
    assert not bad, f"parity failures (rtol {rtol}): {bad}"
    return worst

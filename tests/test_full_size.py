"""Parity at BASELINE.json's full sizes (1/4° 1440×600 and 1/12° 4320×1800) through properties that do not need
the CPU oracle to cover the whole grid:

  * whole-grid parity — the OpenMP oracle recomputes EVERY cell of the same full-grid inputs (7.8 M cells at 1/12°, a few
                    seconds on the box's host cores), for all three flux configurations and both precisions; the CUDA
                    result of the full-grid launch must match everywhere to the north_star tolerance and, in Float64, with
                    the same iteration count in every cell.  (The oracle runs on sub-grid views of the same host arrays
                    — every descriptor's row offset shifted — in blocks of rows, so a failure names its latitude.);
  * determinism   — two launches give bit-identical outputs;
  * decomposition — the pipelined HOST-buffer entry (row-chunked launches, `coflux_update_state_host`) reproduces the
                    single-launch device path bit for bit, i.e. a cell's result does not depend on how the grid is cut;
  * closure       — the net ocean heat flux satisfies Jᵀ·ρ₀c₀ = Q_u + Q_aℓ + Q_c + Q_v (+ Q_ts) cell by cell
                    (/root/reference/experiments/OMIPSimulations/visualize/cache.jl:359-361 recovers W m⁻² this way),
                    the iteration counts are within the stop rule, and every cell reached its fixed point.
"""
import numpy as np
import pytest

import climaocean.jl_b200 as cj
from climaocean.jl_b200 import _abi
from oracle import pyoracle
from tests.common import FLOOR, QUERY_TIME, RTOL, np_dtype, rel_err

pytestmark = pytest.mark.gpu

SIZES = {"quarter": (1440, 600, 10), "twelfth": (4320, 1800, 75)}
BLOCK = 300        # latitude rows per oracle call (the whole grid is covered, block by block)


def _case(res, bits, flux_configuration):
    Nx, Ny, Nz = SIZES[res]
    grid = cj.LatitudeLongitudeGrid((Nx, Ny, 1), latitude=(-75.0, 75.0), halo=(7, 7, 0), dtype=np_dtype(bits))
    host = cj.SurfaceFluxData.synthetic(grid, ring=1)
    dev = host.to_device_columns("cuda:0", Nz)
    cfg = cj.default_config(Nx, Ny, Nz, bits, flux_configuration)
    cfg.grid.ring = 1
    return grid, host, dev, cfg


def _shift_rows(struct, j0, skip=()):
    """Move the row origin of every array descriptor of a ctypes bundle by j0 rows (a sub-grid view of the same memory)."""
    for name, ctype in struct._fields_:
        if ctype is _abi.Array and name not in skip:
            a = getattr(struct, name)
            if a.ptr:
                a.off_j += j0


def _oracle_rows(host, cfg_full, j0, rows):
    """Oracle update_state on rows [j0, j0 + rows) of the full-grid inputs; returns name -> (rows, Nx) arrays."""
    cfg = _abi.Config.from_buffer_copy(cfg_full)
    cfg.grid.Ny = rows
    cfg.grid.Nz = 1
    inp, out = host.update_bundles()
    series = ("u", "v", "T", "q", "p", "Qs", "Ql", "rain", "snow")          # source-grid arrays: indexed through fi, fj
    _shift_rows(inp.atmosphere.contents, j0, skip=series)
    _shift_rows(inp.ocean.contents, j0)
    for b in (out.exchange.contents, out.atmosphere_ocean.contents, out.net_ocean.contents):
        _shift_rows(b, j0)
    pyoracle.update_state(cfg, inp, out, QUERY_TIME)
    res = {}                                   # only the rows of the window (host.outputs() would copy every whole field)
    for grp, d in (("exchange", host.exchange), ("ao", host.ao), ("net", host.net)):
        for n, f in d.items():
            Hx, Hy, _ = f.halo
            a = f.numpy()
            res[f"{grp}.{n}"] = a[0, Hy + j0:Hy + j0 + rows, Hx:a.shape[2] - Hx].copy()
    return res


@pytest.mark.parametrize("res,bits,flux_configuration", [("quarter", 64, "default"), ("quarter", 64, "corrected"), ("quarter", 32, "default"),
                                                         ("quarter", 64, "ncar"), ("twelfth", 64, "default"), ("twelfth", 64, "corrected"),
                                                         ("twelfth", 64, "ncar"), ("twelfth", 32, "default")])
def test_full_size_whole_grid_parity_determinism_closure(res, bits, flux_configuration):
    import torch
    grid, host, dev, cfg = _case(res, bits, flux_configuration)
    eng = cj.Engine(cfg)
    inp, out = dev.update_bundles()
    eng.update_state(inp, out, QUERY_TIME)
    torch.cuda.synchronize()
    first = dev.outputs()
    its = dev.iterations.numpy()[0, 7:-7, 7:-7].copy()

    # parity against the oracle on the same inputs, every row of the grid
    Ny = grid.Ny
    windows = [(j0, min(BLOCK, Ny - j0)) for j0 in range(0, Ny, BLOCK)]
    bad, checked = {}, 0
    scale = {k: float(np.max(np.abs(v))) for k, v in first.items()}
    for j0, rows in windows:
        ref = _oracle_rows(host, cfg, j0, rows)
        checked += rows
        for k, v in ref.items():
            if k in first and (k.startswith("exchange.") or k.startswith("ao.") or k.startswith("net.")):
                a, b = first[k][j0:j0 + rows].astype(np.float64), v.astype(np.float64)
                # denominator floor relative to the WHOLE field's scale (a single row's own maximum would be a moving target)
                e = float(np.max(np.abs(a - b) / np.maximum(np.abs(b), FLOOR[bits] * scale[k]))) if scale[k] > 0 else float(np.max(np.abs(a - b)))
                if not (e <= RTOL[bits]):
                    bad[(k, j0)] = e
        assert np.array_equal(host.iterations.numpy()[0, 7 + j0:7 + j0 + rows, 7:-7], its[j0:j0 + rows]) or bits == 32, f"iteration counts differ in rows {j0}…"
    assert not bad, f"parity failures at {res} f{bits} {flux_configuration}: {dict(list(bad.items())[:8])}"
    assert checked == Ny
    print(f"{res} f{bits} {flux_configuration}: all {checked} rows ({checked * grid.Nx} cells) checked against the oracle")

    # determinism: bit-identical relaunch
    eng.update_state(inp, out, QUERY_TIME)
    torch.cuda.synchronize()
    second = dev.outputs()
    for k in first:
        assert np.array_equal(first[k], second[k]), f"{k} differs between two launches"

    # closure of the heat flux assembly (ocean-only: ℵ = 0, no ice–ocean terms)
    oc = cfg.ocean
    rho0c0 = oc.reference_density * oc.heat_capacity
    SQ = (first["net.upwelling_longwave"].astype(np.float64) + first["net.downwelling_longwave"] + first["ao.sensible_heat"] + first["ao.latent_heat"])
    if not cfg.radiation.shortwave_penetrates:
        SQ = SQ + first["net.downwelling_shortwave"]
    tol = 4e-15 if bits == 64 else 4e-6
    terms = (np.abs(first["net.upwelling_longwave"]) + np.abs(first["net.downwelling_longwave"]) + np.abs(first["ao.sensible_heat"]) +
             np.abs(first["ao.latent_heat"]) + np.abs(first["net.downwelling_shortwave"])).astype(np.float64)
    assert np.max(np.abs(first["net.T"].astype(np.float64) * rho0c0 - SQ) / terms) <= tol

    # stop rule: 1 ≤ iterations ≤ maxiter everywhere; in Float64 every cell converged well before maxiter
    maxit = cfg.atmosphere_ocean.max_iterations
    assert its.min() >= 1 and its.max() <= maxit
    if bits == 64 and flux_configuration != "ncar":
        assert its.max() < maxit
    for k in ("ao.friction_velocity", "ao.temperature_scale", "ao.humidity_scale", "net.T", "net.S", "net.u", "net.v"):
        assert np.isfinite(first[k]).all(), k
    eng.close()


@pytest.mark.parametrize("res,bits", [("quarter", 64), ("twelfth", 64)])
def test_full_size_host_entry_is_bit_identical_to_device_path(res, bits):
    import torch
    grid, host, dev, cfg = _case(res, bits, "default")
    eng = cj.Engine(cfg)
    inp, out = dev.update_bundles()
    eng.update_state(inp, out, QUERY_TIME)
    torch.cuda.synchronize()
    ref = dev.outputs()
    eng.close()

    cfg_h = cj.default_config(grid.Nx, grid.Ny, 1, bits, "default")
    cfg_h.grid.ring = 1
    eng_h = cj.Engine(cfg_h)
    H = grid.halo[0]
    planes = {n: torch.from_numpy(np.ascontiguousarray(host.ocean[n].data[0])).pin_memory() for n in ("u", "v", "T", "S")}
    outs = {n: torch.empty_like(planes["u"]).pin_memory() for n in ("u", "v", "T", "S", "Qv", "Qc")}
    step = _abi.HostStep(planes["u"].data_ptr(), planes["v"].data_ptr(), planes["T"].data_ptr(), planes["S"].data_ptr(),
                         outs["u"].data_ptr(), outs["v"].data_ptr(), outs["T"].data_ptr(), outs["S"].data_ptr(),
                         outs["Qv"].data_ptr(), outs["Qc"].data_ptr(), H, 0)
    eng_h.update_state_host(dev.atmos_series(), step, QUERY_TIME)
    torch.cuda.synchronize()
    for n, key in (("u", "net.u"), ("v", "net.v"), ("T", "net.T"), ("S", "net.S"), ("Qv", "ao.latent_heat"), ("Qc", "ao.sensible_heat")):
        got = outs[n].numpy()[H:-H, H:-H]
        assert np.array_equal(got, ref[key]), f"{key}: row-chunked host entry differs from the single launch"
    eng_h.close()

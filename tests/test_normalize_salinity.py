"""NormalizeSalinity — the per-step salinity-flux post-processing of the reference
(/root/reference/src/OMIPConfigurations/omip_simulation.jl:187-220):

    compute!(n.mean_total)                     # Field(Average(flux_field [+ additional_buffer], dims=(1,2)))
    parent(n.flux_field) .-= n.mean_total      # whole parent, halos included

CPU: the oracle restatement against closed-form answers and a 50-digit evaluation; the 2-rank (gloo) slab logic.
GPU: the CUDA kernels against the oracle (same tolerance as the flux path), run-to-run bit reproducibility, the full
1/12° grid, and the reference-facing `NormalizeSalinity` callable.
"""
import os
import sys

import mpmath as mp
import numpy as np
import pytest

import climaocean.jl_b200 as cj
from climaocean.jl_b200 import _abi, slabs
from climaocean.jl_b200.fields import Field
from oracle import pyoracle
from tests.common import QUERY_TIME, RTOL, make_case, np_dtype, oracle_update

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _host_case(Nx=48, Ny=24, bits=64, land_fraction=0.0):
    grid, host, cfg = make_case(Nx, Ny, 2, bits, land_fraction=land_fraction)
    oracle_update(host, cfg)                       # fills net.S with a realistic salinity flux
    return grid, host, cfg


def _interior(f):
    a = f.numpy()
    Hx, Hy, _ = f.halo
    return a[0, Hy:a.shape[1] - Hy, Hx:a.shape[2] - Hx]


def test_constant_flux_is_removed_everywhere_including_halos():
    grid, host, cfg = _host_case()
    host.net["S"].data[...] = 3.25e-7
    pyoracle.normalize_salinity_flux(cfg, host.salinity_normalization())
    assert np.max(np.abs(host.net["S"].data)) <= 1e-22          # the WHOLE parent: `parent(flux_field) .-= mean`


def test_area_weighted_mean_matches_a_50_digit_evaluation():
    grid, host, cfg = _host_case(Nx=36, Ny=30)
    before = _interior(host.net["S"]).astype(np.float64).copy()
    parent_before = host.net["S"].data.copy()
    num, den = pyoracle.normalize_salinity_flux(cfg, host.salinity_normalization())
    Az = grid.horizontal_areas()[grid.halo[1]:-grid.halo[1]].astype(np.float64)
    mp.mp.dps = 50
    mnum = mp.fsum(mp.mpf(float(before[j, i])) * mp.mpf(float(Az[j])) for j in range(grid.Ny) for i in range(grid.Nx))
    mden = mp.fsum(mp.mpf(float(Az[j])) for j in range(grid.Ny)) * grid.Nx
    mean = float(mnum / mden)
    assert abs(num / den - mean) <= 1e-15 * np.max(np.abs(before))
    assert np.allclose(host.net["S"].data, parent_before - np.float64(num / den), rtol=0, atol=1e-19)
    # Az is R²Δλ(sin φ_n − sin φ_s): the areas of one longitude column add up to the spherical zone
    zone = 6371e3 ** 2 * np.deg2rad(360.0 / grid.Nx) * (np.sin(np.deg2rad(60.0)) - np.sin(np.deg2rad(-60.0)))
    assert abs(Az.sum() - zone) <= 1e-12 * zone


def test_antisymmetric_flux_has_zero_mean_on_a_symmetric_grid():
    grid, host, cfg = _host_case(Nx=16, Ny=20)
    H = grid.halo[1]
    sign = np.where(np.arange(-H, grid.Ny + H) < grid.Ny // 2, -1.0, 1.0)
    host.net["S"].data[...] = 2e-6 * sign[None, :, None]
    before = host.net["S"].data.copy()
    num, den = pyoracle.normalize_salinity_flux(cfg, host.salinity_normalization())
    assert abs(num / den) <= 1e-22 and np.allclose(host.net["S"].data, before, rtol=0, atol=1e-22)


def test_masked_cells_do_not_enter_the_average_and_additional_flux_does():
    grid, host, cfg = _host_case(Nx=40, Ny=24, land_fraction=0.3)
    wet = _interior(host.mask) != 0
    assert 0 < wet.sum() < wet.size
    S = _interior(host.net["S"]).astype(np.float64).copy()
    add = host.net["S"].clone()
    rng = np.random.default_rng(3)
    add.data[...] = rng.normal(0, 1e-7, add.data.shape)
    A = _interior(add).astype(np.float64)
    Az = grid.horizontal_areas()[grid.halo[1]:-grid.halo[1]].astype(np.float64)[:, None] * np.ones((1, grid.Nx))
    expect = ((S + A) * Az)[wet].sum() / Az[wet].sum()
    add_before = add.data.copy()
    num, den = pyoracle.normalize_salinity_flux(cfg, host.salinity_normalization(add))
    assert abs(num / den - expect) <= 1e-14 * np.max(np.abs(S + A))
    assert np.array_equal(add.data, add_before)                 # only the bulk-flux field is corrected
    assert np.allclose(_interior(host.net["S"]), S - num / den, rtol=0, atol=1e-19)     # (long-double quotient vs double quotient)


def _gloo_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pyoracle.set_threads(1)
    Nx, Ny = 48, 20
    full = cj.LatitudeLongitudeGrid((Nx, Ny, 2), latitude=(-60.0, 60.0), halo=(4, 4, 2))
    grid = full.slab(rank, world)
    host = cj.SurfaceFluxData.synthetic(grid, ring=1)
    cfg = cj.default_config(grid.Nx, grid.Ny, 2, 64)
    cfg.grid.ring = 1
    cfg.grid.periodic_x = 0
    inp, out = host.update_bundles()
    pyoracle.update_state(cfg, inp, out, QUERY_TIME)
    keep = host.net["S"].data.copy()
    sums = pyoracle.normalize_salinity_flux(cfg, host.salinity_normalization())      # local sums (and a local subtraction …)
    host.net["S"].data[...] = keep                                                   # … undone: subtract the GLOBAL mean instead
    mean = slabs.combine_partial_sums(sums, dist, world)
    pyoracle.normalize_salinity_flux(cfg, host.salinity_normalization(), mean=mean)
    res = slabs.gather_interior(host.net["S"], dist, world)
    if rank == 0:
        q.put((res, mean))
    dist.barrier()
    dist.destroy_process_group()


def test_two_slabs_subtract_the_global_mean_gloo():
    import torch.multiprocessing as tmp
    ctx = tmp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, 29533, q)) for r in range(2)]
    for p in procs:
        p.start()
    res, mean = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    grid, host, cfg = make_case(48, 20, 2, 64, halo=(4, 4, 2))
    oracle_update(host, cfg)
    num, den = pyoracle.normalize_salinity_flux(cfg, host.salinity_normalization())
    assert abs(mean - num / den) <= 1e-15 * abs(num / den) + 1e-25
    assert np.max(np.abs(res - _interior(host.net["S"]))) <= 1e-15 * np.max(np.abs(res)) + 1e-25


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("bits", [64, 32])
@pytest.mark.parametrize("land_fraction,with_additional", [(0.0, False), (0.3, True)])
def test_cuda_normalization_matches_oracle_and_is_reproducible(bits, land_fraction, with_additional):
    import torch
    grid, host, cfg = make_case(192, 88, 3, bits, land_fraction=land_fraction)
    oracle_update(host, cfg)
    dev = host.to("cuda:0")
    dev.net["S"].data.copy_(torch.from_numpy(host.net["S"].data))
    add_h = add_d = None
    if with_additional:
        add_h = host.net["S"].clone()
        add_h.data[...] = np.random.default_rng(5).normal(0, 1e-7, add_h.data.shape).astype(np_dtype(bits))
        add_d = add_h.to("cuda:0")
    eng = cj.Engine(cfg)
    start = dev.net["S"].data.clone()
    eng.normalize_salinity_flux(dev.salinity_normalization(add_d))
    torch.cuda.synchronize()
    first = dev.net["S"].numpy().copy()
    dev.net["S"].data.copy_(start)
    eng.normalize_salinity_flux(dev.salinity_normalization(add_d))
    torch.cuda.synchronize()
    assert np.array_equal(first, dev.net["S"].numpy())          # fixed-order reduction: bit-reproducible
    num, den = pyoracle.normalize_salinity_flux(cfg, host.salinity_normalization(add_h))
    ref = host.net["S"].data
    scale = np.max(np.abs(ref))
    assert np.max(np.abs(first.astype(np.float64) - ref)) <= RTOL[bits] * scale
    # split form: the sums land in a device buffer, the subtraction uses them
    dev.net["S"].data.copy_(start)
    sums = torch.zeros(2, dtype=torch.float64, device="cuda:0")
    norm = dev.salinity_normalization(add_d)
    eng.salinity_flux_sums(norm, sums.data_ptr())
    eng.subtract_mean_flux(norm, sums.data_ptr())
    torch.cuda.synchronize()
    assert np.array_equal(first, dev.net["S"].numpy())
    s = sums.cpu().numpy()
    assert abs(s[1] - den) <= 1e-13 * den and abs(s[0] - num) <= 1e-12 * scale * den
    eng.close()


@pytest.mark.gpu
def test_full_size_normalization_leaves_a_zero_budget_and_model_callable_works():
    import torch
    from tests.test_full_size import _case
    grid, host, dev, cfg = _case("twelfth", 64, "default")
    eng = cj.Engine(cfg)
    inp, out = dev.update_bundles()
    eng.update_state(inp, out, QUERY_TIME)
    before = dev.net["S"].data.clone()
    eng.normalize_salinity_flux(dev.salinity_normalization())
    torch.cuda.synchronize()
    S = dev.net["S"].data[0, 7:-7, 7:-7].double()
    Az = dev.area.data[0, 7:-7, :].double()
    budget = float((S * Az).sum() / Az.sum() / grid.Nx)
    assert abs(budget) <= 1e-15 * float(before.abs().max())     # the area-weighted budget integrates to zero
    d = (before - dev.net["S"].data)
    assert float(d.max() - d.min()) <= 4 * np.spacing(float(before.abs().max()))   # one and the same mean removed from the whole parent (halos too)
    eng.close()

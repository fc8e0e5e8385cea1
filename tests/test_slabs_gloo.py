"""Longitude-slab decomposition on 2 ranks (gloo, CPU): the slab-wise flux solve — zero-message ring
mode and seam-exchange mode — reproduces the single-domain solve bit for bit.  The arithmetic here is
the CPU oracle (the checker); what is under test is the host-side slab logic (global-index synthetic
data, descriptors, seam exchange, gather) that bench.py --gpus N and the NCCL path share."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, mode, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import climaocean.jl_b200 as cj
    from climaocean.jl_b200 import slabs
    from oracle import pyoracle
    from tests.common import QUERY_TIME
    pyoracle.set_threads(1)
    Nx, Ny, Nz = 48, 20, 3
    full = cj.LatitudeLongitudeGrid((Nx, Ny, Nz), latitude=(-60.0, 60.0), halo=(4, 4, 2))
    grid = full.slab(rank, world)
    ring = 1 if mode == "ring" else 0
    host = cj.SurfaceFluxData.synthetic(grid, ring=ring)
    cfg = cj.default_config(grid.Nx, grid.Ny, Nz, 64)
    cfg.grid.ring = ring
    cfg.grid.periodic_x = 0
    if mode == "ring":
        inp, out = host.update_bundles()
        pyoracle.update_state(cfg, inp, out, QUERY_TIME)
        seam_bytes = 0
    else:
        pyoracle.interpolate_atmosphere(cfg, host.atmos_series(), QUERY_TIME, host.exchange_state())
        pyoracle.atmosphere_ocean_fluxes(cfg, host.exchange_state(), host.ocean_surface(), host.interface_fluxes("ao"))
        seam_bytes = slabs.exchange_seam(host, dist, rank, world)
        pyoracle.assemble_net_ocean_fluxes(cfg, host.exchange_state(), host.ocean_surface(), host.interface_fluxes("ao"),
                                           None, None, host.net_ocean_fluxes())
    res = {n: slabs.gather_interior(host.net[n], dist, world) for n in ("u", "v", "T", "S")}
    res["Qv"] = slabs.gather_interior(host.ao["latent_heat"], dist, world)
    if rank == 0:
        q.put((res, seam_bytes))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["ring", "seam"])
def test_two_slabs_reproduce_the_single_domain_solve(mode):
    sys.path.insert(0, ROOT)
    import climaocean.jl_b200 as cj
    from oracle import pyoracle
    from tests.common import QUERY_TIME
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, mode, q)) for r in range(2)]
    for p in procs:
        p.start()
    res, seam_bytes = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    # single-domain reference
    full = cj.LatitudeLongitudeGrid((48, 20, 3), latitude=(-60.0, 60.0), halo=(4, 4, 2))
    host = cj.SurfaceFluxData.synthetic(full, ring=1)
    cfg = cj.default_config(48, 20, 3, 64)
    inp, out = host.update_bundles()
    pyoracle.update_state(cfg, inp, out, QUERY_TIME)
    ref = host.outputs()
    for n, key in (("u", "net.u"), ("v", "net.v"), ("T", "net.T"), ("S", "net.S"), ("Qv", "ao.latent_heat")):
        assert res[n].shape == ref[key].shape
        a, b = res[n], ref[key]
        if mode == "seam" and n == "v":
            # with ring = 0 the halo row j = -1 of ρτy is not computed, so τy on the southern WALL face
            # (j = 0, a boundary face whose value the ocean never uses) differs from the ring-mode value
            a, b = a[1:], b[1:]
        assert np.array_equal(a, b), (mode, n)
    if mode == "seam":
        assert seam_bytes == (20 + 2 * 4) * 8          # one column of ρτx, Float64, halo rows included


@pytest.mark.parametrize("world", [2, 8])
def test_slab_coordinates_carry_the_bits_of_the_global_grid(world):
    """At 1/12° Δλ = 360/4320 is not representable: a slab that evaluated λ from its own west edge would get fractional
    source indices one ulp away from the one-process solve (seen as a checksum mismatch of slab 1 in bench.py --gpus 2).
    Slabs evaluate the global expression at the global column index instead."""
    import climaocean.jl_b200 as cj
    from climaocean.jl_b200.fields import fractional_indices
    full = cj.LatitudeLongitudeGrid((4320, 60, 1), latitude=(-75.0, 75.0), halo=(7, 7, 0))
    FI, FJ = fractional_indices(full, 640, 320)
    for rank in range(world):
        g = full.slab(rank, world)
        fi, fj = fractional_indices(g, 640, 320)
        nx = g.Nx
        assert np.array_equal(fi[0, :, 1:-1], FI[0, :, 1 + rank * nx:1 + (rank + 1) * nx])
        assert np.array_equal(fi[0, :, 0], FI[0, :, rank * nx]) and np.array_equal(fi[0, :, -1], FI[0, :, 1 + (rank + 1) * nx])
        assert np.array_equal(fj, FJ[:, :, :nx + 2])
        assert np.array_equal(g.horizontal_areas(), full.horizontal_areas())
        host = cj.SurfaceFluxData.synthetic(g, ring=1)
        ref = cj.SurfaceFluxData.synthetic(full, ring=1) if rank == 0 else ref
        for n in ("u", "v", "T", "S"):
            a, b = host.ocean[n].numpy(), ref.ocean[n].numpy()
            H = 7
            assert np.array_equal(a[:, :, H:-H], b[:, :, H + rank * nx:H + (rank + 1) * nx]), n


@pytest.mark.parametrize("world", [2, 3])
def test_slabs_with_land_series_and_sea_ice_match_the_single_domain(world):
    """Zero-message ring mode with the round-2 inputs: land freshwater series (their own source grid and fractional
    indices), sea ice (concentration, ice–ocean terms) and a land mask.  Every slab, solved on its own by the oracle from
    slab-local data, reproduces its columns of the single-domain solve bit for bit."""
    import climaocean.jl_b200 as cj
    from oracle import pyoracle
    from tests.common import QUERY_TIME
    pyoracle.set_threads(2)
    Nx, Ny, Nz = 48, 18, 3
    kw = dict(ring=1, with_land=True, with_ice=True, land_fraction=0.2)
    full = cj.LatitudeLongitudeGrid((Nx, Ny, Nz), latitude=(-70.0, 70.0), halo=(4, 4, 2))
    ref_host = cj.SurfaceFluxData.synthetic(full, **kw)
    cfg = cj.default_config(Nx, Ny, Nz, 64)
    inp, out = ref_host.update_bundles(True)
    pyoracle.sea_ice_ocean_fluxes(cfg, ref_host.ocean_columns(), ref_host.sea_ice_state(), 600.0, ref_host.ice_ocean_fluxes())
    pyoracle.update_state(cfg, inp, out, QUERY_TIME)
    ref = ref_host.outputs()
    nx = Nx // world
    for rank in range(world):
        g = full.slab(rank, world)
        host = cj.SurfaceFluxData.synthetic(g, **kw)
        c = cj.default_config(g.Nx, g.Ny, Nz, 64)
        c.grid.ring = 1
        c.grid.periodic_x = 0
        i2, o2 = host.update_bundles(True)
        pyoracle.sea_ice_ocean_fluxes(c, host.ocean_columns(), host.sea_ice_state(), 600.0, host.ice_ocean_fluxes())
        pyoracle.update_state(c, i2, o2, QUERY_TIME)
        got = host.outputs()
        for k in ("exchange.Mp", "net.T", "net.S", "net.u", "net.v", "ao.latent_heat", "io.frazil_heat"):
            assert np.array_equal(got[k], ref[k][:, rank * nx:(rank + 1) * nx]), (world, rank, k)

"""Longitude slabs on ≥2 GPUs (NCCL for the plumbing): zero-message ring mode and the NVLink seam-push
mode reproduce the single-GPU result bit for bit.  Skipped on a 1-GPU box (run with gpurun --gpus 2)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _worker(rank, world, port, mode, bits, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    import climaocean.jl_b200 as cj
    from climaocean.jl_b200 import slabs
    from tests.common import QUERY_TIME, np_dtype
    Nx, Ny, Nz = 256, 96, 4
    full = cj.LatitudeLongitudeGrid((Nx, Ny, Nz), latitude=(-60.0, 60.0), halo=(4, 4, 2), dtype=np_dtype(bits))
    grid = full.slab(rank, world)
    ring = 1 if mode == "ring" else 0
    host = cj.SurfaceFluxData.synthetic(grid, ring=ring)
    dev = host.to(f"cuda:{rank}")
    cfg = cj.default_config(grid.Nx, grid.Ny, Nz, bits)
    cfg.device = rank
    cfg.grid.ring = ring
    cfg.grid.periodic_x = 0
    eng = cj.Engine(cfg)
    if mode == "seam":
        slabs.attach_seam(eng, dist, rank, world)
    inp, out = dev.update_bundles()
    for step in range(5):                         # several steps: exercises the double-buffered seam + acks
        eng.update_state(inp, out, QUERY_TIME + 600.0 * step)
    torch.cuda.synchronize()
    res = {n: slabs.gather_interior(dev.net[n], dist, world) for n in ("u", "v", "T", "S")}
    res["Qv"] = slabs.gather_interior(dev.ao["latent_heat"], dist, world)
    if rank == 0:
        q.put(res)
    dist.barrier()
    if mode == "seam":
        eng.seam_detach()
    eng.close()
    dist.destroy_process_group()


@pytest.mark.parametrize("bits", [64, 32])
@pytest.mark.parametrize("mode", ["ring", "seam"])
def test_slabs_match_single_gpu(mode, bits):
    import torch
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    import climaocean.jl_b200 as cj
    from tests.common import QUERY_TIME, np_dtype
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, mode, bits, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    full = cj.LatitudeLongitudeGrid((256, 96, 4), latitude=(-60.0, 60.0), halo=(4, 4, 2), dtype=np_dtype(bits))
    host = cj.SurfaceFluxData.synthetic(full, ring=1)
    dev = host.to("cuda:0")
    cfg = cj.default_config(256, 96, 4, bits)
    eng = cj.Engine(cfg)
    inp, out = dev.update_bundles()
    eng.update_state(inp, out, QUERY_TIME + 600.0 * 4)
    torch.cuda.synchronize()
    ref = dev.outputs()
    for n, key in (("u", "net.u"), ("v", "net.v"), ("T", "net.T"), ("S", "net.S"), ("Qv", "ao.latent_heat")):
        a, b = res[n], ref[key]
        if mode == "seam" and n == "v":
            a, b = a[1:], b[1:]            # τy on the southern wall face is not defined without the halo ring
        assert np.array_equal(a, b), (mode, n)


def test_single_context_can_attach_to_itself():
    """world = 1: the periodic single slab in seam mode equals ring mode on all interior faces."""
    import torch
    import climaocean.jl_b200 as cj
    from climaocean.jl_b200 import slabs
    from tests.common import QUERY_TIME
    grid = cj.LatitudeLongitudeGrid((96, 40, 3), latitude=(-60.0, 60.0), halo=(4, 4, 2))
    outs = {}
    for ring in (1, 0):
        host = cj.SurfaceFluxData.synthetic(grid, ring=ring)
        dev = host.to("cuda:0")
        cfg = cj.default_config(96, 40, 3, 64)
        cfg.grid.ring = ring
        cfg.grid.periodic_x = 0
        eng = cj.Engine(cfg)
        if ring == 0:
            slabs.attach_seam(eng, None, 0, 1)
        inp, out = dev.update_bundles()
        for _ in range(4):
            eng.update_state(inp, out, QUERY_TIME)
        torch.cuda.synchronize()
        outs[ring] = dev.outputs()
        eng.close()
    for k in ("net.u", "net.T", "net.S", "ao.x_momentum"):
        assert np.array_equal(outs[0][k], outs[1][k]), k
    assert np.array_equal(outs[0]["net.v"][1:], outs[1]["net.v"][1:])


def _norm_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    import climaocean.jl_b200 as cj
    from climaocean.jl_b200 import slabs
    from tests.common import QUERY_TIME
    full = cj.LatitudeLongitudeGrid((256, 96, 4), latitude=(-60.0, 60.0), halo=(4, 4, 2))
    grid = full.slab(rank, world)
    host = cj.SurfaceFluxData.synthetic(grid, ring=1)
    dev = host.to(f"cuda:{rank}")
    cfg = cj.default_config(grid.Nx, grid.Ny, 4, 64)
    cfg.device = rank
    cfg.grid.ring = 1
    cfg.grid.periodic_x = 0
    eng = cj.Engine(cfg)
    inp, out = dev.update_bundles()
    eng.update_state(inp, out, QUERY_TIME)
    sums = slabs.normalize_salinity_flux(eng, dev.salinity_normalization(), dist, world)   # NCCL all-reduce of 2 doubles
    torch.cuda.synchronize()
    res = slabs.gather_interior(dev.net["S"], dist, world)
    if rank == 0:
        q.put((res, sums.cpu().numpy()))
    dist.barrier()
    eng.close()
    dist.destroy_process_group()


def test_salinity_normalization_across_slabs_nccl():
    """NormalizeSalinity (omip_simulation.jl:187-220) on ≥ 2 slabs: partial sums per GPU, NCCL all-reduce, the same global
    mean subtracted everywhere — equal to the single-GPU result up to the summation order of the slabs."""
    import torch
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    import climaocean.jl_b200 as cj
    from tests.common import QUERY_TIME
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_norm_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res, sums = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    full = cj.LatitudeLongitudeGrid((256, 96, 4), latitude=(-60.0, 60.0), halo=(4, 4, 2))
    host = cj.SurfaceFluxData.synthetic(full, ring=1)
    dev = host.to("cuda:0")
    cfg = cj.default_config(256, 96, 4, 64)
    eng = cj.Engine(cfg)
    inp, out = dev.update_bundles()
    eng.update_state(inp, out, QUERY_TIME)
    eng.normalize_salinity_flux(dev.salinity_normalization())
    torch.cuda.synchronize()
    ref = dev.outputs()["net.S"]
    assert np.max(np.abs(res - ref)) <= 1e-15 * np.max(np.abs(ref))
    eng.close()

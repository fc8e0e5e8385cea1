"""Longitude-slab decomposition helpers: one process per GPU, `Nx/P` columns × full Ny.

Mode A — zero message (`ring = 1`): every slab computes its fluxes into one halo ring from ocean halos
the host model already maintains; nothing to exchange.
Mode B — seam exchange (`ring = 0`): after the interface-flux kernel each rank sends its last interior
column of ρτx to its east neighbour (periodic ring), which needs it for the centre→face stress average
at its first face.  `exchange_seam` does that with torch.distributed point-to-point ops — NCCL over
NVLink for device tensors, gloo for the CPU tests.  Only plumbing lives here.
"""
import numpy as np


def _as_tensor(field):
    import torch
    d = field.data
    return d if type(d).__module__.startswith("torch") else torch.from_numpy(d)


def exchange_seam(data, dist, rank, world):
    """Fill column i = -1 of ao.x_momentum with the west neighbour's last interior column."""
    import torch
    f = data.ao["x_momentum"]
    t = _as_tensor(f)                       # (1, nj, ni)
    Hx = f.halo[0]
    Nx = data.grid.Nx
    send = t[0, :, Hx + Nx - 1].contiguous()
    recv = torch.empty_like(send)
    east, west = (rank + 1) % world, (rank - 1) % world
    if world == 1:
        recv.copy_(send)
    else:
        ops = [dist.P2POp(dist.isend, send, east), dist.P2POp(dist.irecv, recv, west)]
        for r in dist.batch_isend_irecv(ops):
            r.wait()
    t[0, :, Hx - 1] = recv
    return int(send.numel() * send.element_size())


def gather_interior(field, dist, world):
    """All-gather the interior (Ny, Nx_local) of a 2-D field along longitude → (Ny, Nx_global) numpy."""
    import torch
    t = _as_tensor(field)
    Hx, Hy, _ = field.halo
    loc = t[0, Hy:t.shape[1] - Hy, Hx:t.shape[2] - Hx].contiguous()
    if world == 1:
        return loc.cpu().numpy()
    parts = [torch.empty_like(loc) for _ in range(world)]
    dist.all_gather(parts, loc)
    return torch.cat(parts, dim=1).cpu().numpy()


def attach_seam(engine, dist, rank, world):
    """Mode B set-up: all-gather every rank's seam handle and attach the west / east neighbours
    (periodic ring).  After this, coflux_update_state pushes the seam column over NVLink itself."""
    import torch
    h = engine.seam_export()
    if world == 1:
        engine.seam_attach(h, h, 0, 1)
        return
    mine = torch.tensor(list(h), dtype=torch.uint8, device=f"cuda:{torch.cuda.current_device()}")
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    handles = [bytes(p.cpu().tolist()) for p in parts]
    engine.seam_attach(handles[(rank - 1) % world], handles[(rank + 1) % world], rank, world)
    dist.barrier()


def normalize_salinity_flux(engine, norm, dist=None, world=1, stream=None):
    """NormalizeSalinity across longitude slabs (/root/reference/src/OMIPConfigurations/omip_simulation.jl:187-220):
    every rank reduces its slab to (Σ f·Az, Σ Az) on the device, the two doubles are all-reduced (the one real collective
    of this path: NCCL over NVLink for device tensors, gloo in the CPU tests), every rank subtracts the same mean.
    Returns the device tensor holding the global sums."""
    import torch
    sums = torch.zeros(2, dtype=torch.float64, device=f"cuda:{torch.cuda.current_device()}")
    engine.salinity_flux_sums(norm, sums.data_ptr(), stream)
    if world > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    engine.subtract_mean_flux(norm, sums.data_ptr(), stream)
    return sums


def combine_partial_sums(local_sums, dist=None, world=1):
    """Host-side half of the above for backends without device tensors (gloo tests, the CPU oracle): all-reduce the
    (Σ f·Az, Σ Az) pair and return the global mean."""
    import torch
    t = torch.tensor([float(local_sums[0]), float(local_sums[1])], dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t[0] / t[1]) if float(t[1]) != 0.0 else 0.0

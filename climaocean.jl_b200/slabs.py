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

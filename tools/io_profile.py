"""One launch of the sea-ice–ocean kernel at 1/12°, Nz = 75 for ncu (GPU box)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import climaocean.jl_b200 as cj
from bench import make_cfg, NX, NY, NZ
gi = cj.LatitudeLongitudeGrid((NX, NY, 1), latitude=(-75.0, 75.0), halo=(7, 7, 0))
hi = cj.SurfaceFluxData.synthetic(gi, with_ice=True)
di = hi.to_device_columns("cuda:0", NZ, fill_columns=True)
ei = cj.Engine(make_cfg(di.grid, NZ, 64, 0))
ei.compute_sea_ice_ocean_fluxes(di.ocean_columns(), di.sea_ice_state(), 600.0, di.ice_ocean_fluxes())
torch.cuda.synchronize()
print("done")

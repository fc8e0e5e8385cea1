"""Run the atmosphere–sea-ice solve (row a7) on a synthetic grid and dump its outputs (GPU box): used to compare kernel
variants bit for bit.  python tools/ice_dump.py out.npz [bits] [cfg]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import climaocean.jl_b200 as cj
QUERY_TIME = 1.37 * 10800.0          # tests/common.py
out = sys.argv[1]
bits = int(sys.argv[2]) if len(sys.argv) > 2 else 64
name = sys.argv[3] if len(sys.argv) > 3 else "default"
dtype = np.float64 if bits == 64 else np.float32
grid = cj.LatitudeLongitudeGrid((700, 333, 3), latitude=(-60.0, 60.0), halo=(7, 7, 7), dtype=dtype)
host = cj.SurfaceFluxData.synthetic(grid, ring=1, with_ice=True, land_fraction=0.2)
cfg = cj.default_config(700, 333, 3, bits, name)
cfg.grid.ring = 1
dev = host.to("cuda:0")
eng = cj.Engine(cfg)
eng.interpolate_atmosphere_state(dev.atmos_series(), QUERY_TIME, dev.exchange_state())
eng.compute_atmosphere_sea_ice_fluxes(dev.exchange_state(), dev.ocean_surface(), dev.sea_ice_state(), dev.interface_fluxes("ai"))
torch.cuda.synchronize()
res = {k: v for k, v in dev.outputs().items() if k.startswith("ai.")}
res["its"] = dev.iterations_ai.numpy()
res["Ttop"] = dev.ice["top_temperature"].numpy()
np.savez(out, **res)
print(out, {k: float(np.abs(v).sum()) for k, v in list(res.items())[:4]})

"""Run one ocean-only update_state! on a synthetic grid and dump its outputs (GPU box): used to compare library builds bit
for bit.  python tools/ocean_dump.py out.npz [bits] [cfg]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import climaocean.jl_b200 as cj
QUERY_TIME = 1.37 * 10800.0
out = sys.argv[1]
bits = int(sys.argv[2]) if len(sys.argv) > 2 else 64
name = sys.argv[3] if len(sys.argv) > 3 else "default"
dtype = np.float64 if bits == 64 else np.float32
grid = cj.LatitudeLongitudeGrid((700, 333, 3), latitude=(-75.0, 75.0), halo=(7, 7, 7), dtype=dtype)
host = cj.SurfaceFluxData.synthetic(grid, ring=1, land_fraction=0.1)
cfg = cj.default_config(700, 333, 3, bits, name)
cfg.grid.ring = 1
dev = host.to("cuda:0")
eng = cj.Engine(cfg)
inp, o = dev.update_bundles()
eng.update_state(inp, o, QUERY_TIME)
torch.cuda.synchronize()
res = dev.outputs()
res["its"] = dev.iterations.numpy()
np.savez(out, **res)
print(out, bits, name, float(np.abs(res["net.T"]).sum()))

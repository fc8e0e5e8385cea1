"""Time the atmosphere–sea-ice solve (row a7) at 1/12° (GPU box): python tools/ai_bench.py [bits] [cfg]"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import climaocean.jl_b200 as cj
from bench import make_cfg, NX, NY, QUERY_TIME
bits = int(sys.argv[1]) if len(sys.argv) > 1 else 64
name = sys.argv[2] if len(sys.argv) > 2 else "default"
gi = cj.LatitudeLongitudeGrid((NX, NY, 1), latitude=(-75.0, 75.0), halo=(7, 7, 0), dtype=np.float64 if bits == 64 else np.float32)
hi = cj.SurfaceFluxData.synthetic(gi, with_ice=True)
di = hi.to_device_columns("cuda:0", 2)
ei = cj.Engine(make_cfg(di.grid, 2, bits, 0, name))
st = torch.cuda.current_stream()
xch, oc, ice, ai = di.exchange_state(), di.ocean_surface(), di.sea_ice_state(), di.interface_fluxes("ai")
ei.interpolate_atmosphere_state(di.atmos_series(), QUERY_TIME, xch, st)
T0 = di.ice["top_temperature"].data.clone()
ts = []
for k in range(6):
    di.ice["top_temperature"].data.copy_(T0)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st); ei.compute_atmosphere_sea_ice_fluxes(xch, oc, ice, ai, st); b.record(st)
    torch.cuda.synchronize()
    if k >= 2: ts.append(a.elapsed_time(b))
t = float(np.mean(ts))
print(f"f{bits} {name}: atmosphere-sea-ice solve {t:.2f} ms  {NX*NY/t/1e3:.0f} Mcells/s  (ice-covered fraction {(di.ice['concentration'].data > 0).double().mean().item():.2f})")

"""Time the sea-ice–ocean kernel at 1/12°, Nz = 75 (GPU box)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import climaocean.jl_b200 as cj
from bench import make_cfg, NX, NY, NZ, load_peaks
bits = int(sys.argv[1]) if len(sys.argv) > 1 else 64
peak, _ = load_peaks()
gi = cj.LatitudeLongitudeGrid((NX, NY, 1), latitude=(-75.0, 75.0), halo=(7, 7, 0), dtype=np.float64 if bits == 64 else np.float32)
hi = cj.SurfaceFluxData.synthetic(gi, with_ice=True)
di = hi.to_device_columns("cuda:0", NZ, fill_columns=True)
ei = cj.Engine(make_cfg(di.grid, NZ, bits, 0))
cols, ice, io = di.ocean_columns(), di.sea_ice_state(), di.ice_ocean_fluxes()
T0 = di.ocean["T"].data.clone()
st = torch.cuda.current_stream()
tms = []
for k in range(8):
    di.ocean["T"].data.copy_(T0)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st); ei.compute_sea_ice_ocean_fluxes(cols, ice, 600.0, io, st); b.record(st)
    torch.cuda.synchronize()
    if k >= 3: tms.append(a.elapsed_time(b))
t = float(np.mean(tms)); words = 2 * NZ + 13; es = bits // 8
print(f"f{bits} ice_ocean {t:.3f} ms  {NX*NY*words*es/t/1e6:.0f} GB/s algorithmic  frac {NX*NY*words*es/t/1e6/peak:.3f}  frazil frac {(di.ocean['T'].data != T0).float().mean().item():.3f}")

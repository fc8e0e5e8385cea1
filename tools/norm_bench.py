"""Time NormalizeSalinity and the closure front end at 1/12° (GPU box)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import climaocean.jl_b200 as cj
from bench import make_host_case, make_cfg, NX, NY, NZ, QUERY_TIME, load_peaks
peak, _ = load_peaks()
grid, host = make_host_case(NX, NY, 64, 0, 1)
dev = host.to_device_columns("cuda:0", NZ)
eng = cj.Engine(make_cfg(dev.grid, NZ, 64, 0))
inp, out = dev.update_bundles()
eng.update_state(inp, out, QUERY_TIME)
st = torch.cuda.current_stream()
def timeit(fn, n=10):
    ts = []
    for k in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st); fn(); b.record(st); torch.cuda.synchronize()
        if k >= 3: ts.append(a.elapsed_time(b))
    return float(np.mean(ts))
norm = dev.salinity_normalization()
t = timeit(lambda: eng.normalize_salinity_flux(norm, st))
print(f"normalize_salinity: {t*1e3:.1f} us  {NX*NY*24/t/1e6:.0f} GB/s (3 words/cell)  frac {NX*NY*24/t/1e6/peak:.3f}")
f = dev.closure_forcing()
net = dev.net_ocean_fluxes()
t = timeit(lambda: eng.closure_surface_forcing(net, f, st))
print(f"closure_surface_forcing (stand-alone): {t*1e3:.1f} us  {NX*NY*80/t/1e6:.0f} GB/s (6 reads + 4 writes/cell)  frac {NX*NY*80/t/1e6/peak:.3f}")
eng.attach_closure_forcing(f)
eng.profile(True); eng.profile_read()
for _ in range(5): eng.update_state(inp, out, QUERY_TIME, st)
torch.cuda.synchronize()
fl, sm, n = eng.profile_read()
print(f"stress kernel with fused closure by-products: {sm/n*1e3:.1f} us (without: ~67 us)")

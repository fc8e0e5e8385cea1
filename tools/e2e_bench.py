"""Time the HOST-buffer entry (coflux_update_state_host) at 1/12° (GPU box): python tools/e2e_bench.py [bits]
COFLUX_HOST_CHUNKS=n overrides the number of row chunks of the pipeline."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import climaocean.jl_b200 as cj
from bench import make_host_case, make_cfg, time_e2e_steps, NX, NY
bits = int(sys.argv[1]) if len(sys.argv) > 1 else 64
grid, host = make_host_case(NX, NY, bits, 0, 1)
dev = host.to_device_columns("cuda:0", 1)
eng = cj.Engine(make_cfg(grid, 1, bits, 0))
ms, h2d, d2h, _ = time_e2e_steps(eng, dev, host, grid, bits, 12, 3, None, "cuda:0")
print(f"f{bits} chunks {os.environ.get('COFLUX_HOST_CHUNKS', 'auto')}: e2e {ms:.3f} ms/step  {NX*NY/ms/1e3:.0f} Mcells/s  h2d {h2d/1e6:.0f} MB d2h {d2h/1e6:.0f} MB")

"""Small driver for ncu (run on the GPU box): a few update_state! steps on the 1/4° grid."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import climaocean.jl_b200 as cj
from bench import make_host_case, make_cfg, QUERY_TIME
bits = int(sys.argv[1]) if len(sys.argv) > 1 else 64
name = sys.argv[2] if len(sys.argv) > 2 else "default"
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
res = sys.argv[4] if len(sys.argv) > 4 else "quarter"
NXY = {"quarter": (1440, 600, 10), "twelfth": (4320, 1800, 75)}[res]
grid, host = make_host_case(NXY[0], NXY[1], bits, 0, 1)
dev = host.to_device_columns("cuda:0", NXY[2])
eng = cj.Engine(make_cfg(dev.grid, NXY[2], bits, 0, name))
inp, out = dev.update_bundles()
for _ in range(steps):
    eng.update_state(inp, out, QUERY_TIME)
torch.cuda.synchronize()
print("done", eng.launches)

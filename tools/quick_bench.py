"""Quick device-resident timing of update_state on the GPU box: python tools/quick_bench.py [bits] [cfg...]"""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import climaocean.jl_b200 as cj
from bench import make_host_case, make_cfg, time_device_steps, NX, NY, NZ
bits = int(sys.argv[1]) if len(sys.argv) > 1 else 64
names = sys.argv[2:] or ["default", "corrected"]
NX = int(os.environ.get('QB_NX', NX))      # e.g. QB_NX=540: one of 8 longitude slabs of the 1/12 degree grid
NY = int(os.environ.get('QB_NY', NY))      # e.g. QB_NX=1440 QB_NY=600: the 1/4 degree grid
grid, host = make_host_case(NX, NY, bits, 0, 1)
dev = host.to_device_columns("cuda:0", NZ)
for name in names:
    eng = cj.Engine(make_cfg(dev.grid, NZ, bits, 0, name))
    ms, launches, fms, sms, _ = time_device_steps(eng, dev, 10, 3, None, "cuda:0")
    its = dev.iterations.numpy()[0, 7:-7, 7:-7]
    print(f"f{bits} {name:10s} step {ms:8.3f} ms  flux {fms:8.3f} ms  stress {sms:6.3f} ms  {NX*NY/ms/1e3:9.1f} Mcells/s  its mean {its.mean():.2f} max {its.max()}", flush=True)
    eng.close()

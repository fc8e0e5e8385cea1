"""Diagnostic (GPU box): CUDA vs oracle on the first rows of the 1/12° grid — worst cells per field."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
import climaocean.jl_b200 as cj
from tests.test_full_size import _case, _oracle_rows
from tests.common import QUERY_TIME
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 48
grid, host, dev, cfg = _case("twelfth", 64, sys.argv[2] if len(sys.argv) > 2 else "default")
eng = cj.Engine(cfg)
inp, out = dev.update_bundles()
eng.update_state(inp, out, QUERY_TIME); torch.cuda.synchronize()
gpu = dev.outputs(); its_g = dev.iterations.numpy()[0, 7:-7, 7:-7][:rows].copy()
ref = _oracle_rows(host, cfg, 0, rows); its_r = host.iterations.numpy()[0, 7:-7, 7:-7][:rows].copy()
print("iteration mismatches", int((its_g != its_r).sum()), "of", its_r.size)
for k in sorted(ref):
    if k not in gpu or not (k.startswith("ao.") or k.startswith("net.") or k.startswith("exchange.")): continue
    a = gpu[k][:rows].astype(np.float64); b = ref[k].astype(np.float64); sc = np.abs(b).max()
    if sc == 0: continue
    e = np.abs(a - b) / np.maximum(np.abs(b), 1e-3 * sc)
    j, i = np.unravel_index(np.argmax(e), e.shape)
    print(f"{k:32s} max {e.max():.2e} p99.99 {np.quantile(e, 0.9999):.2e} at ({i},{j}) ref {b[j,i]:.6e} its {its_r[j,i]}/{its_g[j,i]}  n>1e-12: {(e>1e-12).sum()}")

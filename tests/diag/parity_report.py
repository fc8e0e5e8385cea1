"""Diagnostic (run on the GPU box): per-field CUDA-vs-oracle error at several denominator floors,
iteration-count agreement, and the cells carrying the largest errors."""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
from tests.common import make_case, oracle_update, gpu_update

def report(Nx, Ny, bits, cfgname):
    grid, host, cfg = make_case(Nx, Ny, 4, bits, flux_configuration=cfgname)
    ref = oracle_update(host, cfg)
    gpu, dev = gpu_update(host, cfg)
    its_ref = host.iterations.numpy()[0, 7:-7, 7:-7]
    its_gpu = gpu["_iterations"][0, 7:-7, 7:-7]
    print(f"== {Nx}x{Ny} f{bits} {cfgname}: iteration mismatches {(its_ref != its_gpu).sum()} of {its_ref.size}; mean its {its_ref.mean():.2f}")
    for k, b in ref.items():
        if k not in gpu or k.startswith("ai.") or k.startswith("io."):
            continue
        a = gpu[k].astype(np.float64); b = b.astype(np.float64)
        scale = np.abs(b).max()
        if scale == 0:
            continue
        d = np.abs(a - b)
        row = [f"{k:30s} max|d|/max|b| {d.max()/scale:.2e}"]
        for fl in (1e-6, 1e-3, 1e-2, 1e-1):
            row.append(f"floor{fl:g}: {np.max(d/np.maximum(np.abs(b), fl*scale)):.2e}")
        # un-floored relative error of the cells that carry signal (|ref| > 1e-6·max): percentiles, and the share of cells over the
        # north_star tolerance at the 1e-3 floor — the evidence behind tests/common.py::FLOOR
        sig = np.abs(b) > 1e-6 * scale
        r = (d[sig] / np.abs(b[sig])) if sig.any() else np.zeros(1)
        tol = 1e-12 if bits == 64 else 1e-5
        over = np.mean(d / np.maximum(np.abs(b), 1e-3 * scale) > tol)
        row.append("raw rel p50/p99/p99.9/max " + "/".join(f"{np.percentile(r, q):.1e}" for q in (50, 99, 99.9, 100)) + f"  cells>tol@1e-3: {over:.2e}")
        j, i = np.unravel_index(np.argmax(d / np.maximum(np.abs(b), 1e-3 * scale)), d.shape)
        row.append(f"worst@({i},{j}) ref {b[j,i]:.6e} its {its_ref[j,i]}/{its_gpu[j,i]}")
        print("  ".join(row))

if __name__ == "__main__":
    for bits in (64, 32):
        for name in ("default", "corrected", "ncar"):
            report(256, 128, bits, name)

"""Generate the committed golden fixtures.

    python tests/golden/make_golden.py

IMPORTANT: these vectors are ORACLE-generated (oracle/libcoflux_oracle.so, the CPU restatement), not
reference-generated: the Julia reference cannot run in this environment and ships no golden vector of
its own.  They pin the oracle against accidental change (CPU test) and let the
GPU test compare the CUDA path without executing the oracle.  Inputs are not stored: they are
regenerated bit-identically from climaocean.jl_b200/synth.py (integer hashing + IEEE basic ops).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.common import QUERY_TIME, make_case, oracle_update  # noqa: E402

KEYS = ("exchange.T", "exchange.q", "exchange.Mp", "ao.latent_heat", "ao.sensible_heat", "ao.water_vapor", "ao.x_momentum",
        "ao.y_momentum", "ao.friction_velocity", "net.u", "net.v", "net.T", "net.S", "net.upwelling_longwave")
CASES = {"c1_default_f64": dict(Nx=64, Ny=32, Nz=8, bits=64, flux_configuration="default"),
         "c1_corrected_f64": dict(Nx=64, Ny=32, Nz=8, bits=64, flux_configuration="corrected"),
         "c1_ncar_f64": dict(Nx=64, Ny=32, Nz=8, bits=64, flux_configuration="ncar"),
         "c1_default_f32": dict(Nx=64, Ny=32, Nz=8, bits=32, flux_configuration="default")}


def run_case(spec):
    grid, host, cfg = make_case(spec["Nx"], spec["Ny"], spec["Nz"], spec["bits"], flux_configuration=spec["flux_configuration"])
    out = oracle_update(host, cfg, QUERY_TIME)
    res = {k: out[k] for k in KEYS}
    res["iterations"] = host.iterations.numpy()[0, 7:-7, 7:-7].astype(np.int16)
    return res


if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    for name, spec in CASES.items():
        res = run_case(spec)
        np.savez_compressed(os.path.join(here, name + ".npz"), **res)
        print(name, {k: float(np.abs(v).max()) for k, v in list(res.items())[:3]})

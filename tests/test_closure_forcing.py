"""Surface-forcing front ends of the ocean mixing closures that consume the net fluxes:
    KPP       /root/reference/src/OMIPConfigurations/KPP/kpp_surface_forcing.jl:18-29
    NEMO-TKE  /root/reference/src/OMIPConfigurations/NEMOTKE/nemo_tke_surface_forcing.jl:14-22
CPU: the oracle restatement on hand-computed values.  GPU: the stand-alone kernel and the form fused into the
centre→face stress kernel, bit for bit against the oracle (only IEEE +, ×, √ and max are involved)."""
import numpy as np
import pytest

import climaocean.jl_b200 as cj
from oracle import pyoracle
from tests.common import QUERY_TIME, make_case, np_dtype, oracle_update


def _interior(f):
    a = f.numpy()
    Hx, Hy, _ = f.halo
    return a[0, Hy:a.shape[1] - Hy, Hx:a.shape[2] - Hx]


def test_oracle_known_answers():
    grid, host, cfg = make_case(16, 8, 2, 64)
    host.net["u"].data[...] = 3e-4
    host.net["v"].data[...] = -4e-4
    host.net["T"].data[...] = 2e-5
    host.net["S"].data[...] = -1e-6
    f = host.closure_forcing(gravitational_acceleration=9.8)
    host.eos["alpha"].data[...] = 2e-4
    host.eos["beta"].data[...] = 8e-4
    pyoracle.closure_surface_forcing(cfg, host.net_ocean_fluxes(), f)
    c = host.closure
    assert np.all(_interior(c["friction_velocity_squared"]) == np.sqrt(3e-4 ** 2 + 4e-4 ** 2))          # |τ| = 5e-4
    assert np.allclose(_interior(c["friction_velocity"]), np.sqrt(5e-4), rtol=1e-15)
    assert np.allclose(_interior(c["surface_tke"]), 3.75 * 5e-4, rtol=1e-15)                            # > e_min0 = 1e-4
    assert np.allclose(_interior(c["buoyancy_flux"]), -9.8 * (2e-4 * 2e-5 - 8e-4 * -1e-6), rtol=1e-15)  # stabilising positive
    # floors: no stress → u★ = u★_min, e = e_min0
    host.net["u"].data[...] = 0.0
    host.net["v"].data[...] = 0.0
    pyoracle.closure_surface_forcing(cfg, host.net_ocean_fluxes(), f)
    assert np.all(_interior(c["friction_velocity"]) == 1e-6) and np.all(_interior(c["surface_tke"]) == 1e-4)
    assert np.all(_interior(c["friction_velocity_squared"]) == 0.0)


@pytest.mark.gpu
@pytest.mark.parametrize("bits", [64, 32])
def test_cuda_closure_forcing_standalone_and_fused_match_oracle_bitwise(bits):
    import torch
    grid, host, cfg = make_case(160, 72, 3, bits)
    fh = host.closure_forcing()
    oracle_update(host, cfg)
    pyoracle.closure_surface_forcing(cfg, host.net_ocean_fluxes(), fh)
    ref = {k: _interior(v).copy() for k, v in host.closure.items()}

    dev = host.to("cuda:0")
    fd = dev.closure_forcing()
    eng = cj.Engine(cfg)
    inp, out = dev.update_bundles()
    # fused: attached → update_state emits the by-products from the stress kernel (still 2 launches)
    eng.attach_closure_forcing(fd)
    l0 = eng.launches
    eng.update_state(inp, out, QUERY_TIME)
    torch.cuda.synchronize()
    assert eng.launches - l0 == 2
    net_gpu = {k: _interior(v).copy() for k, v in dev.net.items()}
    fused = {k: _interior(v).copy() for k, v in dev.closure.items()}
    # the by-products must be exactly the reference formulas applied to the GPU's own net fluxes …
    tx, ty = net_gpu["u"], net_gpu["v"]
    u2 = np.sqrt(tx * tx + ty * ty)
    assert np.array_equal(fused["friction_velocity_squared"], u2)
    assert np.array_equal(fused["friction_velocity"], np.maximum(np.sqrt(u2), np_dtype(bits)(1e-6)))
    assert np.array_equal(fused["surface_tke"], np.maximum(np_dtype(bits)(1e-4), np_dtype(bits)(3.75) * u2))
    # … and within the flux tolerance of the oracle's end-to-end values
    tol = 1e-12 if bits == 64 else 1e-5
    for k in ref:
        scale = np.max(np.abs(ref[k]))
        assert np.max(np.abs(fused[k].astype(np.float64) - ref[k])) <= tol * scale * 10, k
    # stand-alone kernel on the same net fluxes: bit-identical to the fused by-products
    eng.attach_closure_forcing(None)
    for v in dev.closure.values():
        v.data.zero_()
    eng.closure_surface_forcing(dev.net_ocean_fluxes(), fd)
    torch.cuda.synchronize()
    for k in fused:
        assert np.array_equal(_interior(dev.closure[k]), fused[k]), k
    eng.close()

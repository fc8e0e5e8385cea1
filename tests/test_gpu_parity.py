"""Parity of the CUDA path (through the C ABI) against the CPU oracle on identical seeded inputs.
Tolerances: BASELINE.json north_star — ≤1e-12 relative in Float64, ≤1e-5 in Float32 (denominator
floored at 0.1 % of the field's largest magnitude, tests/common.py).  Needs a B200."""
import ctypes as C

import numpy as np
import pytest

import climaocean.jl_b200 as cj
from climaocean.jl_b200 import _abi
from oracle import pyoracle
from tests.common import QUERY_TIME, RTOL, compare, compare_sea_ice, gpu_update, make_case, oracle_update, rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("bits", [64, 32])
@pytest.mark.parametrize("flux_configuration", ["default", "corrected", "ncar"])
def test_update_state_config1(bits, flux_configuration):
    """BASELINE config 1: 64×32×8 lat-lon grid, synthetic prescribed atmosphere, one update_state!."""
    grid, host, cfg = make_case(64, 32, 8, bits, flux_configuration=flux_configuration)
    ref = oracle_update(host, cfg)
    gpu, dev = gpu_update(host, cfg)
    worst = compare(gpu, ref, bits)
    assert gpu["_launches"] == 2
    if bits == 64:
        its_ref = host.iterations.numpy()
        assert np.array_equal(gpu["_iterations"], its_ref)       # same iteration path, cell by cell
    print(flux_configuration, bits, max(worst.values()))


@pytest.mark.parametrize("bits", [64, 32])
def test_update_state_wind_velocity_and_land_mask(bits):
    grid, host, cfg = make_case(96, 40, 5, bits, velocity="wind", land_fraction=0.3)
    ref = oracle_update(host, cfg)
    gpu, _ = gpu_update(host, cfg)
    compare(gpu, ref, bits)
    wet = host.mask.data[0, 7:-7, 7:-7] != 0
    assert np.all(gpu["ao.latent_heat"][~wet] == 0) and np.all(gpu["net.T"][~wet] == 0)


@pytest.mark.parametrize("bits", [64, 32])
def test_ragged_sizes_and_ring0_periodic(bits):
    """Sizes that are not multiples of the block size; ring = 0 uses the periodic wrap for τx."""
    for (Nx, Ny, ring) in ((1, 1, 1), (33, 7, 1), (130, 3, 0), (257, 65, 0)):
        grid, host, cfg = make_case(Nx, Ny, 3, bits, halo=(4, 4, 2), ring=ring)
        ref = oracle_update(host, cfg)
        gpu, _ = gpu_update(host, cfg)
        compare(gpu, ref, bits)


@pytest.mark.parametrize("bits", [64, 32])
def test_unfused_entry_points_match_oracle_and_fused_path(bits):
    grid, host, cfg = make_case(80, 36, 4, bits)
    ref = oracle_update(host, cfg)
    fused, _ = gpu_update(host, cfg)
    import torch
    dev = host.to("cuda:0")
    eng = cj.Engine(cfg)
    s, x, o = dev.atmos_series(), dev.exchange_state(), dev.ocean_surface()
    f, n = dev.interface_fluxes("ao"), dev.net_ocean_fluxes()
    eng.interpolate_atmosphere_state(s, QUERY_TIME, x)
    eng.compute_atmosphere_ocean_fluxes(x, o, f)
    eng.compute_net_ocean_fluxes(x, o, f, None, None, n)
    torch.cuda.synchronize()
    assert eng.launches == 3
    un = dev.outputs()
    compare(un, ref, bits)
    for k in ref:                                   # fused and un-fused are the same arithmetic: bit-identical
        if k in un and k in fused and not k.startswith("ai.") and not k.startswith("io."):
            assert np.array_equal(un[k], fused[k]), k


@pytest.mark.parametrize("bits", [64, 32])
@pytest.mark.parametrize("flux_configuration", ["default", "corrected"])
def test_sea_ice_ocean_fluxes_and_ice_aware_assembly(bits, flux_configuration):
    import torch
    grid, host, cfg = make_case(72, 30, 12, bits, with_ice=True, frazil=True, flux_configuration=flux_configuration)
    dev = host.to("cuda:0")
    dt = 900.0
    # oracle
    pyoracle.sea_ice_ocean_fluxes(cfg, host.ocean_columns(), host.sea_ice_state(), dt, host.ice_ocean_fluxes())
    ref = oracle_update(host, cfg, with_ice_terms=True)
    # CUDA
    eng = cj.Engine(cfg)
    eng.compute_sea_ice_ocean_fluxes(dev.ocean_columns(), dev.sea_ice_state(), dt, dev.ice_ocean_fluxes())
    inp, out = dev.update_bundles(with_ice_terms=True)
    eng.update_state(inp, out, QUERY_TIME)
    torch.cuda.synchronize()
    gpu = dev.outputs()
    compare(gpu, ref, bits)
    # the frazil sweep edits the ocean temperature in place: compare the whole column array
    assert rel_err(dev.ocean["T"].numpy(), host.ocean["T"].numpy(), bits) <= RTOL[bits]
    assert np.array_equal(dev.ice["previous_thickness"].numpy(), host.ice["previous_thickness"].numpy())
    assert np.any(gpu["io.frazil_heat"] != 0)


@pytest.mark.parametrize("bits", [64, 32])
@pytest.mark.parametrize("flux_configuration", ["default", "corrected", "ncar", "default+linearized_longwave"])
def test_atmosphere_sea_ice_fluxes(bits, flux_configuration):
    import torch
    grid, host, cfg = make_case(64, 28, 4, bits, with_ice=True, flux_configuration=flux_configuration.split("+")[0])
    if flux_configuration.endswith("+linearized_longwave"):          # include/coflux.h: coflux_skin_temperature_update
        cfg.atmosphere_sea_ice.skin_temperature_update = _abi.SKIN_LINEARIZED_LONGWAVE
    dev = host.to("cuda:0")
    pyoracle.interpolate_atmosphere(cfg, host.atmos_series(), QUERY_TIME, host.exchange_state())
    pyoracle.atmosphere_sea_ice_fluxes(cfg, host.exchange_state(), host.ocean_surface(), host.sea_ice_state(), host.interface_fluxes("ai"))
    eng = cj.Engine(cfg)
    eng.interpolate_atmosphere_state(dev.atmos_series(), QUERY_TIME, dev.exchange_state())
    eng.compute_atmosphere_sea_ice_fluxes(dev.exchange_state(), dev.ocean_surface(), dev.sea_ice_state(), dev.interface_fluxes("ai"))
    torch.cuda.synchronize()
    ref, gpu = host.outputs(), dev.outputs()
    its_ref = host.iterations_ai.numpy()[0, 7:-7, 7:-7]
    its_gpu = dev.iterations_ai.numpy()[0, 7:-7, 7:-7]
    gpu["ice.top_temperature"] = dev.ice["top_temperature"].numpy()[0, 7:-7, 7:-7]
    ref["ice.top_temperature"] = host.ice["top_temperature"].numpy()[0, 7:-7, 7:-7]
    # north_star tolerance (1e-12 / 1e-5) on every cell that converges; limit-cycle cells: tests/common.py::compare_sea_ice
    worst, frac = compare_sea_ice(gpu, ref, bits, its_gpu, its_ref, cfg.atmosphere_sea_ice.max_iterations,
                                  [k for k in ref if k.startswith("ai.")] + ["ice.top_temperature"])
    print(flux_configuration, bits, "worst", max(worst.values()), "limit-cycle cells", frac)
    assert np.any(gpu["ai.sensible_heat"] != 0)


def test_calm_cells_and_exact_zero_wind():
    """≥1 cell with Δu = Δv = 0 exactly: stress is exactly zero, nothing is NaN."""
    import torch
    grid, host, cfg = make_case(40, 20, 3, 64)
    for n in ("u", "v"):
        host.atmos[n].data[...] = 0.0
        host.ocean[n].data[...] = 0.0
    cfg.atmosphere_ocean.gustiness_parameter = 0.0
    ref = oracle_update(host, cfg)
    gpu, _ = gpu_update(host, cfg)
    for k, v in gpu.items():
        if not k.startswith("_"):
            assert not np.isnan(v).any(), k
    assert np.all(gpu["ao.x_momentum"] == 0) and np.all(gpu["net.u"] == 0) and np.all(gpu["ao.latent_heat"] == 0)
    compare(gpu, ref, 64)


def test_run_to_run_determinism_and_independence_from_stream():
    import torch
    grid, host, cfg = make_case(128, 48, 3, 64)
    a, _ = gpu_update(host, cfg)
    b, _ = gpu_update(host, cfg)
    for k in a:
        if not k.startswith("_"):
            assert np.array_equal(a[k], b[k]), k          # one thread per cell, no atomics: bitwise reproducible
    dev = host.to("cuda:0")
    eng = cj.Engine(cfg)
    st = torch.cuda.Stream()
    inp, out = dev.update_bundles()
    with torch.cuda.stream(st):
        eng.update_state(inp, out, QUERY_TIME, st)
    st.synchronize()
    c = dev.outputs()
    for k in c:
        assert np.array_equal(a[k], c[k]), k


@pytest.mark.parametrize("bits", [64, 32])
def test_reference_facing_model_api(bits):
    """OceanSeaIceModel / time_step! / update_state! mirror drives the same kernels."""
    import torch
    grid, host, cfg = make_case(64, 32, 8, bits)
    ref = oracle_update(host, cfg, time=0.0)
    dev = host.to("cuda:0")
    ocean = cj.ocean_simulation(grid, dev)
    atmosphere = cj.PrescribedAtmosphere(dev)
    model = cj.OceanSeaIceModel(ocean, atmosphere=atmosphere)        # constructor ends with update_state!
    torch.cuda.synchronize()
    compare(dev.outputs(), ref, bits)
    assert model.interfaces.net_fluxes.ocean.T is dev.net["T"]
    cj.time_step(model, 1200.0)
    torch.cuda.synchronize()
    assert model.clock.time == 1200.0 and model.clock.iteration == 1
    ref2 = oracle_update(host, cfg, time=1200.0)
    compare(dev.outputs(), ref2, bits)
    with pytest.raises(ValueError, match="Unknown flux_configuration"):
        cj.build_coupled_model(ocean, None, atmosphere, None, None, "shear_aware")


@pytest.mark.parametrize("bits", [64, 32])
def test_host_buffer_entry_matches_device_path(bits):
    import torch
    grid, host, cfg = make_case(96, 44, 6, bits)
    ref = oracle_update(host, cfg)
    dev = host.to("cuda:0")
    eng = cj.Engine(cfg)
    H, kN = grid.halo[0], grid.Nz - 1 + grid.halo[2]
    planes = {n: torch.from_numpy(np.ascontiguousarray(host.ocean[n].data[kN])).pin_memory() for n in ("u", "v", "T", "S")}
    outs = {n: torch.empty_like(planes["u"]).pin_memory() for n in ("u", "v", "T", "S", "Qv", "Qc")}
    step = _abi.HostStep(planes["u"].data_ptr(), planes["v"].data_ptr(), planes["T"].data_ptr(), planes["S"].data_ptr(),
                         outs["u"].data_ptr(), outs["v"].data_ptr(), outs["T"].data_ptr(), outs["S"].data_ptr(),
                         outs["Qv"].data_ptr(), outs["Qc"].data_ptr(), H, 0)
    h2d, d2h = eng.update_state_host(dev.atmos_series(), step, QUERY_TIME)
    esz = 8 if bits == 64 else 4
    plane_bytes = (grid.Nx + 2 * H) * (grid.Ny + 2 * H) * esz
    assert h2d == 4 * plane_bytes and d2h == 6 * plane_bytes
    for n, key in (("u", "net.u"), ("v", "net.v"), ("T", "net.T"), ("S", "net.S"), ("Qv", "ao.latent_heat"), ("Qc", "ao.sensible_heat")):
        got = outs[n].numpy()[H:-H, H:-H]
        assert rel_err(got, ref[key], bits) <= RTOL[bits], key


@pytest.mark.parametrize("flux_configuration", ["default", "ncar"])
@pytest.mark.parametrize("bits", [64, 32])
def test_mixed_parent_layouts_take_the_strided_kernel(bits, flux_configuration):
    """The tile kernel shares one element offset among all 2-D surface arrays (uniform layout: what Oceananigans parents on
    one grid have).  A caller whose fields are padded differently is still served — by the one-cell-per-thread kernel,
    which addresses every array through its own strides — with the same results."""
    import torch
    from climaocean.jl_b200.fields import Field
    grid, host, cfg = make_case(72, 30, 4, bits, land_fraction=0.2, flux_configuration=flux_configuration)
    ref = oracle_update(host, cfg)
    uniform, _ = gpu_update(host, cfg)
    dev = host.to("cuda:0")
    size2 = (grid.Nx, grid.Ny, 1)
    dev.net["T"] = Field.zeros(size2, (9, 8, 0), dev.dtype, "cuda:0", "net_T_wide")                 # wider halo → other row pitch
    dev.ao["latent_heat"] = Field.zeros(size2, (7, 11, 0), dev.dtype, "cuda:0", "Qv_tall")
    eng = cj.Engine(cfg)
    inp, out = dev.update_bundles()
    eng.update_state(inp, out, QUERY_TIME)
    torch.cuda.synchronize()
    mixed = dev.outputs()
    compare(mixed, ref, bits)
    for k in ("net.T", "ao.latent_heat", "net.S", "net.u"):
        assert mixed[k].shape == uniform[k].shape
        assert rel_err(mixed[k], uniform[k], bits) <= RTOL[bits], k
        if flux_configuration == "ncar":       # one and the same kernel, two ways of addressing: the same bits
            assert np.array_equal(mixed[k], uniform[k]), k
    eng.close()


@pytest.mark.parametrize("bits", [64, 32])
def test_series_with_their_own_padding_take_the_strided_kernel(bits):
    """One gather-offset set serves all nine atmosphere series only when they share a layout; a series padded differently
    (here: the short-wave radiation with a wider halo) sends the call to the strided kernel — same results."""
    import torch
    from climaocean.jl_b200.fields import FieldTimeSeries
    grid, host, cfg = make_case(72, 30, 4, bits)
    ref = oracle_update(host, cfg)
    uniform, _ = gpu_update(host, cfg)
    dev = host.to("cuda:0")
    q = dev.atmos["Qs"]
    H = q.halo[0]
    wide = torch.nn.functional.pad(q.data, (3, 3, 2, 2))                      # (Nt, 1, nj + 4, ni + 6)
    dev.atmos["Qs"] = FieldTimeSeries(wide.contiguous(), (H + 3, q.halo[1] + 2, 0), q.times, "Qs_wide")
    eng = cj.Engine(cfg)
    inp, out = dev.update_bundles()
    eng.update_state(inp, out, QUERY_TIME)
    torch.cuda.synchronize()
    mixed = dev.outputs()
    compare(mixed, ref, bits)
    for k in ("net.T", "exchange.Qs", "net.downwelling_shortwave"):
        assert rel_err(mixed[k], uniform[k], bits) <= RTOL[bits], k
    eng.close()

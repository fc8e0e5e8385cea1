"""GPU parity (through the C ABI) of the round-2 rows against the CPU oracle: land freshwater (JRA55PrescribedLand),
wind rotation on a curvilinear grid, compute_net_sea_ice_fluxes! with the prescribed and the CCSM3 albedo, time-averaged
flux diagnostics accumulated in the kernels' epilogues, and the device forcing ring with prefetch.
Reference call sites: /root/reference/src/OMIPConfigurations/atmosphere.jl:22-46, omip_diagnostics.jl:77-89,125-158,
examples/one_degree_tripolar_ocean_sea_ice.jl:17-42.  Needs a B200."""
import numpy as np
import pytest

import climaocean.jl_b200 as cj
from climaocean.jl_b200 import _abi
from climaocean.jl_b200.fields import Field
from oracle import pyoracle
from tests.common import QUERY_TIME, RTOL, compare, compare_sea_ice, gpu_update, make_case, oracle_update, rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("bits", [64, 32])
@pytest.mark.parametrize("flux_configuration", ["default", "ncar"])
def test_land_freshwater_fused_and_stand_alone(bits, flux_configuration):
    import torch
    grid, host, cfg = make_case(72, 34, 4, bits, with_land=True, flux_configuration=flux_configuration)
    t = 1.3 * 86400.0                                    # exercises the land series' own (daily) time weights
    ref = oracle_update(host, cfg, time=t)
    gpu, dev = gpu_update(host, cfg, time=t)
    compare(gpu, ref, bits)
    assert np.any(gpu["exchange.Mp"] != 0)
    # stand-alone: interpolate_atmosphere_state! then the land kernel; same arithmetic as the fused phase A → same bits
    eng = cj.Engine(cfg)
    x = dev.exchange_state()
    for f in dev.exchange.values():
        f.data.zero_()
    eng.interpolate_atmosphere_state(dev.atmos_series(), t, x)
    eng.interpolate_land(dev.land_series(), t, x)
    torch.cuda.synchronize()
    assert np.array_equal(dev.outputs()["exchange.Mp"], gpu["exchange.Mp"])
    # and the land really contributes
    land, host.land = host.land, None
    ref_noland = oracle_update(host, cfg, time=t)
    host.land = land
    assert np.max(np.abs(ref_noland["net.S"] - ref["net.S"])) > 0


@pytest.mark.parametrize("bits", [64, 32])
def test_wind_rotation_into_the_grid_frame(bits):
    """cos θ / sin θ of a curvilinear (tripolar) grid: the interpolated winds are rotated before they are stored and used."""
    grid, host, cfg = make_case(64, 30, 3, bits)
    ring = 1
    jj, ii = np.meshgrid(np.arange(-ring, grid.Ny + ring), np.arange(-ring, grid.Nx + ring), indexing="ij")
    theta = 0.9 * np.sin(2 * np.pi * ii / grid.Nx) * (jj / grid.Ny) ** 2 + 0.2     # up to ~1 rad near the "north fold"
    host.rotation = (Field(np.cos(theta)[None].astype(grid.dtype), (ring, ring, 0), "cos_theta"),
                     Field(np.sin(theta)[None].astype(grid.dtype), (ring, ring, 0), "sin_theta"))
    ref = oracle_update(host, cfg)
    gpu, _ = gpu_update(host, cfg)
    compare(gpu, ref, bits)
    host.rotation = None
    unrot = oracle_update(host, cfg)
    assert np.max(np.abs(unrot["exchange.u"] - ref["exchange.u"])) > 1.0           # the rotation really acts
    # rotation preserves the wind speed
    assert np.allclose(np.hypot(ref["exchange.u"], ref["exchange.v"]), np.hypot(unrot["exchange.u"], unrot["exchange.v"]),
                       rtol=1e-12 if bits == 64 else 1e-5)


@pytest.mark.parametrize("bits", [64, 32])
@pytest.mark.parametrize("albedo", ["prescribed", "ccsm3"])
def test_net_sea_ice_fluxes(bits, albedo):
    import torch
    grid, host, cfg = make_case(60, 26, 6, bits, with_ice=True, frazil=True, land_fraction=0.15)
    if albedo == "ccsm3":
        cfg.radiation.sea_ice_albedo_kind = _abi.SEA_ICE_ALBEDO_CCSM3
    dev = host.to("cuda:0")
    dt = 900.0
    # oracle: a3 → a7 → a8 → net sea-ice fluxes
    pyoracle.interpolate_atmosphere(cfg, host.atmos_series(), QUERY_TIME, host.exchange_state())
    pyoracle.atmosphere_sea_ice_fluxes(cfg, host.exchange_state(), host.ocean_surface(), host.sea_ice_state(), host.interface_fluxes("ai"))
    pyoracle.sea_ice_ocean_fluxes(cfg, host.ocean_columns(), host.sea_ice_state(), dt, host.ice_ocean_fluxes())
    pyoracle.assemble_net_sea_ice_fluxes(cfg, host.exchange_state(), host.ocean_surface(), host.sea_ice_state(),
                                         host.interface_fluxes("ai"), host.ice_ocean_fluxes(), host.net_sea_ice_fluxes())
    eng = cj.Engine(cfg)
    eng.interpolate_atmosphere_state(dev.atmos_series(), QUERY_TIME, dev.exchange_state())
    eng.compute_atmosphere_sea_ice_fluxes(dev.exchange_state(), dev.ocean_surface(), dev.sea_ice_state(), dev.interface_fluxes("ai"))
    eng.compute_sea_ice_ocean_fluxes(dev.ocean_columns(), dev.sea_ice_state(), dt, dev.ice_ocean_fluxes())
    n0 = eng.launches
    eng.compute_net_sea_ice_fluxes(dev.exchange_state(), dev.ocean_surface(), dev.sea_ice_state(), dev.interface_fluxes("ai"),
                                   dev.ice_ocean_fluxes(), dev.net_sea_ice_fluxes())
    torch.cuda.synchronize()
    assert eng.launches == n0 + 1
    ref, gpu = host.outputs(), dev.outputs()
    # the atmosphere–sea-ice solve feeds this kernel: converged cells at the north_star tolerance, limit-cycle cells as in
    # test_atmosphere_sea_ice_fluxes (tests/common.py::compare_sea_ice)
    its_ref = host.iterations_ai.numpy()[0, 7:-7, 7:-7]
    its_gpu = dev.iterations_ai.numpy()[0, 7:-7, 7:-7]
    compare_sea_ice(gpu, ref, bits, its_gpu, its_ref, cfg.atmosphere_sea_ice.max_iterations,
                    [k for k in ref if k.startswith("net_ice.")], stress_keys=("net_ice.top_u", "net_ice.top_v"))
    assert np.any(gpu["net_ice.top_heat"] != 0) and np.any(gpu["net_ice.bottom_heat"] != 0) and np.any(gpu["net_ice.top_u"] != 0)
    if albedo == "ccsm3":      # the albedo changed the absorbed short wave, hence the skin temperature and the top flux
        cfg2 = cj.default_config(60, 26, 6, bits)
        assert cfg2.radiation.sea_ice_albedo_kind == _abi.SEA_ICE_ALBEDO_PRESCRIBED


@pytest.mark.parametrize("bits", [64, 32])
def test_time_averaged_fluxes_accumulate_in_the_kernel_epilogues(bits):
    """Three coupled steps of different length inside one averaging window: the running averages kept by the fused path
    (flux kernel + stress kernel + ice–ocean kernel epilogues) equal the oracle's WindowedTimeAverage of the same fields."""
    import torch
    grid, host, cfg = make_case(56, 24, 8, bits, with_ice=True, frazil=True)
    dev = host.to("cuda:0")
    host.allocate_averages()
    dev.allocate_averages()
    eng = cj.Engine(cfg)
    T = 0.0
    for step, dt in enumerate((600.0, 1500.0, 900.0)):
        t = QUERY_TIME + 1800.0 * step
        # oracle: the un-fused sequence, then the stand-alone accumulation
        pyoracle.sea_ice_ocean_fluxes(cfg, host.ocean_columns(), host.sea_ice_state(), dt, host.ice_ocean_fluxes())
        oracle_update(host, cfg, time=t, with_ice_terms=True)
        pyoracle.accumulate_flux_averages(cfg, host.net_ocean_fluxes(), host.interface_fluxes("ao"), host.sea_ice_state(),
                                          host.ice_ocean_fluxes(), host.flux_averages(T, dt))
        # CUDA: averages attached → no extra launch
        eng.attach_flux_averages(dev.flux_averages(T, dt))
        n0 = eng.launches
        eng.compute_sea_ice_ocean_fluxes(dev.ocean_columns(), dev.sea_ice_state(), dt, dev.ice_ocean_fluxes())
        inp, out = dev.update_bundles(with_ice_terms=True)
        eng.update_state(inp, out, t)
        assert eng.launches == n0 + 3
        T += dt
    torch.cuda.synchronize()
    ref, gpu = host.outputs(), dev.outputs()
    compare(gpu, ref, bits, keys=[k for k in ref if k.startswith("avg.")])
    for k in ("avg.tau_x", "avg.JT", "avg.JS", "avg.Qc", "avg.Qv", "avg.JT_atmosphere_ocean", "avg.JT_ice_ocean", "avg.JT_frazil"):
        assert np.any(gpu[k] != 0), k
    # stand-alone form on the GPU: one launch, same numbers as the oracle for a single collection
    dev.allocate_averages(); host.allocate_averages()
    eng.attach_flux_averages(None)
    eng.accumulate_flux_averages(dev.net_ocean_fluxes(), dev.interface_fluxes("ao"), dev.sea_ice_state(), dev.ice_ocean_fluxes(),
                                 dev.flux_averages(0.0, 300.0))
    pyoracle.accumulate_flux_averages(cfg, host.net_ocean_fluxes(), host.interface_fluxes("ao"), host.sea_ice_state(),
                                      host.ice_ocean_fluxes(), host.flux_averages(0.0, 300.0))
    torch.cuda.synchronize()
    compare(dev.outputs(), host.outputs(), bits, keys=[k for k in ref if k.startswith("avg.")])


@pytest.mark.parametrize("mode", [_abi.TIME_LINEAR, _abi.TIME_CYCLICAL])
def test_device_forcing_ring_is_bit_identical_and_never_stalls(mode):
    """40 coupled steps crossing ≥ 5 window moves: the ring (capacity 4, one level prefetched) gives the same bits as the
    whole series resident on the device, every level is uploaded once, and no step waits for an upload."""
    import torch
    Nt = 12
    grid, host, cfg = make_case(160, 72, 3, 64, Nt=Nt)
    host.time_indexing = mode
    full = host.to("cuda:0")
    ring = host.to("cuda:0")
    eng_full, eng_ring = cj.Engine(cfg), cj.Engine(cfg)
    source = {n: f.numpy() for n, f in host.atmos.items()}
    halo = host.atmos["u"].halo
    w = cj.DeviceForcingWindow(eng_ring, source, host.times, capacity=4, halo=halo, time_indexing=mode, prefetch=1)
    st = torch.cuda.Stream()
    dt_level = host.times[1] - host.times[0]
    span = (Nt - 1) * dt_level if mode == _abi.TIME_LINEAR else 1.6 * Nt * dt_level       # cyclical: run past the wrap
    nsteps = 40
    times = [0.013 * dt_level + span * k / nsteps for k in range(nsteps)]
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(nsteps)]
    moves, last_start = 0, None
    with torch.cuda.stream(st):
        for k, t in enumerate(times):
            teff = w.apply(ring, t, st)
            if last_start is not None and ring.ring_start != last_start:
                moves += 1
            last_start = ring.ring_start
            inp, out = ring.update_bundles()
            ev[k][0].record(st)
            eng_ring.update_state(inp, out, teff, st)
            ev[k][1].record(st)
            w.release(st)
            # reference: whole series on the device, same step
            inp2, out2 = full.update_bundles()
            eng_full.update_state(inp2, out2, t, st)
            st.synchronize()
            a, b = ring.outputs(), full.outputs()
            for key in b:
                assert np.array_equal(a[key], b[key]), (k, key)
    assert moves >= 5
    nbytes, nlevels = w.stats()
    plane = int(np.prod(source["u"].shape[1:])) * 8 * len(source)
    assert nbytes == nlevels * plane
    needed = len({int(t // dt_level) for t in times} | {int(t // dt_level) + 1 for t in times})
    assert nlevels <= needed + 2                              # every level once (+ the prefetched ones at the end)
    ms = np.array([a.elapsed_time(b) for a, b in ev])
    assert ms[5:].max() < 5.0 * np.median(ms[5:]) + 0.2, ms   # no step stalls behind an upload
    w.close()


@pytest.mark.parametrize("bits", [64])
def test_coupled_model_mirror_with_land_and_sea_ice(bits):
    """build_coupled_model(ocean, sea_ice, atmosphere, radiation, land, :corrected): land is used (it was dropped),
    SeaIceAlbedo selects the CCSM3 form, update_state! ends with compute_net_sea_ice_fluxes!."""
    import torch
    grid, host, cfg0 = make_case(48, 22, 5, bits, with_ice=True, with_land=True, frazil=True)
    dev = host.to("cuda:0")
    ocean, sea_ice = cj.ocean_simulation(grid, dev), cj.sea_ice_simulation(grid, dev)
    atmosphere, land = cj.PrescribedAtmosphere(dev), cj.PrescribedLand(dev)
    radiation = cj.Radiation(ocean_surface=cj.SurfaceRadiationProperties(0.06, 1.0),
                             sea_ice_surface=cj.SurfaceRadiationProperties(cj.SeaIceAlbedo(), 1.0))
    model = cj.build_coupled_model(ocean, sea_ice, atmosphere, radiation, land, "corrected")
    torch.cuda.synchronize()
    cfg = model.interfaces.cfg
    assert cfg.radiation.sea_ice_albedo_kind == _abi.SEA_ICE_ALBEDO_CCSM3
    # oracle, same sequence as update_state!
    t = 0.0
    pyoracle.interpolate_atmosphere(cfg, host.atmos_series(), t, host.exchange_state())
    pyoracle.interpolate_land(cfg, host.land_series(), t, host.exchange_state())
    pyoracle.atmosphere_ocean_fluxes(cfg, host.exchange_state(), host.ocean_surface(), host.interface_fluxes("ao"))
    pyoracle.atmosphere_sea_ice_fluxes(cfg, host.exchange_state(), host.ocean_surface(), host.sea_ice_state(), host.interface_fluxes("ai"))
    pyoracle.sea_ice_ocean_fluxes(cfg, host.ocean_columns(), host.sea_ice_state(), 1.0, host.ice_ocean_fluxes())
    pyoracle.assemble_net_ocean_fluxes(cfg, host.exchange_state(), host.ocean_surface(), host.interface_fluxes("ao"),
                                       host.sea_ice_state(), host.ice_ocean_fluxes(), host.net_ocean_fluxes())
    pyoracle.assemble_net_sea_ice_fluxes(cfg, host.exchange_state(), host.ocean_surface(), host.sea_ice_state(),
                                         host.interface_fluxes("ai"), host.ice_ocean_fluxes(), host.net_sea_ice_fluxes())
    gpu, ref = dev.outputs(), host.outputs()
    ice_keys = [k for k in ref if k.startswith("ai.") or k.startswith("net_ice.")]
    compare(gpu, ref, bits, keys=[k for k in ref if k not in ice_keys])
    compare_sea_ice(gpu, ref, bits, dev.iterations_ai.numpy()[0, 7:-7, 7:-7], host.iterations_ai.numpy()[0, 7:-7, 7:-7],
                    cfg.atmosphere_sea_ice.max_iterations, ice_keys, stress_keys=("net_ice.top_u", "net_ice.top_v"))
    assert model.interfaces.net_fluxes.sea_ice.top.heat is dev.net_ice["top_heat"]
    with pytest.raises(NotImplementedError):
        cj.ComponentInterfaces(atmosphere, ocean, sea_ice, land="rivers.nc")


@pytest.mark.parametrize("bits", [64, 32])
def test_config5_one_degree_tripolar_ocean_sea_ice_flux_set(bits):
    """BASELINE config 5: 1° TripolarGrid (360×180) ocean + sea ice — the whole flux set of update_state! (atmosphere–ocean,
    atmosphere–sea-ice, sea-ice–ocean, net ocean, net sea-ice) with winds rotated into the grid frame and north halos
    filled by the fold (examples/one_degree_tripolar_ocean_sea_ice.jl:17-42; OceanConfigurations/one_degree_tripolar.jl:20-73)."""
    import torch
    grid = cj.TripolarGrid((360, 180, 6), halo=(7, 7, 7), dtype=np.float64 if bits == 64 else np.float32)
    host = cj.SurfaceFluxData.synthetic(grid, with_ice=True, with_land=True, frazil=True, land_fraction=0.3, ring=1)
    assert host.rotation is not None
    cfg = cj.default_config(360, 180, 6, bits, "corrected")
    cfg.grid.ring = 1
    cfg.radiation.sea_ice_albedo_kind = _abi.SEA_ICE_ALBEDO_CCSM3
    dev = host.to("cuda:0")
    t, dt = QUERY_TIME, 1200.0

    def sequence(d, e):
        x, o, ice, io = d.exchange_state(), d.ocean_surface(), d.sea_ice_state(), d.ice_ocean_fluxes()
        ao, ai = d.interface_fluxes("ao"), d.interface_fluxes("ai")
        e.interpolate_atmosphere(d.atmos_series(), t, x)
        e.interpolate_land(d.land_series(), t, x)
        e.atmosphere_ocean(x, o, ao)
        e.atmosphere_sea_ice(x, o, ice, ai)
        e.sea_ice_ocean(d.ocean_columns(), ice, dt, io)
        e.net_ocean(x, o, ao, ice, io, d.net_ocean_fluxes())
        e.net_sea_ice(x, o, ice, ai, io, d.net_sea_ice_fluxes())

    class Oracle:
        interpolate_atmosphere = staticmethod(lambda s, tt, x: pyoracle.interpolate_atmosphere(cfg, s, tt, x))
        interpolate_land = staticmethod(lambda s, tt, x: pyoracle.interpolate_land(cfg, s, tt, x))
        atmosphere_ocean = staticmethod(lambda x, o, f: pyoracle.atmosphere_ocean_fluxes(cfg, x, o, f))
        atmosphere_sea_ice = staticmethod(lambda x, o, i, f: pyoracle.atmosphere_sea_ice_fluxes(cfg, x, o, i, f))
        sea_ice_ocean = staticmethod(lambda c, i, d_, f: pyoracle.sea_ice_ocean_fluxes(cfg, c, i, d_, f))
        net_ocean = staticmethod(lambda x, o, ao, i, io, n: pyoracle.assemble_net_ocean_fluxes(cfg, x, o, ao, i, io, n))
        net_sea_ice = staticmethod(lambda x, o, i, ai, io, n: pyoracle.assemble_net_sea_ice_fluxes(cfg, x, o, i, ai, io, n))

    eng = cj.Engine(cfg)

    class Cuda:
        interpolate_atmosphere = staticmethod(lambda s, tt, x: eng.interpolate_atmosphere_state(s, tt, x))
        interpolate_land = staticmethod(lambda s, tt, x: eng.interpolate_land(s, tt, x))
        atmosphere_ocean = staticmethod(lambda x, o, f: eng.compute_atmosphere_ocean_fluxes(x, o, f))
        atmosphere_sea_ice = staticmethod(lambda x, o, i, f: eng.compute_atmosphere_sea_ice_fluxes(x, o, i, f))
        sea_ice_ocean = staticmethod(lambda c, i, d_, f: eng.compute_sea_ice_ocean_fluxes(c, i, d_, f))
        net_ocean = staticmethod(lambda x, o, ao, i, io, n: eng.compute_net_ocean_fluxes(x, o, ao, i, io, n))
        net_sea_ice = staticmethod(lambda x, o, i, ai, io, n: eng.compute_net_sea_ice_fluxes(x, o, i, ai, io, n))

    sequence(host, Oracle)
    sequence(dev, Cuda)
    torch.cuda.synchronize()
    assert eng.launches == 7
    ref, gpu = host.outputs(), dev.outputs()
    ice_keys = [k for k in ref if k.startswith("ai.") or k.startswith("net_ice.")]
    compare(gpu, ref, bits, keys=[k for k in ref if k not in ice_keys and not k.startswith("avg.")])
    compare_sea_ice(gpu, ref, bits, dev.iterations_ai.numpy()[0, 7:-7, 7:-7], host.iterations_ai.numpy()[0, 7:-7, 7:-7],
                    cfg.atmosphere_sea_ice.max_iterations, ice_keys, stress_keys=("net_ice.top_u", "net_ice.top_v"))
    # the rotation acts in the cap only, and strongly near the fold
    cs = host.rotation[0].numpy()[0, 1:-1, 1:-1]
    assert np.all(cs[:grid.j0 - 1] > 1 - 1e-6) and cs[grid.j0 + 2:].min() < 0.1
    assert np.any(gpu["net_ice.top_heat"] != 0) and np.any(gpu["io.frazil_heat"] != 0) and np.any(gpu["net.S"] != 0)

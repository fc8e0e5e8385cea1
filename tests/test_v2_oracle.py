"""Known-answer tests (CPU) of the round-2 additions to the oracle: land freshwater, CCSM3 sea-ice albedo, net sea-ice
fluxes, running time averages, ring-buffer series.  The oracle is the checker of the GPU tests; these pin it on closed
forms.  Reference call sites: /root/reference/src/OMIPConfigurations/atmosphere.jl:31-46, jra55_data_staging.jl:8,
omip_diagnostics.jl:77-89,125-158."""
import ctypes as C

import numpy as np
import pytest

import climaocean.jl_b200 as cj
from climaocean.jl_b200 import _abi
from oracle import pyoracle
from tests.common import QUERY_TIME, make_case, oracle_update


def _interior(f):
    a = f.numpy()
    Hx, Hy, _ = f.halo
    return a[0, Hy:a.shape[1] - Hy, Hx:a.shape[2] - Hx]


def test_land_freshwater_adds_to_the_exchange_freshwater_flux():
    grid, host, cfg = make_case(24, 12, 3, 64, with_land=True)
    # constant fields are reproduced exactly by bilinear × linear-in-time interpolation (weights sum to one)
    host.land["rivers"].data[...] = 3.0e-5
    host.land["icebergs"].data[...] = 1.0e-5
    ref0 = dict(oracle_update(host, cfg))                       # with land
    host_noland = host
    land = host.land
    host.land = None
    ref1 = dict(oracle_update(host_noland, cfg))                # without
    host.land = land
    d = ref0["exchange.Mp"] - ref1["exchange.Mp"]
    assert np.allclose(d, 4.0e-5, rtol=1e-12, atol=0)
    # it enters the salinity flux like rain: J^S changes by +S·ΔMp/ρ_f (fresher ocean ⇒ upward salt-equivalent flux … sign as A9)
    S = host.ocean["S"].numpy()[-1 - grid.halo[2], grid.halo[1]:-grid.halo[1], grid.halo[0]:-grid.halo[0]]
    assert np.allclose(ref0["net.S"] - ref1["net.S"], S * 4.0e-5 / 1000.0, rtol=1e-9)
    # the heat flux does not see it
    assert np.array_equal(ref0["net.T"], ref1["net.T"])


def test_land_interpolation_is_linear_in_time_and_space():
    grid, host, cfg = make_case(16, 8, 2, 64, with_land=True)
    nt = host.land["rivers"].data.shape[0]
    for n in range(nt):
        host.land["rivers"].data[n] = 1e-5 * (n + 1)
    host.land["icebergs"].data[...] = 0.0
    x = host.exchange_state()
    for f in host.exchange.values():
        f.data[...] = 0.0
    t = 0.25 * host.land_times[1] + 0.75 * host.land_times[2]
    pyoracle.interpolate_land(cfg, host.land_series(), t, x)
    assert np.allclose(_interior(host.exchange["Mp"]), 1e-5 * (0.25 * 2 + 0.75 * 3), rtol=1e-13)


def _ccsm3(cfg, hi, hs, TsC):
    """CCSM3 albedo through the net-sea-ice entry point: top = −(1−α)Qs when everything else is zero."""
    grid, host, _ = make_case(4, 3, 2, 64, with_ice=True)
    cfg.grid.Nx, cfg.grid.Ny, cfg.grid.Nz = 4, 3, 2
    for f in host.exchange.values():
        f.data[...] = 0.0
    host.exchange["Qs"].data[...] = 100.0
    for f in list(host.ai.values()) + list(host.io.values()):
        f.data[...] = 0.0
    host.ice["thickness"].data[...] = hi
    host.ice["snow_thickness"].data[...] = hs
    host.ice["top_temperature"].data[...] = TsC
    host.ice["concentration"].data[...] = 1.0
    cfg.radiation.stefan_boltzmann_constant = 0.0               # no emitted long wave: isolates the short-wave term
    pyoracle.assemble_net_sea_ice_fluxes(cfg, host.exchange_state(), host.ocean_surface(), host.sea_ice_state(),
                                         host.interface_fluxes("ai"), host.ice_ocean_fluxes(), host.net_sea_ice_fluxes())
    top = _interior(host.net_ice["top_heat"])
    return 1.0 + top[0, 0] / 100.0


def test_ccsm3_albedo_limits():
    cfg = cj.default_config(4, 3, 2, 64)
    cfg.radiation.sea_ice_albedo_kind = _abi.SEA_ICE_ALBEDO_CCSM3
    vis, nir, f = 0.78, 0.36, 0.52
    # thick, cold, bare ice: the band-weighted bare-ice albedo
    assert _ccsm3(cfg, 2.0, 0.0, -20.0) == pytest.approx(f * vis + (1 - f) * nir, rel=1e-14)
    # at the melting point bare ice is 0.075 darker in both bands
    assert _ccsm3(cfg, 2.0, 0.0, 0.0) == pytest.approx(f * vis + (1 - f) * nir - 0.075, rel=1e-13)
    # half way through the 1.5 K melt range: half the darkening
    assert _ccsm3(cfg, 2.0, 0.0, -0.75) == pytest.approx(f * vis + (1 - f) * nir - 0.0375, rel=1e-12)
    # vanishing ice fades into the ocean albedo
    assert _ccsm3(cfg, 1e-12, 0.0, -20.0) == pytest.approx(0.06, abs=1e-10)
    # thickness dependence below h_max: atan(4h)/atan(4·0.3)
    fh = np.arctan(4 * 0.1) / np.arctan(1.2)
    assert _ccsm3(cfg, 0.1, 0.0, -20.0) == pytest.approx(f * (vis * fh + 0.06 * (1 - fh)) + (1 - f) * (nir * fh + 0.06 * (1 - fh)), rel=1e-13)
    # deep cold snow: the snow albedo; snow fraction h_s/(h_s + 0.02)
    fs = 0.5 / 0.52
    bare = f * vis + (1 - f) * nir
    snow = f * 0.98 + (1 - f) * 0.70
    assert _ccsm3(cfg, 2.0, 0.5, -20.0) == pytest.approx((1 - fs) * bare + fs * snow, rel=1e-13)
    # prescribed kind ignores all of this
    cfg.radiation.sea_ice_albedo_kind = _abi.SEA_ICE_ALBEDO_PRESCRIBED
    assert _ccsm3(cfg, 0.1, 0.3, -1.0) == pytest.approx(0.7, rel=1e-14)


def test_net_sea_ice_fluxes_identities():
    grid, host, cfg = make_case(20, 10, 3, 64, with_ice=True, land_fraction=0.2)
    rng = np.random.default_rng(3)
    for grp in (host.ai, host.io):
        for f in grp.values():
            f.data[...] = rng.normal(size=f.data.shape)
    pyoracle.interpolate_atmosphere(cfg, host.atmos_series(), QUERY_TIME, host.exchange_state())
    pyoracle.assemble_net_sea_ice_fluxes(cfg, host.exchange_state(), host.ocean_surface(), host.sea_ice_state(),
                                         host.interface_fluxes("ai"), host.ice_ocean_fluxes(), host.net_sea_ice_fluxes())
    top, bot = _interior(host.net_ice["top_heat"]), _interior(host.net_ice["bottom_heat"])
    conc, wet = _interior(host.ice["concentration"]), _interior(host.mask) != 0
    assert np.all(top[conc == 0] == 0)                                  # no ice, no top flux
    assert np.all(top[~wet] == 0) and np.all(bot[~wet] == 0)            # land
    Qf, Qi = _interior(host.io["frazil_heat"]), _interior(host.io["interface_heat"])
    assert np.array_equal(bot[wet], (Qf + Qi)[wet])
    # top = Qd + Qu + Qc + Qv with the constant default albedo 0.7, ε = 1
    Ts = _interior(host.ice["top_temperature"]) + 273.15
    Qs, Ql = _interior(host.exchange["Qs"]), _interior(host.exchange["Ql"])
    want = -(1 - 0.7) * Qs - Ql + 5.67e-8 * Ts ** 4 + _interior(host.ai["sensible_heat"]) + _interior(host.ai["latent_heat"])
    sel = wet & (conc > 0)
    assert np.allclose(top[sel], want[sel], rtol=1e-13)
    # stresses: plain centre → face averages of ρτ
    rtx = host.ai["x_momentum"].numpy()[0]
    Hx, Hy = grid.halo[0], grid.halo[1]
    face = 0.5 * (rtx[Hy:-Hy, Hx - 1:-Hx - 1] + rtx[Hy:-Hy, Hx:-Hx])
    m = host.mask.numpy()[0]
    both = (m[Hy:-Hy, Hx - 1:-Hx - 1] != 0) & wet
    assert np.array_equal(_interior(host.net_ice["top_u"])[both], face[both])


def test_running_average_is_the_time_weighted_mean():
    grid, host, cfg = make_case(12, 6, 2, 64)
    host.allocate_averages()
    rng = np.random.default_rng(5)
    dts = [600.0, 900.0, 300.0, 1200.0]
    acc, T = None, 0.0
    for dt in dts:
        for f in (host.net["u"], host.net["T"], host.ao["sensible_heat"]):
            f.data[...] = rng.normal(size=f.data.shape)
        v = host.flux_averages(T, dt)
        pyoracle.accumulate_flux_averages(cfg, host.net_ocean_fluxes(), host.interface_fluxes("ao"), None, None, v)
        x = np.stack([_interior(host.net["u"]), _interior(host.net["T"]), _interior(host.ao["sensible_heat"])])
        acc = x * dt if acc is None else acc + x * dt
        T += dt
    got = np.stack([_interior(host.averages["tau_x"]), _interior(host.averages["JT"]), _interior(host.averages["Qc"])])
    assert np.allclose(got, acc / T, rtol=1e-13, atol=1e-15)          # WindowedTimeAverage = left Riemann sum / window
    # ocean-only: JTao = JT, the ice parts are zero
    assert np.allclose(_interior(host.averages["JT_atmosphere_ocean"]), _interior(host.averages["JT"]), rtol=1e-13, atol=1e-15)
    assert np.all(_interior(host.averages["JT_ice_ocean"]) == 0) and np.all(_interior(host.averages["JS_ice_ocean"]) == 0)
    # a new window (T = 0) forgets the old one
    v = host.flux_averages(0.0, 450.0)
    pyoracle.accumulate_flux_averages(cfg, host.net_ocean_fluxes(), host.interface_fluxes("ao"), None, None, v)
    assert np.allclose(_interior(host.averages["tau_x"]), _interior(host.net["u"]), rtol=4e-16, atol=0)   # (0·x + u·Δt)/Δt


def test_ring_buffer_series_equals_plain_series():
    """Levels stored at ring slots (start + n) mod capacity give the same interpolation as the plain layout."""
    grid, host, cfg = make_case(16, 8, 2, 64, Nt=6)
    ref = dict(oracle_update(host, cfg, time=2.4 * 10800.0))
    # rotate the time axis of every series by 4 slots in a ring of capacity 6
    start, cap = 4, 6
    for n, f in host.atmos.items():
        f.data = np.ascontiguousarray(np.roll(f.data, start, axis=0))
    host.ring_start, host.ring_capacity = start, cap
    got = dict(oracle_update(host, cfg, time=2.4 * 10800.0))
    for k in ref:
        assert np.array_equal(ref[k], got[k]), k

"""Pin the CPU oracle.  The reference holds no golden vector / known-answer test for this path
(parity unpinned), so the oracle is anchored on (a) mpmath high-precision
restatements of the published formulas, (b) analytic limits of Monin–Obukhov theory, (c) exactness
identities of the interpolation and assembly.  CPU only."""
import ctypes as C

import mpmath as mp
import numpy as np
import pytest

import climaocean.jl_b200 as cj
from climaocean.jl_b200 import _abi
from oracle import pyoracle
from tests.common import QUERY_TIME, make_case, oracle_update

mp.mp.dps = 40
lib = pyoracle.load()


def probe_psi(kind, zeta, bits=64):
    dt = np.float64 if bits == 64 else np.float32
    z = np.ascontiguousarray(zeta, dtype=dt)
    pm, ps = np.empty_like(z), np.empty_like(z)
    f = getattr(lib, f"oracle_probe_psi_f{bits}")
    f(C.c_int(kind), C.c_int(z.size), z.ctypes.data_as(C.c_void_p), pm.ctypes.data_as(C.c_void_p), ps.ctypes.data_as(C.c_void_p))
    return pm, ps


# ---- mpmath restatements (Edson et al. 2013 / COARE 3.5; Paulson 1970; Grachev et al. 2007) ----
def mp_conv(y):
    r3 = mp.sqrt(3)
    return mp.mpf(1.5) * mp.log((1 + y + y * y) / 3) - r3 * mp.atan((1 + 2 * y) / r3) + mp.pi / r3


def mp_edson_m(z):
    z = mp.mpf(z)
    if z >= 0:
        dz = min(mp.mpf(50), mp.mpf("0.35") * z)
        return -(mp.mpf("0.7") * z + mp.mpf("0.75") * (z - 5 / mp.mpf("0.35")) * mp.exp(-dz) + mp.mpf("0.75") * 5 / mp.mpf("0.35"))
    x = (1 - 15 * z) ** mp.mpf("0.25")
    pk = 2 * mp.log((1 + x) / 2) + mp.log((1 + x * x) / 2) - 2 * mp.atan(x) + mp.pi / 2
    pc = mp_conv(mp.cbrt(1 - mp.mpf("10.15") * z))
    f = z * z / (1 + z * z)
    return (1 - f) * pk + f * pc


def mp_edson_s(z):
    z = mp.mpf(z)
    if z >= 0:
        dz = min(mp.mpf(50), mp.mpf("0.35") * z)
        return -((1 + mp.mpf(2) / 3 * z) ** mp.mpf("1.5") + mp.mpf(2) / 3 * (z - mp.mpf("14.28")) * mp.exp(-dz) + mp.mpf("8.525"))
    x = mp.sqrt(1 - 15 * z)
    pk = 2 * mp.log((1 + x) / 2)
    pc = mp_conv(mp.cbrt(1 - mp.mpf("34.15") * z))
    f = z * z / (1 + z * z)
    return (1 - f) * pk + f * pc


def mp_paulson_m(z):
    x = (1 - 16 * mp.mpf(z)) ** mp.mpf("0.25")
    return 2 * mp.log((1 + x) / 2) + mp.log((1 + x * x) / 2) - 2 * mp.atan(x) + mp.pi / 2


def mp_paulson_s(z):
    return 2 * mp.log((1 + mp.sqrt(1 - 16 * mp.mpf(z))) / 2)


def mp_sheba_m(z):
    z = mp.mpf(z)
    if z < 0:
        return mp_paulson_m(z)
    a, b = mp.mpf(5), mp.mpf(5) / mp.mpf("6.5")          # Grachev et al. (2007): a_m = 5, b_m = a_m / 6.5
    x, B, r3 = mp.cbrt(1 + z), mp.cbrt((1 - b) / b), mp.sqrt(3)
    return -3 * a * (x - 1) / b + a * B / (2 * b) * (2 * mp.log((x + B) / (1 + B)) - mp.log((x * x - B * x + B * B) / (1 - B + B * B))
                                                      + 2 * r3 * (mp.atan((2 * x - B) / (r3 * B)) - mp.atan((2 - B) / (r3 * B))))


def mp_sheba_s(z):
    z = mp.mpf(z)
    if z < 0:
        return mp_paulson_s(z)
    a, b, c = mp.mpf(5), mp.mpf(5), mp.mpf(3)
    B = mp.sqrt(c * c - 4)
    return -b / 2 * mp.log(1 + c * z + z * z) + (-a / B + b * c / (2 * B)) * (mp.log((2 * z + c - B) / (2 * z + c + B)) - mp.log((c - B) / (c + B)))


ZETAS = np.array([-50.0, -10.0, -2.5, -1.0, -0.3, -1e-2, -1e-6, 0.0, 1e-6, 1e-2, 0.3, 1.0, 2.5, 10.0, 50.0, 200.0])


@pytest.mark.parametrize("kind,fm,fs", [(_abi.STABILITY_EDSON, mp_edson_m, mp_edson_s),
                                        (_abi.STABILITY_SHEBA_PAULSON, mp_sheba_m, mp_sheba_s),
                                        (_abi.STABILITY_LARGE_YEAGER, lambda z: -5 * mp.mpf(z) if z >= 0 else mp_paulson_m(z),
                                         lambda z: -5 * mp.mpf(z) if z >= 0 else mp_paulson_s(z))])
def test_stability_functions_match_mpmath(kind, fm, fs):
    pm, ps = probe_psi(kind, ZETAS)
    for z, a, b in zip(ZETAS, pm, ps):
        # the literals of the oracle are binary doubles (0.35, 10.15 ...): compare at 1e-13 relative, abs floor 1e-14
        rm, rs = float(fm(float(z))), float(fs(float(z)))
        assert abs(a - rm) <= 2e-13 * max(1.0, abs(rm)), (kind, z, a, rm)
        assert abs(b - rs) <= 2e-13 * max(1.0, abs(rs)), (kind, z, b, rs)


def test_sheba_momentum_function_is_the_integral_of_its_gradient_function():
    """φ_m = 1 − ζ dψ_m/dζ must equal 1 + 6.5 ζ (1+ζ)^{1/3} / (1.3 + ζ) (Grachev et al. 2007, eq. 9a)."""
    for z in (0.1, 1.0, 5.0, 30.0):
        d = mp.diff(mp_sheba_m, z)
        phi = 1 + mp.mpf("6.5") * z * mp.cbrt(1 + mp.mpf(z)) / (mp.mpf("1.3") + z)
        assert abs((1 - z * d) - phi) < mp.mpf("1e-20")


def test_stability_functions_vanish_at_neutral_and_are_continuous():
    for kind in (_abi.STABILITY_EDSON, _abi.STABILITY_SHEBA_PAULSON, _abi.STABILITY_LARGE_YEAGER, _abi.STABILITY_NEUTRAL):
        pm, ps = probe_psi(kind, np.array([0.0, -1e-12, 1e-12]))
        assert np.all(np.abs(pm) < 1e-10)
        if kind == _abi.STABILITY_EDSON:
            # COARE's stable scalar function has a published −0.005 offset at ζ = 0⁺: −1 + ⅔·14.28 − 8.525
            assert abs(ps[1]) < 1e-10 and ps[0] == pytest.approx(-0.005, abs=1e-12) and ps[2] == pytest.approx(-0.005, abs=1e-9)
        else:
            assert np.all(np.abs(ps) < 1e-10)


def test_stability_float32_close_to_float64():
    for kind in (_abi.STABILITY_EDSON, _abi.STABILITY_SHEBA_PAULSON, _abi.STABILITY_LARGE_YEAGER):
        pm64, ps64 = probe_psi(kind, ZETAS, 64)
        pm32, ps32 = probe_psi(kind, ZETAS, 32)
        assert np.allclose(pm32, pm64, rtol=2e-5, atol=2e-5)
        assert np.allclose(ps32, ps64, rtol=2e-5, atol=2e-5)


def test_saturation_vapor_pressure_matches_mpmath_and_triple_point():
    cfg = cj.default_config(4, 4, 2)
    T = np.array([233.0, 250.0, 273.16, 288.15, 303.15, 310.0])
    pl, pi = np.empty_like(T), np.empty_like(T)
    lib.oracle_probe_saturation_f64(C.byref(cfg), C.c_int(T.size), T.ctypes.data_as(C.c_void_p), pl.ctypes.data_as(C.c_void_p), pi.ctypes.data_as(C.c_void_p))
    t = cfg.atmosphere.thermodynamics
    Rv = mp.mpf(t.gas_constant) / mp.mpf(t.water_molar_mass)

    def ref(Tk, LH0, dcp):
        Tk = mp.mpf(Tk)
        return mp.mpf(t.triple_point_pressure) * (Tk / mp.mpf(t.triple_point_temperature)) ** (dcp / Rv) * mp.exp(
            (LH0 - dcp * mp.mpf(t.reference_temperature)) / Rv * (1 / mp.mpf(t.triple_point_temperature) - 1 / Tk))
    for k, Tk in enumerate(T):
        rl = float(ref(Tk, mp.mpf(t.reference_vaporization_enthalpy), mp.mpf(t.water_vapor_heat_capacity) - mp.mpf(t.liquid_water_heat_capacity)))
        ri = float(ref(Tk, mp.mpf(t.reference_sublimation_enthalpy), mp.mpf(t.water_vapor_heat_capacity) - mp.mpf(t.ice_heat_capacity)))
        assert abs(pl[k] - rl) <= 1e-13 * rl and abs(pi[k] - ri) <= 1e-13 * ri
    assert pl[2] == pytest.approx(611.657, rel=1e-15) and pi[2] == pytest.approx(611.657, rel=1e-15)
    assert 1690 < pl[3] < 1720      # ≈ 17 hPa at 15 °C
    assert pi[1] < pl[1]            # ice saturation below liquid saturation under freezing


def test_thermodynamic_state_dry_and_supersaturated():
    cfg = cj.default_config(4, 4, 2)
    out = (C.c_double * 7)()
    lib.oracle_probe_thermo_f64.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_void_p]
    lib.oracle_probe_thermo_f64(C.byref(cfg), 101325.0, 288.15, 0.0, out)
    Rd = 8.3144598 / 0.02897
    assert out[0] == pytest.approx(101325.0 / (Rd * 288.15), rel=1e-14)          # dry-air density
    assert out[1] == pytest.approx(Rd / (2.0 / 7.0), rel=1e-14)                    # cp_d
    assert out[2] == 0.0 and out[4] == 0.0 and out[5] == 0.0
    lib.oracle_probe_thermo_f64(C.byref(cfg), 101325.0, 260.0, 0.02, out)           # far above saturation at 260 K
    assert out[4] + out[5] > 0.015 and out[2] < 0.005                               # condensate removed from vapour
    assert out[2] + out[4] + out[5] == pytest.approx(0.02, rel=1e-13)


def solve(cfg, ua, va, Ta, pa, qa, uo, vo, ToK, So, bits=64):
    dt = np.float64 if bits == 64 else np.float32
    x = np.array([ua, va, Ta, pa, qa, uo, vo, ToK, So], dtype=dt)
    o = np.zeros(4, dtype=dt)
    getattr(lib, f"oracle_probe_solve_f{bits}")(C.byref(cfg), x.ctypes.data_as(C.c_void_p), o.ctypes.data_as(C.c_void_p))
    return o


def neutral_cfg():
    cfg = cj.default_config(4, 4, 2)
    p = cfg.atmosphere_ocean
    p.stability_functions = _abi.STABILITY_NEUTRAL
    p.gustiness_parameter = 0.0
    p.momentum_roughness.kind = _abi.ROUGHNESS_FIXED
    p.momentum_roughness.fixed_length = 1e-4
    p.temperature_roughness.kind = _abi.ROUGHNESS_FIXED
    p.temperature_roughness.fixed_length = 1e-5
    p.water_vapor_roughness.kind = _abi.ROUGHNESS_FIXED
    p.water_vapor_roughness.fixed_length = 2e-5
    return cfg


def test_neutral_limit_is_the_log_law():
    """ψ ≡ 0, fixed roughness, no gustiness: u★ = κU/ln(h/ℓu), θ★ = κΔθ/ln(h/ℓθ), q★ = κΔq/ln(h/ℓq) exactly."""
    cfg = neutral_cfg()
    ua, va, uo, vo = 7.0, -3.0, 0.4, 0.2
    o = solve(cfg, ua, va, 285.0, 101000.0, 0.006, uo, vo, 290.0, 35.0)
    U = np.hypot(ua - uo, va - vo)
    assert o[0] == pytest.approx(0.4 * U / np.log(10.0 / 1e-4), rel=1e-14)
    assert o[3] == 2                                   # second pass reproduces the first exactly → drift 0
    # ratios are independent of the thermodynamics: θ★/q★ · (Δq/Δθ) = ln(h/ℓq)/ln(h/ℓθ)
    o2 = solve(cfg, ua, va, 285.0, 101000.0, 0.006, uo, vo, 291.0, 35.0)
    assert o2[0] == pytest.approx(o[0], rel=1e-15)     # u★ does not depend on ΔT in the neutral limit


def test_wind_velocity_formulation_ignores_ocean_current():
    cfg = neutral_cfg()
    cfg.atmosphere_ocean.velocity_formulation = _abi.VELOCITY_WIND
    a = solve(cfg, 7.0, -3.0, 285.0, 101000.0, 0.006, 0.4, 0.2, 290.0, 35.0)
    b = solve(cfg, 7.0, -3.0, 285.0, 101000.0, 0.006, -0.9, 0.7, 290.0, 35.0)
    assert np.array_equal(a, b)
    assert a[0] == pytest.approx(0.4 * np.hypot(7.0, 3.0) / np.log(1e5), rel=1e-14)


def test_zero_differences_give_zero_scales():
    cfg = cj.default_config(4, 4, 2)
    # Δu = Δv = 0 with gustiness switched off → calm-cell guard: all scales vanish
    cfg.atmosphere_ocean.gustiness_parameter = 0.0
    o = solve(cfg, 0.3, -0.1, 285.0, 101000.0, 0.006, 0.3, -0.1, 290.0, 35.0)
    assert o[0] == 0.0 and o[1] == 0.0 and o[2] == 0.0


def test_default_solver_converges_to_a_fixed_point_and_respects_maxiter():
    cfg = cj.default_config(4, 4, 2)
    o = solve(cfg, 9.0, 2.0, 283.0, 100800.0, 0.005, 0.1, 0.0, 291.0, 34.0)
    assert 3 <= o[3] < 100 and o[0] > 0 and o[1] < 0 and o[2] < 0     # unstable: ocean warmer & moister than the air
    cfg.atmosphere_ocean.max_iterations = int(o[3]) + 5                 # a larger cap must not change the answer
    o2 = solve(cfg, 9.0, 2.0, 283.0, 100800.0, 0.005, 0.1, 0.0, 291.0, 34.0)
    assert np.array_equal(o, o2)
    cfg.atmosphere_ocean.max_iterations = 3
    o3 = solve(cfg, 9.0, 2.0, 283.0, 100800.0, 0.005, 0.1, 0.0, 291.0, 34.0)
    assert o3[3] == 3
    cfg.atmosphere_ocean.stop_kind = _abi.STOP_FIXED_ITERATIONS
    cfg.atmosphere_ocean.max_iterations = 40
    o4 = solve(cfg, 9.0, 2.0, 283.0, 100800.0, 0.005, 0.1, 0.0, 291.0, 34.0)
    assert o4[3] == 40 and np.allclose(o4[:3], o[:3], rtol=1e-7)        # converged value is the fixed point


def test_stable_vs_unstable_stratification_orders_the_drag():
    cfg = cj.default_config(4, 4, 2)
    unstable = solve(cfg, 8.0, 0.0, 283.0, 101000.0, 0.004, 0.0, 0.0, 293.0, 35.0)
    stable = solve(cfg, 8.0, 0.0, 298.0, 101000.0, 0.012, 0.0, 0.0, 283.0, 35.0)
    assert unstable[0] > stable[0] > 0
    assert stable[1] > 0 and unstable[1] < 0


def test_large_yeager_neutral_drag_at_10m():
    cfg = cj.default_config(4, 4, 2, flux_configuration="ncar")
    # Δθ = Δq ≈ 0 is not reachable exactly; use weak stratification and compare Cd with the LY polynomial within 5 %
    o = solve(cfg, 10.0, 0.0, 288.15, 101325.0, 0.0104, 0.0, 0.0, 288.16, 35.0)
    cd = (o[0] / 10.0) ** 2
    cdn = 1e-3 * (2.7 / 10 + 0.142 + 10 / 13.09 - 3.14807e-10 * 10 ** 6)
    assert o[3] == 5
    assert cd == pytest.approx(cdn, rel=0.05)


def test_time_indices_linear_cyclical_clamp():
    times = np.arange(8, dtype=np.float64) * 10800.0
    tp = times.ctypes.data_as(C.POINTER(C.c_double))
    n1, n2, fr = C.c_int32(), C.c_int32(), C.c_double()

    def q(mode, t, period=0.0):
        assert lib.oracle_time_indices(tp, 8, mode, period, t, C.byref(n1), C.byref(n2), C.byref(fr)) == 0
        return n1.value, n2.value, fr.value
    assert q(_abi.TIME_LINEAR, QUERY_TIME) == (1, 2, pytest.approx(0.37))
    assert q(_abi.TIME_LINEAR, 0.0) == (0, 1, 0.0)
    assert q(_abi.TIME_LINEAR, 7 * 10800.0) == (6, 7, 1.0)
    assert q(_abi.TIME_LINEAR, 8 * 10800.0)[2] == pytest.approx(2.0)          # linear extrapolation
    assert q(_abi.TIME_CLAMP, 8 * 10800.0) == (6, 7, 1.0)
    assert q(_abi.TIME_CLAMP, -5.0) == (0, 1, 0.0)
    assert q(_abi.TIME_CYCLICAL, 7.5 * 10800.0) == (7, 0, pytest.approx(0.5))   # wrap interval last → first
    assert q(_abi.TIME_CYCLICAL, 8 * 10800.0 + QUERY_TIME) == (1, 2, pytest.approx(0.37))
    assert q(_abi.TIME_CYCLICAL, -10800.0 * 0.5) == (7, 0, pytest.approx(0.5))


def test_interpolation_reproduces_constant_and_linear_fields():
    grid, host, cfg = make_case(48, 24, 4)
    # constant series → constant on the ocean grid, exactly (weights sum to one up to rounding)
    host.atmos["T"].data[...] = 287.5
    # series linear in source index i, j and in time
    a = host.atmos["p"].data
    Nt, _, nj, ni = a.shape
    ii, jj = np.arange(ni) - 3.0, np.arange(nj) - 3.0
    for n in range(Nt):
        a[n, 0] = 1.0e5 + 2.0 * ii[None, :] + 3.0 * jj[:, None] + 5.0 * n
    out = oracle_update(host, cfg)
    assert np.max(np.abs(out["exchange.T"] - 287.5)) <= 6e-14 * 287.5
    fi = host.fi.data[0, 1:-1, 1:-1].astype(np.float64)
    fj = host.fj.data[0, 1:-1, 1:-1].astype(np.float64)
    frac = QUERY_TIME / 10800.0
    expect = 1.0e5 + 2.0 * fi + 3.0 * fj + 5.0 * frac
    # the periodic seam of the source grid is not linear in i: exclude cells whose stencil crosses it
    ok = fi < (ni - 6 - 1)
    assert np.max(np.abs(out["exchange.p"] - expect)[ok]) <= 1e-9
    # Mp = rain + snow summed after interpolation
    host.atmos["rain"].data[...] = 1e-4
    host.atmos["snow"].data[...] = 2e-4
    out = oracle_update(host, cfg)
    assert np.max(np.abs(out["exchange.Mp"] - 3e-4)) <= 1e-18 + 1e-15 * 3e-4


def test_assembly_identities():
    grid, host, cfg = make_case(32, 16, 4, with_ice=True)
    rho0, c0 = cfg.ocean.reference_density, cfg.ocean.heat_capacity
    out = oracle_update(host, cfg)                                   # ice terms not passed: ℵ ≡ 0
    SQ = out["net.upwelling_longwave"] + out["net.downwelling_longwave"] + out["ao.sensible_heat"] + out["ao.latent_heat"]
    assert np.allclose(out["net.T"] * rho0 * c0, SQ, rtol=1e-13, atol=1e-10)   # W/m² = Jᵀ ρ₀ c₀ (visualize/cache.jl:359-361)
    assert np.allclose(out["net.penetrating_shortwave"] * rho0 * c0, out["net.downwelling_shortwave"], rtol=1e-13)
    assert np.allclose(out["net.downwelling_shortwave"], -(1 - 0.06) * out["exchange.Qs"], rtol=1e-15)
    S = host.ocean["S"].data[grid.Nz - 1 + 7, 7:-7, 7:-7]
    F = (-out["exchange.Mp"] + out["ao.water_vapor"]) / 1000.0
    assert np.allclose(out["net.S"], -S * F, rtol=1e-13, atol=1e-20)
    # stress: face value is the mean of the two adjacent centre values divided by ρ₀
    rtx = host.ao["x_momentum"].data[0, 7:-7, 6:-7]
    assert np.allclose(out["net.u"], 0.5 * (rtx[:, :-1] + rtx[:, 1:]) / rho0, rtol=1e-13, atol=1e-20)
    # full ice cover (ℵ = 1) removes every atmosphere–ocean contribution
    host.ice["concentration"].data[...] = 1.0
    for f in host.io.values():
        f.data[...] = 0.0
    out1 = oracle_update(host, cfg, with_ice_terms=True)
    for k in ("net.T", "net.S", "net.u", "net.v", "net.penetrating_shortwave"):
        assert np.all(out1[k] == 0.0), k


def test_minimum_salinity_suppresses_freshening_only():
    grid, host, cfg = make_case(16, 8, 2)
    cfg.ocean.minimum_salinity = 100.0                                  # every cell is "too fresh"
    out = oracle_update(host, cfg)
    assert np.all(out["net.S"] <= 0.0)                                  # salt-extracting (positive) fluxes suppressed
    assert np.any(out["net.S"] < 0.0) or np.all(out["exchange.Mp"] > out["ao.water_vapor"])


def test_land_cells_produce_zero_fluxes():
    grid, host, cfg = make_case(32, 16, 4, land_fraction=0.3)
    out = oracle_update(host, cfg)
    wet = host.mask.data[0, 7:-7, 7:-7] != 0
    assert 0.5 < wet.mean() < 0.9
    for k in ("ao.latent_heat", "ao.sensible_heat", "ao.x_momentum", "net.T", "net.S"):
        assert np.all(out[k][~wet] == 0.0), k
    assert np.any(out["ao.latent_heat"][wet] != 0.0)


def test_frazil_heat_budget_and_ice_bath():
    grid, host, cfg = make_case(24, 12, 6, with_ice=True, frazil=True)
    T_before = host.ocean["T"].data.copy()
    dt = 600.0
    pyoracle.sea_ice_ocean_fluxes(cfg, host.ocean_columns(), host.sea_ice_state(), dt, host.ice_ocean_fluxes())
    T_after = host.ocean["T"].data
    S = host.ocean["S"].data
    Tm = 0.0 - 0.054 * S
    sl = (slice(7, -7), slice(7, -7), slice(7, -7))
    assert np.all(T_after[sl] >= Tm[sl] - 1e-15)                         # nothing below freezing afterwards
    changed = T_after[sl] != T_before[sl]
    assert changed.any() and np.all(T_before[sl][changed] < Tm[sl][changed])
    dz = grid.dz()[7]
    heat = cfg.ocean.reference_density * cfg.ocean.heat_capacity * ((T_after - T_before)[sl] * dz).sum(axis=0) / dt
    Qf = host.io["frazil_heat"].data[0, 7:-7, 7:-7]
    assert np.allclose(Qf, -heat, rtol=1e-12, atol=1e-9)                 # frazil heat = −(column heating)/Δt
    # ice bath: Q_io = ρ₀c₀ u_m★ (T_N − T_m) ℵ
    TN, SN = T_after[grid.Nz - 1 + 7, 7:-7, 7:-7], S[grid.Nz - 1 + 7, 7:-7, 7:-7]
    conc = host.ice["concentration"].data[0, 7:-7, 7:-7]
    expect = cfg.ocean.reference_density * cfg.ocean.heat_capacity * 1e-5 * (TN - (0.0 - 0.054 * SN)) * conc
    assert np.allclose(host.io["interface_heat"].data[0, 7:-7, 7:-7], expect, rtol=1e-12, atol=1e-12)
    # previous thickness is rolled forward
    assert np.array_equal(host.ice["previous_thickness"].data[0, 7:-7, 7:-7], host.ice["thickness"].data[0, 7:-7, 7:-7])


def test_three_equation_interface_satisfies_its_equations():
    grid, host, cfg = make_case(16, 8, 4, with_ice=True, flux_configuration="corrected")
    io = cfg.ice_ocean
    assert io.heat_flux == _abi.ICE_OCEAN_THREE_EQUATION and io.friction_velocity == _abi.FRICTION_VELOCITY_MOMENTUM_BASED
    io.friction_velocity = _abi.FRICTION_VELOCITY_CONSTANT
    pyoracle.sea_ice_ocean_fluxes(cfg, host.ocean_columns(), host.sea_ice_state(), 600.0, host.ice_ocean_fluxes())
    rho0, c0 = cfg.ocean.reference_density, cfg.ocean.heat_capacity
    TN = host.ocean["T"].data[grid.Nz - 1 + 7, 7:-7, 7:-7]
    SN = host.ocean["S"].data[grid.Nz - 1 + 7, 7:-7, 7:-7]
    conc = host.ice["concentration"].data[0, 7:-7, 7:-7]
    Si = host.ice["salinity"].data[0, 7:-7, 7:-7]
    Q = host.io["interface_heat"].data[0, 7:-7, 7:-7]
    gT, gS = io.heat_transfer_coefficient * io.constant_friction_velocity, io.salt_transfer_coefficient * io.constant_friction_velocity
    ok = conc > 0
    Tb = TN[ok] - Q[ok] / conc[ok] / (rho0 * c0 * gT)                   # invert the heat equation
    Sb = -Tb / io.liquidus_slope                                         # liquidus
    w = rho0 * c0 * gT * (TN[ok] - Tb) / (io.ice_density * io.ice_latent_heat)
    assert np.allclose(gS * (SN[ok] - Sb), w * (Sb - Si[ok]), rtol=1e-9, atol=1e-14)   # salt balance closes
    assert np.all(Sb > 0)

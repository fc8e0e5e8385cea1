"""50-digit (mpmath) restatement of the WHOLE per-cell interface solve of SURVEY.md Appendix A (A1–A7): thermodynamic
states, surface humidity, stability functions, roughness lengths, the fixed-point iteration with its stop rule, the
Large–Yeager coefficient form, and the skin-temperature update of the atmosphere–sea-ice solve.

TEST INFRASTRUCTURE.  Written from the appendix, independently of oracle/oracle_impl.h (different language, different
arithmetic, no shared code); tests/test_oracle_whole_cell_mpmath.py holds the C oracle to 1e-13 of it, which pins
"oracle = Appendix A evaluated exactly".  It does NOT pin the oracle to NumericalEarth (parity stays unpinned).
Every parameter is read from the same coflux_config the oracle gets; binary doubles are converted exactly."""
import mpmath as mp

from climaocean.jl_b200 import _abi

mp.mp.dps = 50
F = mp.mpf


class Consts:
    def __init__(self, t):
        self.R_d = F(t.gas_constant) / F(t.dry_air_molar_mass)
        self.R_v = F(t.gas_constant) / F(t.water_molar_mass)
        self.eps = F(t.dry_air_molar_mass) / F(t.water_molar_mass)
        self.cp_d = self.R_d / F(t.dry_air_adiabatic_exponent)
        self.cp_v, self.cp_l, self.cp_i = F(t.water_vapor_heat_capacity), F(t.liquid_water_heat_capacity), F(t.ice_heat_capacity)
        self.LH_v0, self.LH_s0 = F(t.reference_vaporization_enthalpy), F(t.reference_sublimation_enthalpy)
        self.T_0, self.T_tr, self.p_tr = F(t.reference_temperature), F(t.triple_point_temperature), F(t.triple_point_pressure)
        self.T_fr, self.T_in = F(t.water_freezing_temperature), F(t.total_ice_nucleation_temperature)


def p_sat(c, T, LH_0, dcp):                                   # A1, Clausius–Clapeyron with constant Δcp
    return c.p_tr * (T / c.T_tr) ** (dcp / c.R_v) * mp.exp((LH_0 - dcp * c.T_0) / c.R_v * (1 / c.T_tr - 1 / T))


def liquid_fraction(c, T):
    if T > c.T_fr:
        return F(1)
    if T <= c.T_in:
        return F(0)
    return (T - c.T_in) / (c.T_fr - c.T_in)


def moist_air(c, p, T, q):                                    # A1: state from (p, T, q_tot) with saturation adjustment
    lam = liquid_fraction(c, T)
    LH_0 = lam * c.LH_v0 + (1 - lam) * c.LH_s0
    dcp = lam * (c.cp_v - c.cp_l) + (1 - lam) * (c.cp_v - c.cp_i)
    ps = p_sat(c, T, LH_0, dcp)
    q_vs = (c.R_d / c.R_v) * (1 - q) * ps / (p - ps) if p - ps > 0 else mp.inf
    q_c = max(q - q_vs, F(0))
    q_liq, q_ice = lam * q_c, (1 - lam) * q_c
    R_m = c.R_d * (1 + (c.eps - 1) * q - c.eps * q_c)
    return dict(rho=p / (R_m * T), cp=c.cp_d + (c.cp_v - c.cp_d) * q + (c.cp_l - c.cp_v) * q_liq + (c.cp_i - c.cp_v) * q_ice,
                q_vap=q - q_liq - q_ice, T_v=T * R_m / c.R_d)


# ---- A5 stability functions (literals are the binary doubles the C code holds) ----
def _conv(y):
    r3 = mp.sqrt(3)
    return F(1.5) * mp.log((1 + y + y * y) / 3) - r3 * mp.atan((1 + 2 * y) / r3) + mp.pi / r3


def psi_m(kind, z):
    if kind == _abi.STABILITY_NEUTRAL:
        return F(0)
    if kind == _abi.STABILITY_EDSON:
        if z >= 0:
            dz = min(F(50), F(0.35) * z)
            return -F(0.7) * z - F(0.75) * (z - F(5) / F(0.35)) * mp.exp(-dz) - F(0.75) * F(5) / F(0.35)
        x = mp.sqrt(mp.sqrt(1 - 15 * z))
        pk = 2 * mp.log((1 + x) / 2) + mp.log((1 + x * x) / 2) - 2 * mp.atan(x) + mp.pi / 2
        pc = _conv(mp.cbrt(1 - F(10.15) * z))
        f = z * z / (1 + z * z)
        return (1 - f) * pk + f * pc
    if z < 0:                                                  # Paulson (1970), coefficient 16
        x = mp.sqrt(mp.sqrt(1 - 16 * z))
        return 2 * mp.log((1 + x) / 2) + mp.log((1 + x * x) / 2) - 2 * mp.atan(x) + mp.pi / 2
    if kind == _abi.STABILITY_LARGE_YEAGER:
        return -5 * z
    a, b = F(5), F(5) / F(6.5)                                 # Grachev et al. (2007)
    x, B, r3 = mp.cbrt(1 + z), mp.cbrt((1 - b) / b), mp.sqrt(3)
    return -3 * a * (x - 1) / b + a * B / (2 * b) * (2 * mp.log((x + B) / (1 + B)) - mp.log((x * x - B * x + B * B) / (1 - B + B * B))
                                                      + 2 * r3 * (mp.atan((2 * x - B) / (r3 * B)) - mp.atan((2 - B) / (r3 * B))))


def psi_s(kind, z):
    if kind == _abi.STABILITY_NEUTRAL:
        return F(0)
    if kind == _abi.STABILITY_EDSON:
        if z >= 0:
            dz = min(F(50), F(0.35) * z)
            return -(1 + F(2) / F(3) * z) ** F(1.5) - F(2) / F(3) * (z - F(14.28)) * mp.exp(-dz) - F(8.525)
        x = mp.sqrt(1 - 15 * z)
        pk = 2 * mp.log((1 + x) / 2)
        pc = _conv(mp.cbrt(1 - F(34.15) * z))
        f = z * z / (1 + z * z)
        return (1 - f) * pk + f * pc
    if z < 0:
        return 2 * mp.log((1 + mp.sqrt(1 - 16 * z)) / 2)
    if kind == _abi.STABILITY_LARGE_YEAGER:
        return -5 * z
    a, b, c = F(5), F(5), F(3)
    B = mp.sqrt(c * c - 4)
    return -b / 2 * mp.log(1 + c * z + z * z) + (-a / B + b * c / (2 * B)) * (mp.log((2 * z + c - B) / (2 * z + c + B)) - mp.log((c - B) / (c + B)))


# ---- A6 roughness lengths ----
def viscosity(v, T):
    if v.kind == _abi.VISCOSITY_CONSTANT:
        return F(v.nu)
    Tp = T - F(273.15)
    return F(v.c0) + F(v.c1) * Tp + F(v.c2) * Tp * Tp + F(v.c3) * Tp * Tp * Tp


def momentum_roughness(r, ustar, U, Ts):
    if r.kind == _abi.ROUGHNESS_FIXED:
        return F(r.fixed_length)
    nu = viscosity(r.viscosity, Ts)
    if r.wave_formulation == _abi.WAVES_WIND_DEPENDENT:
        alpha = max(F(r.wind_a1) * min(U, F(r.wind_umax)) + F(r.wind_a2), F(r.wind_alpha_min))
    else:
        alpha = F(r.gravity_wave_parameter)
    lR = F(r.maximum_length) if ustar == 0 else F(r.smooth_wall_parameter) * nu / ustar
    return min(alpha * ustar * ustar / F(r.gravitational_acceleration) + lR, F(r.maximum_length))


def scalar_roughness(r, lu, ustar, Ts):
    if r.kind == _abi.ROUGHNESS_FIXED:
        return F(r.fixed_length)
    Rstar = lu * ustar / viscosity(r.viscosity, Ts)
    lq = F(0) if Rstar == 0 else F(r.reynolds_A) / Rstar ** F(r.reynolds_b)
    return min(lq, F(r.maximum_length))


def profile(form, stab, scalar, h, l, L):
    ps = psi_s if scalar else psi_m
    if l == 0:                                                 # ℓ = 0 (R★ = 0): ln(h/ℓ) = +∞ ⇒ χ = κ/∞ = 0
        return mp.inf
    chi = mp.log(h / l) - ps(stab, h / L)
    if form == _abi.PROFILE_LOGARITHMIC:
        chi += ps(stab, l / L)
    return chi


def ly_cdn(U):
    if U >= 33:
        return F(2.34e-3)
    return F(1e-3) * (F(2.7) / U + F(0.142) + U / F(13.09) - F(3.14807e-10) * U ** 6)


def solve_cell(cfg, P, cell, surface_kind):
    """cell: dict ua va Ta pa qa us vs Ts So [Qs Ql h_ice S_ice albedo].  Returns (u★, θ★, q★, T_s, iterations)."""
    c = Consts(cfg.atmosphere.thermodynamics)
    A = cfg.atmosphere
    g, h, hbl = F(A.gravitational_acceleration), F(A.surface_layer_height), F(A.boundary_layer_height)
    kappa = F(P.von_karman_constant)
    v = {k: F(float(x)) for k, x in cell.items()}
    atm = moist_air(c, v["pa"], v["Ta"], v["qa"])
    if P.velocity_formulation == _abi.VELOCITY_RELATIVE:
        du, dv = v["ua"] - v["us"], v["va"] - v["vs"]
    else:
        du, dv = v["ua"], v["va"]
    x = F(1)
    if surface_kind == 0:
        o = cfg.ocean
        s = v["So"] / 1000
        alpha = F(o.salt_water_molar_mass) * sum(F(o.constituent_mass_fraction[k]) / F(o.constituent_molar_mass[k]) for k in range(4))
        x = (1 - s) / (1 - s + alpha * s)
    theta_a = v["Ta"] + g * h / atm["cp"]
    Ts = v["Ts"]
    us = ts = qs = F(P.initial_scale)
    ly = P.formulation == _abi.FLUXES_COEFFICIENT_LARGE_YEAGER

    def surface(Ts):
        if surface_kind == 0:
            ps = p_sat(c, Ts, c.LH_v0, c.cp_v - c.cp_l)
        else:
            ps = p_sat(c, Ts, c.LH_s0, c.cp_v - c.cp_i)
        qsurf = ps / (atm["rho"] * c.R_v * Ts) * x
        return qsurf, moist_air(c, v["pa"], Ts, qsurf)

    U_ly = rcdn_ly = F(0)
    if ly:
        qsurf, _ = surface(Ts)
        U_ly = max(mp.sqrt(du * du + dv * dv), F(P.ly_minimum_wind))
        dth0, dq0 = theta_a - Ts, atm["q_vap"] - qsurf
        cdn = ly_cdn(U_ly)
        rcdn = mp.sqrt(cdn)
        chn = (F(18e-3) if dth0 > 0 else F(32.7e-3)) * rcdn
        cen = F(34.6e-3) * rcdn
        rcdn_ly = rcdn
        us, ts, qs = rcdn * U_ly, chn / rcdn * dth0, cen / rcdn * dq0
    it, maxit, tol = 0, P.max_iterations, F(P.tolerance)
    prev = None
    while True:
        if P.stop_kind == _abi.STOP_FIXED_ITERATIONS:
            go = it < maxit
        else:
            if prev is None:
                go = True
            else:
                drift = abs(us - prev[0]) + abs(ts - prev[1]) + abs(qs - prev[2])
                go = not (drift < tol or it >= maxit)
        if not go:
            break
        prev = (us, ts, qs)
        if P.interface_temperature == _abi.TEMPERATURE_SKIN:      # A7 / row a7: conductive flux balance, clamped
            I, R = cfg.ice_ocean, cfg.radiation
            Toff = F(273.15) if cfg.ocean.temperature_units == _abi.TEMPERATURE_CELSIUS else F(0)
            Tb = F(I.liquidus_freshwater_melting_temperature) - F(I.liquidus_slope) * v["S_ice"] + Toff
            Tm = F(I.liquidus_freshwater_melting_temperature) + Toff
            Ls = c.LH_s0 + (c.cp_v - c.cp_i) * (v["Ta"] - c.T_0)
            emis, sigma = F(R.sea_ice_emissivity), F(R.stefan_boltzmann_constant)
            Qa = (-atm["rho"] * Ls * us * qs) + emis * sigma * Ts ** 4 + (-atm["rho"] * atm["cp"] * us * ts) + \
                 (-(1 - v["albedo"]) * v["Qs"] - emis * v["Ql"])
            k_ice = F(I.ice_conductivity)
            Tstar = Tb - Qa * v["h_ice"] / k_ice
            if P.skin_temperature_update == _abi.SKIN_LINEARIZED_LONGWAVE:
                Qrest = Qa - emis * sigma * Ts ** 4
                Tstar = (Tb - Qrest * v["h_ice"] / k_ice) / (1 + sigma * emis * Ts ** 3 / k_ice * v["h_ice"])
            Tstar = max(F(0), Tstar)
            Tnew = Tstar if v["h_ice"] >= F(I.ice_consolidation_thickness) else Tb
            dT = Tnew - Ts
            step = min(F(P.skin_max_delta_T), abs(dT)) * mp.sign(dT)
            Ts = min(Ts + step, Tm)
        qsurf, surf = surface(Ts)
        dq, dth = atm["q_vap"] - qsurf, theta_a - Ts
        bstar = g / surf["T_v"] * (ts * (1 + (c.eps - 1) * surf["q_vap"]) + (c.eps - 1) * surf["T_v"] * qs)
        if ly:
            zeta = max(F(-10), min(F(10), kappa * bstar * h / (us * us)))
            pm_, ph_ = psi_m(_abi.STABILITY_LARGE_YEAGER, zeta), psi_s(_abi.STABILITY_LARGE_YEAGER, zeta)
            lnh = mp.log(h / 10)
            U10N = max(U_ly / (1 + rcdn_ly / kappa * (lnh - pm_)), F(P.ly_minimum_wind))
            cdn = ly_cdn(U10N)
            rcdn = mp.sqrt(cdn)
            cen = F(34.6e-3) * rcdn
            chn = (F(18e-3) if zeta > 0 else F(32.7e-3)) * rcdn
            xm = 1 + rcdn / kappa * (lnh - pm_)
            cd = cdn / (xm * xm)
            ch = chn / (1 + chn / (kappa * rcdn) * (lnh - ph_)) * mp.sqrt(cd / cdn)
            ce = cen / (1 + cen / (kappa * rcdn) * (lnh - ph_)) * mp.sqrt(cd / cdn)
            rcdn_ly = rcdn
            rcd = mp.sqrt(cd)
            us, ts, qs = rcd * U_ly, ch / rcd * dth, ce / rcd * dq
        else:
            Jb = -us * bstar
            UG = max(F(P.gustiness_parameter) * mp.cbrt(Jb * hbl) if Jb >= 0 else -F(P.gustiness_parameter) * mp.cbrt(-Jb * hbl),
                     F(P.minimum_gustiness))
            U = mp.sqrt(du * du + dv * dv + UG * UG)
            if U == 0:
                us = ts = qs = F(0)
            else:
                lu = momentum_roughness(P.momentum_roughness, us, U, Ts)
                lq = scalar_roughness(P.water_vapor_roughness, lu, us, Ts)
                lt = scalar_roughness(P.temperature_roughness, lu, us, Ts)
                L = mp.inf if bstar == 0 else us * us / (kappa * bstar)
                pu = profile(P.similarity_form, P.stability_functions, False, h, lu, L)
                if not pu > 0:
                    us = ts = qs = F(0)
                else:
                    pt = profile(P.similarity_form, P.stability_functions, True, h, lt, L)
                    pq = profile(P.similarity_form, P.stability_functions, True, h, lq, L)
                    us = kappa / pu * U
                    ts = (kappa / pt if pt > 0 else F(0)) * dth
                    qs = (kappa / pq if pq > 0 else F(0)) * dq
        it += 1
    return us, ts, qs, Ts, it

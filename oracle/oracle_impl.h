/* oracle_impl.h — body of the CPU oracle, included twice (FT = double, FT = float).
 *
 * TEST INFRASTRUCTURE ONLY.  See oracle/coflux_oracle.c for the header note (parity unpinned).
 *
 * Written as a literal, un-optimised restatement of SURVEY.md Appendix A (A1..A10): no hoisting,
 * one function per formula, both stability branches spelled out.  The CUDA product code in
 * climaocean.jl_b200/csrc/ is a separate implementation; the two share only include/coflux.h
 * (struct layouts).
 *
 * Macros supplied by the includer: FT, SUF(name), and the libm spellings LOG EXP SQRT CBRT ATAN
 * POW FABS FLOOR FMIN FMAX TRUNC.
 */

typedef struct {
  FT rho, cp_m, q_vap, T_v, T, q_liq, q_ice, q_tot, p;
} SUF(thermo_state);

typedef struct {
  FT R_d, R_v, eps, cp_d, cp_v, cp_l, cp_i, LH_v0, LH_s0, T_0, T_tr, p_tr, T_fr, T_in;
} SUF(thermo_consts);

/* A1: derived thermodynamic constants (atmosphere thermodynamics parameters) */
static SUF(thermo_consts) SUF(make_thermo)(const coflux_thermodynamics* t) {
  SUF(thermo_consts) c;
  c.R_d = (FT)t->gas_constant / (FT)t->dry_air_molar_mass;
  c.R_v = (FT)t->gas_constant / (FT)t->water_molar_mass;
  c.eps = (FT)t->dry_air_molar_mass / (FT)t->water_molar_mass;
  c.cp_d = c.R_d / (FT)t->dry_air_adiabatic_exponent;
  c.cp_v = (FT)t->water_vapor_heat_capacity;
  c.cp_l = (FT)t->liquid_water_heat_capacity;
  c.cp_i = (FT)t->ice_heat_capacity;
  c.LH_v0 = (FT)t->reference_vaporization_enthalpy;
  c.LH_s0 = (FT)t->reference_sublimation_enthalpy;
  c.T_0 = (FT)t->reference_temperature;
  c.T_tr = (FT)t->triple_point_temperature;
  c.p_tr = (FT)t->triple_point_pressure;
  c.T_fr = (FT)t->water_freezing_temperature;
  c.T_in = (FT)t->total_ice_nucleation_temperature;
  return c;
}

/* A1: Clausius–Clapeyron with constant Δcp:  p_tr (T/T_tr)^(Δcp/R_v) exp[(LH0 − Δcp T0)/R_v (1/T_tr − 1/T)] */
static FT SUF(saturation_vapor_pressure_generic)(const SUF(thermo_consts)* c, FT T, FT LH_0, FT dcp) {
  return c->p_tr * POW(T / c->T_tr, dcp / c->R_v) *
         EXP((LH_0 - dcp * c->T_0) / c->R_v * ((FT)1 / c->T_tr - (FT)1 / T));
}
static FT SUF(saturation_vapor_pressure_liquid)(const SUF(thermo_consts)* c, FT T) {
  return SUF(saturation_vapor_pressure_generic)(c, T, c->LH_v0, c->cp_v - c->cp_l);
}
static FT SUF(saturation_vapor_pressure_ice)(const SUF(thermo_consts)* c, FT T) {
  return SUF(saturation_vapor_pressure_generic)(c, T, c->LH_s0, c->cp_v - c->cp_i);
}
/* supercooled-liquid ramp between total ice nucleation and freezing */
static FT SUF(liquid_fraction)(const SUF(thermo_consts)* c, FT T) {
  if (T > c->T_fr) return (FT)1;
  if (T <= c->T_in) return (FT)0;
  return (T - c->T_in) / (c->T_fr - c->T_in);
}
static FT SUF(saturation_vapor_pressure_mixed)(const SUF(thermo_consts)* c, FT T) {
  FT lam = SUF(liquid_fraction)(c, T);
  FT LH_0 = lam * c->LH_v0 + ((FT)1 - lam) * c->LH_s0;
  FT dcp = lam * (c->cp_v - c->cp_l) + ((FT)1 - lam) * (c->cp_v - c->cp_i);
  return SUF(saturation_vapor_pressure_generic)(c, T, LH_0, dcp);
}

/* A1: moist-air state from (p, T, q_tot) with saturation adjustment of the condensate */
static SUF(thermo_state) SUF(phase_equil_pTq)(const SUF(thermo_consts)* c, FT p, FT T, FT q) {
  SUF(thermo_state) s;
  FT lam = SUF(liquid_fraction)(c, T);
  FT ps = SUF(saturation_vapor_pressure_mixed)(c, T);
  FT denom = p - ps;
  FT q_vs = (denom > (FT)0) ? (c->R_d / c->R_v) * ((FT)1 - q) * ps / denom : (FT)HUGE_VAL;
  FT q_c = FMAX(q - q_vs, (FT)0);
  s.q_liq = lam * q_c;
  s.q_ice = ((FT)1 - lam) * q_c;
  s.q_tot = q;
  s.T = T;
  s.p = p;
  FT R_m = c->R_d * ((FT)1 + (c->eps - (FT)1) * q - c->eps * q_c);
  s.rho = p / (R_m * T);
  s.cp_m = c->cp_d + (c->cp_v - c->cp_d) * q + (c->cp_l - c->cp_v) * s.q_liq + (c->cp_i - c->cp_v) * s.q_ice;
  s.q_vap = q - s.q_liq - s.q_ice;
  s.T_v = T * R_m / c->R_d;
  return s;
}
static FT SUF(latent_heat_vapor)(const SUF(thermo_consts)* c, FT T) {
  return c->LH_v0 + (c->cp_v - c->cp_l) * (T - c->T_0);
}
static FT SUF(latent_heat_sublimation)(const SUF(thermo_consts)* c, FT T) {
  return c->LH_s0 + (c->cp_v - c->cp_i) * (T - c->T_0);
}

/* A2: Raoult water mole fraction of sea water */
static FT SUF(water_mole_fraction)(const coflux_ocean_properties* o, FT S) {
  FT s = S / (FT)1000;
  FT alpha = (FT)0;
  for (int k = 0; k < 4; ++k)
    alpha += (FT)o->constituent_mass_fraction[k] / (FT)o->constituent_molar_mass[k];
  alpha = (FT)o->salt_water_molar_mass * alpha;
  return ((FT)1 - s) / ((FT)1 - s + alpha * s);
}
/* A2: q_s = x · p_sat(T_s) / (ρ_a R_v T_s);  phase 0 = liquid (ocean), 1 = ice (x = 1) */
static FT SUF(surface_specific_humidity)(const SUF(thermo_consts)* c, FT rho_a, FT Ts, FT x, int ice_phase) {
  FT ps = ice_phase ? SUF(saturation_vapor_pressure_ice)(c, Ts) : SUF(saturation_vapor_pressure_liquid)(c, Ts);
  FT qstar = ps / (rho_a * c->R_v * Ts);
  return qstar * x;
}

/* A5: stability functions ψ(ζ) ------------------------------------------------------------- */
static FT SUF(psi_edson_momentum)(FT z) {
  if (z >= (FT)0) {
    FT dz = FMIN((FT)50, (FT)0.35 * z);
    return -(FT)0.7 * z - (FT)0.75 * (z - (FT)5 / (FT)0.35) * EXP(-dz) - (FT)0.75 * (FT)5 / (FT)0.35;
  } else {
    FT x = SQRT(SQRT((FT)1 - (FT)15 * z));
    FT psik = (FT)2 * LOG(((FT)1 + x) / (FT)2) + LOG(((FT)1 + x * x) / (FT)2) - (FT)2 * ATAN(x) + (FT)M_PI / (FT)2;
    FT y = CBRT((FT)1 - (FT)10.15 * z);
    FT rt3 = SQRT((FT)3);
    FT psic = (FT)1.5 * LOG(((FT)1 + y + y * y) / (FT)3) - rt3 * ATAN(((FT)1 + (FT)2 * y) / rt3) + (FT)M_PI / rt3;
    FT f = z * z / ((FT)1 + z * z);
    return ((FT)1 - f) * psik + f * psic;
  }
}
static FT SUF(psi_edson_scalar)(FT z) {
  if (z >= (FT)0) {
    FT dz = FMIN((FT)50, (FT)0.35 * z);
    return -POW((FT)1 + (FT)2 / (FT)3 * z, (FT)1.5) - (FT)2 / (FT)3 * (z - (FT)14.28) * EXP(-dz) - (FT)8.525;
  } else {
    FT x = SQRT((FT)1 - (FT)15 * z);
    FT psik = (FT)2 * LOG(((FT)1 + x) / (FT)2);
    FT y = CBRT((FT)1 - (FT)34.15 * z);
    FT rt3 = SQRT((FT)3);
    FT psic = (FT)1.5 * LOG(((FT)1 + y + y * y) / (FT)3) - rt3 * ATAN(((FT)1 + (FT)2 * y) / rt3) + (FT)M_PI / rt3;
    FT f = z * z / ((FT)1 + z * z);
    return ((FT)1 - f) * psik + f * psic;
  }
}
static FT SUF(psi_paulson_momentum)(FT z) { /* unstable branch, z < 0 */
  FT x = SQRT(SQRT((FT)1 - (FT)16 * z));
  return (FT)2 * LOG(((FT)1 + x) / (FT)2) + LOG(((FT)1 + x * x) / (FT)2) - (FT)2 * ATAN(x) + (FT)M_PI / (FT)2;
}
static FT SUF(psi_paulson_scalar)(FT z) {
  FT x = SQRT((FT)1 - (FT)16 * z);
  return (FT)2 * LOG(((FT)1 + x) / (FT)2);
}
static FT SUF(psi_sheba_momentum)(FT z) {
  if (z < (FT)0) return SUF(psi_paulson_momentum)(z);
  /* Grachev et al. (2007) eq. 12: a_m = 5, b_m = a_m/6.5  (φ_m = 1 + 6.5 ζ (1+ζ)^{1/3} / (1.3 + ζ)) */
  const FT a = (FT)5, b = (FT)5 / (FT)6.5;
  FT x = CBRT((FT)1 + z);
  FT B = CBRT(((FT)1 - b) / b);
  FT rt3 = SQRT((FT)3);
  FT p1 = -(FT)3 * a * (x - (FT)1) / b;
  FT p2 = a * B / ((FT)2 * b) *
          ((FT)2 * LOG((x + B) / ((FT)1 + B)) - LOG((x * x - B * x + B * B) / ((FT)1 - B + B * B)) +
           (FT)2 * rt3 * (ATAN(((FT)2 * x - B) / (rt3 * B)) - ATAN(((FT)2 - B) / (rt3 * B))));
  return p1 + p2;
}
static FT SUF(psi_sheba_scalar)(FT z) {
  if (z < (FT)0) return SUF(psi_paulson_scalar)(z);
  const FT a = (FT)5, b = (FT)5, c = (FT)3;
  FT B = SQRT(c * c - (FT)4);
  FT p1 = -b / (FT)2 * LOG((FT)1 + c * z + z * z);
  FT p2 = (-a / B + b * c / ((FT)2 * B)) *
          (LOG(((FT)2 * z + c - B) / ((FT)2 * z + c + B)) - LOG((c - B) / (c + B)));
  return p1 + p2;
}
static FT SUF(psi_momentum)(int kind, FT z) {
  switch (kind) {
    case COFLUX_STABILITY_EDSON: return SUF(psi_edson_momentum)(z);
    case COFLUX_STABILITY_SHEBA_PAULSON: return SUF(psi_sheba_momentum)(z);
    case COFLUX_STABILITY_LARGE_YEAGER: return (z >= (FT)0) ? -(FT)5 * z : SUF(psi_paulson_momentum)(z);
    default: return (FT)0;
  }
}
static FT SUF(psi_scalar)(int kind, FT z) {
  switch (kind) {
    case COFLUX_STABILITY_EDSON: return SUF(psi_edson_scalar)(z);
    case COFLUX_STABILITY_SHEBA_PAULSON: return SUF(psi_sheba_scalar)(z);
    case COFLUX_STABILITY_LARGE_YEAGER: return (z >= (FT)0) ? -(FT)5 * z : SUF(psi_paulson_scalar)(z);
    default: return (FT)0;
  }
}

/* A6: roughness lengths ---------------------------------------------------------------------- */
static FT SUF(air_viscosity)(const coflux_air_viscosity* v, FT T) {
  if (v->kind == COFLUX_VISCOSITY_CONSTANT) return (FT)v->nu;
  FT Tp = T - (FT)273.15;
  return (FT)v->c0 + (FT)v->c1 * Tp + (FT)v->c2 * Tp * Tp + (FT)v->c3 * Tp * Tp * Tp;
}
static FT SUF(momentum_roughness_length)(const coflux_momentum_roughness* r, FT ustar, FT U, FT Ts) {
  if (r->kind == COFLUX_ROUGHNESS_FIXED) return (FT)r->fixed_length;
  FT g = (FT)r->gravitational_acceleration;
  FT lm = (FT)r->maximum_length;
  FT nu = SUF(air_viscosity)(&r->viscosity, Ts);
  FT alpha;
  if (r->wave_formulation == COFLUX_WAVES_WIND_DEPENDENT) {
    alpha = (FT)r->wind_a1 * FMIN(U, (FT)r->wind_umax) + (FT)r->wind_a2;
    alpha = FMAX(alpha, (FT)r->wind_alpha_min);
  } else {
    alpha = (FT)r->gravity_wave_parameter;
  }
  FT lR = (ustar == (FT)0) ? lm : (FT)r->smooth_wall_parameter * nu / ustar;
  return FMIN(alpha * ustar * ustar / g + lR, lm);
}
static FT SUF(scalar_roughness_length)(const coflux_scalar_roughness* r, FT lu, FT ustar, FT Ts) {
  if (r->kind == COFLUX_ROUGHNESS_FIXED) return (FT)r->fixed_length;
  FT nu = SUF(air_viscosity)(&r->viscosity, Ts);
  FT Rstar = lu * ustar / nu;
  FT lq = (Rstar == (FT)0) ? (FT)0 : (FT)r->reynolds_A / POW(Rstar, (FT)r->reynolds_b);
  return FMIN(lq, (FT)r->maximum_length);
}

/* A4: similarity profile χ = ln(h/ℓ) − ψ(h/L) [+ ψ(ℓ/L)] */
static FT SUF(similarity_profile)(int form, int stab, int scalar, FT h, FT l, FT L) {
  FT zeta = h / L;
  FT psi_h = scalar ? SUF(psi_scalar)(stab, zeta) : SUF(psi_momentum)(stab, zeta);
  if (form == COFLUX_PROFILE_COARE_LOGARITHMIC) return LOG(h / l) - psi_h;
  FT theta = l / L;
  FT psi_l = scalar ? SUF(psi_scalar)(stab, theta) : SUF(psi_momentum)(stab, theta);
  return LOG(h / l) - psi_h + psi_l;
}

/* A4: buoyancy scale from the surface thermodynamic state */
static FT SUF(buoyancy_scale)(FT theta_star, FT q_star, FT T_v, FT q_vap, FT eps, FT g) {
  FT delta = eps - (FT)1;
  return g / T_v * (theta_star * ((FT)1 + delta * q_vap) + delta * T_v * q_star);
}

typedef struct { FT ustar, tstar, qstar, Ts, qs; int iterations; } SUF(scales);

typedef struct {
  /* per-cell inputs */
  FT ua, va, Ta, pa, qa, Qs, Ql;
  FT uo, vo, To /* Kelvin */, So;
  /* sea ice (SKIN temperature) */
  FT h_ice, S_ice, albedo;
} SUF(cell_in);

/* A7: Large & Yeager neutral 10 m transfer coefficients */
static FT SUF(ly_cdn)(FT U) {
  if (U >= (FT)33) return (FT)2.34e-3;
  FT U2 = U * U, U6 = U2 * U2 * U2;
  return (FT)1e-3 * ((FT)2.7 / U + (FT)0.142 + U / (FT)13.09 - (FT)3.14807e-10 * U6);
}

/* One fixed-point update of (u★, θ★, q★) — MOST form (A4) */
static void SUF(iterate_similarity)(const coflux_flux_params* P, const coflux_atmosphere_properties* A,
                                    const SUF(thermo_consts)* c, FT du, FT dv, FT dtheta, FT dq,
                                    const SUF(thermo_state)* surf, FT Ts, FT* us, FT* ts, FT* qs_) {
  FT g = (FT)A->gravitational_acceleration;
  FT h = (FT)A->surface_layer_height;
  FT hbl = (FT)A->boundary_layer_height;
  FT kappa = (FT)P->von_karman_constant;
  FT ustar = *us, tstar = *ts, qstar = *qs_;

  FT bstar = SUF(buoyancy_scale)(tstar, qstar, surf->T_v, surf->q_vap, c->eps, g);
  FT Jb = -ustar * bstar;
  FT UG = (FT)P->gustiness_parameter * CBRT(Jb * hbl);
  UG = FMAX(UG, (FT)P->minimum_gustiness);
  FT U = SQRT(du * du + dv * dv + UG * UG);
  if (U == (FT)0) { *us = (FT)0; *ts = (FT)0; *qs_ = (FT)0; return; } /* documented calm-cell guard */

  FT lu = SUF(momentum_roughness_length)(&P->momentum_roughness, ustar, U, Ts);
  FT lq = SUF(scalar_roughness_length)(&P->water_vapor_roughness, lu, ustar, Ts);
  FT lt = SUF(scalar_roughness_length)(&P->temperature_roughness, lu, ustar, Ts);

  FT Lstar = (bstar == (FT)0) ? (FT)HUGE_VAL : ustar * ustar / (kappa * bstar);

  FT prof_u = SUF(similarity_profile)(P->similarity_form, P->stability_functions, 0, h, lu, Lstar);
  FT prof_t = SUF(similarity_profile)(P->similarity_form, P->stability_functions, 1, h, lt, Lstar);
  FT prof_q = SUF(similarity_profile)(P->similarity_form, P->stability_functions, 1, h, lq, Lstar);
  /* documented guard: the profile function ∫φ dz/z is positive by construction in the standard form;
   * the COARE form (no ψ(ℓ/L) term) can turn non-positive in wild transients (tiny u★, ℓ → ℓ_max).
   * A non-positive momentum profile resets the iterate to zero scales (the next pass re-enters
   * through the neutral log law, like a calm cell); a non-positive scalar profile zeroes that scale. */
  if (!(prof_u > (FT)0)) { *us = (FT)0; *ts = (FT)0; *qs_ = (FT)0; return; }
  FT chi_u = kappa / prof_u;
  FT chi_t = (prof_t > (FT)0) ? kappa / prof_t : (FT)0;
  FT chi_q = (prof_q > (FT)0) ? kappa / prof_q : (FT)0;

  *us = chi_u * U;
  *ts = chi_t * dtheta;
  *qs_ = chi_q * dq;
}

/* One update of the coefficient-based Large–Yeager form (A7).  State carried: (u★,θ★,q★) and
 * the neutral drag coefficient is recomputed from the shifted wind each pass. */
static void SUF(ly_scales_from_coeffs)(FT cd, FT ch, FT ce, FT U, FT dtheta, FT dq, FT* us, FT* ts, FT* qs_) {
  FT rcd = SQRT(cd);
  *us = rcd * U;
  *ts = ch / rcd * dtheta;
  *qs_ = ce / rcd * dq;
}
static void SUF(iterate_large_yeager)(const coflux_flux_params* P, const coflux_atmosphere_properties* A,
                                      const SUF(thermo_consts)* c, FT U, FT dtheta, FT dq,
                                      const SUF(thermo_state)* surf, FT* rcdn_io, FT* us, FT* ts, FT* qs_) {
  FT g = (FT)A->gravitational_acceleration;
  FT h = (FT)A->surface_layer_height;
  FT kappa = (FT)P->von_karman_constant;
  FT ustar = *us, tstar = *ts, qstar = *qs_;
  FT bstar = SUF(buoyancy_scale)(tstar, qstar, surf->T_v, surf->q_vap, c->eps, g);
  FT zeta = kappa * bstar * h / (ustar * ustar);
  zeta = FMAX((FT)-10, FMIN((FT)10, zeta));
  FT psim = SUF(psi_momentum)(COFLUX_STABILITY_LARGE_YEAGER, zeta);
  FT psih = SUF(psi_scalar)(COFLUX_STABILITY_LARGE_YEAGER, zeta);
  FT lnh = LOG(h / (FT)10);
  /* shift the wind to 10 m, neutral, with the previous neutral drag coefficient */
  FT U10N = U / ((FT)1 + *rcdn_io / kappa * (lnh - psim));
  U10N = FMAX(U10N, (FT)P->ly_minimum_wind);
  FT cdn = SUF(ly_cdn)(U10N);
  FT rcdn = SQRT(cdn);
  FT cen = (FT)34.6e-3 * rcdn;
  FT chn = ((zeta > (FT)0) ? (FT)18e-3 : (FT)32.7e-3) * rcdn;
  FT xm = (FT)1 + rcdn / kappa * (lnh - psim);
  FT cd = cdn / (xm * xm);
  FT ch = chn / ((FT)1 + chn / (kappa * rcdn) * (lnh - psih)) * SQRT(cd / cdn);
  FT ce = cen / ((FT)1 + cen / (kappa * rcdn) * (lnh - psih)) * SQRT(cd / cdn);
  *rcdn_io = rcdn;
  SUF(ly_scales_from_coeffs)(cd, ch, ce, U, dtheta, dq, us, ts, qs_);
}

/* SKIN interface temperature (row a7): conductive flux balance through the ice slab */
static FT SUF(skin_temperature)(const coflux_flux_params* P, const coflux_ice_ocean_params* I,
                                const coflux_radiation_properties* R, const coflux_ocean_properties* O,
                                FT Ts_prev, FT ustar, FT tstar, FT qstar, FT rho_a, FT cp_a, FT Ls,
                                FT Qs, FT Ql, FT albedo, FT h_ice, FT S_ice) {
  FT sigma = (FT)R->stefan_boltzmann_constant;
  FT emis = (FT)R->sea_ice_emissivity;
  FT k = (FT)I->ice_conductivity;
  FT hc = (FT)I->ice_consolidation_thickness;
  FT Toff = (O->temperature_units == COFLUX_TEMPERATURE_CELSIUS) ? (FT)273.15 : (FT)0;
  FT Tb = (FT)I->liquidus_freshwater_melting_temperature - (FT)I->liquidus_slope * S_ice + Toff; /* bottom at melting */
  FT Tm = (FT)I->liquidus_freshwater_melting_temperature + Toff;
  FT Qu = emis * sigma * Ts_prev * Ts_prev * Ts_prev * Ts_prev;
  FT Qd = -((FT)1 - albedo) * Qs - emis * Ql;
  FT Qc = -rho_a * cp_a * ustar * tstar;
  FT Qv = -rho_a * Ls * ustar * qstar;
  FT Qa = Qv + Qu + Qc + Qd;
  FT Tstar = Tb - Qa * h_ice / k;
  if (P->skin_temperature_update == COFLUX_SKIN_LINEARIZED_LONGWAVE) {
    /* emitted long wave implicit: Q_u ≈ σ ε T_s⁻³ · T_s⁺ (include/coflux.h: coflux_skin_temperature_update) */
    FT alpha = sigma * emis * Ts_prev * Ts_prev * Ts_prev / k;
    Tstar = (Tb - (Qd + Qc + Qv) * h_ice / k) / ((FT)1 + alpha * h_ice);
  }
  if (Tstar != Tstar) Tstar = Ts_prev;
  Tstar = FMAX((FT)0, Tstar);
  FT Tnew = (h_ice >= hc) ? Tstar : Tb;
  FT dT = Tnew - Ts_prev;
  FT maxdT = (FT)P->skin_max_delta_T;
  FT adT = FMIN(maxdT, FABS(dT));
  FT sgn = (dT > (FT)0) ? (FT)1 : ((dT < (FT)0) ? (FT)-1 : (FT)0);
  Tnew = Ts_prev + adT * sgn;
  return FMIN(Tnew, Tm);
}

/* A3/A4: full per-cell solve.  surface_kind: 0 ocean (liquid, Raoult), 1 sea ice (ice phase). */
static SUF(scales) SUF(solve_cell)(const coflux_flux_params* P, const coflux_config* cfg,
                                   const SUF(thermo_consts)* c, const SUF(cell_in)* in, int surface_kind,
                                   SUF(thermo_state)* atm_out, FT* du_out, FT* dv_out) {
  const coflux_atmosphere_properties* A = &cfg->atmosphere;
  FT g = (FT)A->gravitational_acceleration;
  FT h = (FT)A->surface_layer_height;

  SUF(thermo_state) atm = SUF(phase_equil_pTq)(c, in->pa, in->Ta, in->qa);
  *atm_out = atm;
  FT du, dv;
  if (P->velocity_formulation == COFLUX_VELOCITY_RELATIVE) { du = in->ua - in->uo; dv = in->va - in->vo; }
  else { du = in->ua; dv = in->va; }
  *du_out = du; *dv_out = dv;

  FT x = (surface_kind == 0) ? SUF(water_mole_fraction)(&cfg->ocean, in->So) : (FT)1;
  FT Ts = in->To;
  FT qs = SUF(surface_specific_humidity)(c, atm.rho, Ts, x, surface_kind);
  FT theta_a = in->Ta + g * h / atm.cp_m;

  SUF(scales) cur, prev;
  cur.ustar = cur.tstar = cur.qstar = (FT)P->initial_scale;
  cur.Ts = Ts; cur.qs = qs; cur.iterations = 0;
  prev = cur;
  int it = 0;
  int maxit = P->max_iterations;
  FT tol = (FT)P->tolerance;

  /* Large–Yeager: first guess from neutral coefficients at the floored wind */
  FT U_ly = (FT)0, rcdn_ly = (FT)0;
  if (P->formulation == COFLUX_FLUXES_COEFFICIENT_LARGE_YEAGER) {
    U_ly = FMAX(SQRT(du * du + dv * dv), (FT)P->ly_minimum_wind);
    FT dtheta0 = theta_a - Ts, dq0 = atm.q_vap - qs;
    FT cdn = SUF(ly_cdn)(U_ly), rcdn = SQRT(cdn);
    FT chn = ((dtheta0 > (FT)0) ? (FT)18e-3 : (FT)32.7e-3) * rcdn;
    FT cen = (FT)34.6e-3 * rcdn;
    rcdn_ly = rcdn;
    SUF(ly_scales_from_coeffs)(cdn, chn, cen, U_ly, dtheta0, dq0, &cur.ustar, &cur.tstar, &cur.qstar);
    prev = cur;
  }

  for (;;) {
    int go;
    if (P->stop_kind == COFLUX_STOP_FIXED_ITERATIONS) {
      go = it < maxit;
    } else {
      FT drift = FABS(cur.ustar - prev.ustar) + FABS(cur.tstar - prev.tstar) + FABS(cur.qstar - prev.qstar);
      int converged = drift < tol;
      int reached = it >= maxit;
      go = (!(converged || reached)) || (it == 0);
    }
    if (!go) break;
    prev = cur;

    /* interface temperature */
    if (P->interface_temperature == COFLUX_TEMPERATURE_SKIN) {
      FT Ls = SUF(latent_heat_sublimation)(c, atm.T);
      Ts = SUF(skin_temperature)(P, &cfg->ice_ocean, &cfg->radiation, &cfg->ocean, prev.Ts, prev.ustar, prev.tstar,
                                 prev.qstar, atm.rho, atm.cp_m, Ls, in->Qs, in->Ql, in->albedo, in->h_ice, in->S_ice);
    } else {
      Ts = in->To;
    }
    qs = SUF(surface_specific_humidity)(c, atm.rho, Ts, x, surface_kind);
    FT dq = atm.q_vap - qs;
    FT dtheta = theta_a - Ts;
    SUF(thermo_state) surf = SUF(phase_equil_pTq)(c, atm.p, Ts, qs);

    FT us = prev.ustar, ts = prev.tstar, qq = prev.qstar;
    if (P->formulation == COFLUX_FLUXES_COEFFICIENT_LARGE_YEAGER)
      SUF(iterate_large_yeager)(P, A, c, U_ly, dtheta, dq, &surf, &rcdn_ly, &us, &ts, &qq);
    else
      SUF(iterate_similarity)(P, A, c, du, dv, dtheta, dq, &surf, Ts, &us, &ts, &qq);
    cur.ustar = us; cur.tstar = ts; cur.qstar = qq; cur.Ts = Ts; cur.qs = qs;
    ++it;
  }
  cur.iterations = it;
  return cur;
}

/* ---------------------------------------------------------------------------------------------
 * Array helpers
 * ------------------------------------------------------------------------------------------- */
static inline int64_t SUF(idx)(const coflux_array* a, int i, int j, int k, int n) {
  return (int64_t)(i + a->off_i) * a->stride_i + (int64_t)(j + a->off_j) * a->stride_j +
         (int64_t)(k + a->off_k) * a->stride_k + (int64_t)n * a->stride_n;
}
static inline FT SUF(ld)(const coflux_array* a, int i, int j, int k, int n) {
  return ((const FT*)a->ptr)[SUF(idx)(a, i, j, k, n)];
}
static inline void SUF(st)(const coflux_array* a, int i, int j, int k, FT v) {
  if (a->ptr) ((FT*)a->ptr)[SUF(idx)(a, i, j, k, 0)] = v;
}
static inline int SUF(active)(const coflux_array* mask, int i, int j) {
  if (!mask->ptr) return 1;
  return ((const uint8_t*)mask->ptr)[SUF(idx)(mask, i, j, 0, 0)] != 0;
}

/* device ring buffer of a series (coflux_forcing_window): logical level n lives in time slot (start + n) mod capacity */
static int SUF(ring_slot)(int n, int start, int capacity) { return capacity > 0 ? (start + n) % capacity : n; }

/* A8: bilinear (space) × linear (time) interpolation of one series */
static FT SUF(interp_series)(const coflux_array* a, FT fi, FT fj, int n1, int n2, FT nfrac) {
  int i0 = (int)TRUNC(fi), j0 = (int)TRUNC(fj);
  int si = (fi > (FT)0) - (fi < (FT)0), sj = (fj > (FT)0) - (fj < (FT)0);
  int i1 = i0 + si, j1 = j0 + sj;
  FT xi = fi - FLOOR(fi), eta = fj - FLOOR(fj);
  FT one = (FT)1;
  FT p1 = (one - xi) * (one - eta) * SUF(ld)(a, i0, j0, 0, n1) + (one - xi) * eta * SUF(ld)(a, i0, j1, 0, n1) +
          xi * (one - eta) * SUF(ld)(a, i1, j0, 0, n1) + xi * eta * SUF(ld)(a, i1, j1, 0, n1);
  FT p2 = (one - xi) * (one - eta) * SUF(ld)(a, i0, j0, 0, n2) + (one - xi) * eta * SUF(ld)(a, i0, j1, 0, n2) +
          xi * (one - eta) * SUF(ld)(a, i1, j0, 0, n2) + xi * eta * SUF(ld)(a, i1, j1, 0, n2);
  return p2 * nfrac + p1 * (one - nfrac);
}

int SUF(oracle_interpolate_atmosphere)(const coflux_config* cfg, const coflux_atmos_series* in, double time,
                                       coflux_exchange_state* out) {
  int n1, n2; double frac;
  int rc = oracle_time_indices(in->times, in->Nt, in->time_indexing, in->cycle_period, time, &n1, &n2, &frac);
  if (rc) return rc;
  const int Nx = cfg->grid.Nx, Ny = cfg->grid.Ny, r = cfg->grid.ring;
  FT nf = (FT)frac;
  n1 = SUF(ring_slot)(n1, in->ring_start, in->ring_capacity);
  n2 = SUF(ring_slot)(n2, in->ring_start, in->ring_capacity);
#pragma omp parallel for schedule(static)
  for (int j = -r; j < Ny + r; ++j)
    for (int i = -r; i < Nx + r; ++i) {
      FT fi = SUF(ld)(&in->fi, i, j, 0, 0), fj = SUF(ld)(&in->fj, i, j, 0, 0);
      FT u = SUF(interp_series)(&in->u, fi, fj, n1, n2, nf);
      FT v = SUF(interp_series)(&in->v, fi, fj, n1, n2, nf);
      FT T = SUF(interp_series)(&in->T, fi, fj, n1, n2, nf);
      FT q = SUF(interp_series)(&in->q, fi, fj, n1, n2, nf);
      FT p = SUF(interp_series)(&in->p, fi, fj, n1, n2, nf);
      FT Qs = SUF(interp_series)(&in->Qs, fi, fj, n1, n2, nf);
      FT Ql = SUF(interp_series)(&in->Ql, fi, fj, n1, n2, nf);
      FT Mp = (FT)0;
      if (in->rain.ptr) Mp += SUF(interp_series)(&in->rain, fi, fj, n1, n2, nf);
      if (in->snow.ptr) Mp += SUF(interp_series)(&in->snow, fi, fj, n1, n2, nf);
      if (in->cos_theta.ptr && in->sin_theta.ptr) {
        FT cs = SUF(ld)(&in->cos_theta, i, j, 0, 0), sn = SUF(ld)(&in->sin_theta, i, j, 0, 0);
        FT ur = u * cs + v * sn, vr = -u * sn + v * cs;   /* extrinsic → intrinsic */
        u = ur; v = vr;
      }
      SUF(st)(&out->u, i, j, 0, u); SUF(st)(&out->v, i, j, 0, v); SUF(st)(&out->T, i, j, 0, T);
      SUF(st)(&out->p, i, j, 0, p); SUF(st)(&out->q, i, j, 0, q); SUF(st)(&out->Qs, i, j, 0, Qs);
      SUF(st)(&out->Ql, i, j, 0, Ql); SUF(st)(&out->Mp, i, j, 0, Mp);
    }
  return 0;
}

/* Land freshwater (JRA55PrescribedLand, /root/reference/src/OMIPConfigurations/atmosphere.jl:46; variables friver, licalvf:
 * jra55_data_staging.jl:8): river runoff and iceberg calving, interpolated like the atmosphere on their own source grid
 * and time axis and ADDED to the exchange freshwater flux:  Mp ← Mp + (M_rivers + M_icebergs). */
int SUF(oracle_interpolate_land)(const coflux_config* cfg, const coflux_land_series* in, double time, coflux_exchange_state* x) {
  int n1, n2; double frac;
  int rc = oracle_time_indices(in->times, in->Nt, in->time_indexing, in->cycle_period, time, &n1, &n2, &frac);
  if (rc) return rc;
  const int Nx = cfg->grid.Nx, Ny = cfg->grid.Ny, r = cfg->grid.ring;
  FT nf = (FT)frac;
  n1 = SUF(ring_slot)(n1, in->ring_start, in->ring_capacity);
  n2 = SUF(ring_slot)(n2, in->ring_start, in->ring_capacity);
#pragma omp parallel for schedule(static)
  for (int j = -r; j < Ny + r; ++j)
    for (int i = -r; i < Nx + r; ++i) {
      FT fi = SUF(ld)(&in->fi, i, j, 0, 0), fj = SUF(ld)(&in->fj, i, j, 0, 0);
      FT Mr = in->rivers.ptr ? SUF(interp_series)(&in->rivers, fi, fj, n1, n2, nf) : (FT)0;
      FT Mi = in->icebergs.ptr ? SUF(interp_series)(&in->icebergs, fi, fj, n1, n2, nf) : (FT)0;
      SUF(st)(&x->Mp, i, j, 0, SUF(ld)(&x->Mp, i, j, 0, 0) + (Mr + Mi));
    }
  return 0;
}

/* CCSM3 sea-ice albedo (Briegleb et al. 2004; the "ccsm3" shortwave option of CICE), include/coflux.h coflux_ccsm3_albedo;
 * reference call site: SeaIceAlbedo(hi, hs, Ts), /root/reference/src/OMIPConfigurations/atmosphere.jl:31-44 */
static FT SUF(ccsm3_albedo)(const coflux_ccsm3_albedo* a, FT hi, FT hs, FT TsK) {
  FT fh = FMIN(ATAN((FT)4 * hi) / ATAN((FT)4 * (FT)a->thickness_scale), (FT)1);
  FT fT = FMIN(FMAX((FT)1 - ((FT)a->melting_temperature - TsK) / (FT)a->melt_temperature_range, (FT)0), (FT)1);
  FT aiv = (FT)a->ice_visible * fh + (FT)a->ocean_albedo * ((FT)1 - fh) - (FT)a->ice_melt_change * fT;
  FT ain = (FT)a->ice_near_infrared * fh + (FT)a->ocean_albedo * ((FT)1 - fh) - (FT)a->ice_melt_change * fT;
  FT asv = (FT)a->snow_visible - (FT)a->snow_visible_melt_change * fT;
  FT asn = (FT)a->snow_near_infrared - (FT)a->snow_near_infrared_melt_change * fT;
  FT fs = hs / (hs + (FT)a->snow_patchiness);
  FT av = ((FT)1 - fs) * aiv + fs * asv;
  FT an = ((FT)1 - fs) * ain + fs * asn;
  return (FT)a->visible_fraction * av + ((FT)1 - (FT)a->visible_fraction) * an;
}
static FT SUF(sea_ice_albedo)(const coflux_config* cfg, const coflux_sea_ice_state* ice, int i, int j, FT TsK) {
  if (cfg->radiation.sea_ice_albedo_kind == COFLUX_SEA_ICE_ALBEDO_CCSM3) {
    FT hs = ice->snow_thickness.ptr ? SUF(ld)(&ice->snow_thickness, i, j, 0, 0) : (FT)0;
    return SUF(ccsm3_albedo)(&cfg->radiation.ccsm3, SUF(ld)(&ice->thickness, i, j, 0, 0), hs, TsK);
  }
  return ice->albedo.ptr ? SUF(ld)(&ice->albedo, i, j, 0, 0) : (FT)cfg->radiation.sea_ice_albedo;
}

static FT SUF(to_kelvin)(const coflux_ocean_properties* o, FT T) {
  return (o->temperature_units == COFLUX_TEMPERATURE_CELSIUS) ? T + (FT)273.15 : T;
}
static FT SUF(from_kelvin)(const coflux_ocean_properties* o, FT T) {
  return (o->temperature_units == COFLUX_TEMPERATURE_CELSIUS) ? T - (FT)273.15 : T;
}

/* A3: atmosphere–ocean interface state + turbulent fluxes */
int SUF(oracle_atmosphere_ocean_fluxes)(const coflux_config* cfg, const coflux_exchange_state* atmos,
                                        const coflux_ocean_surface* ocean, coflux_interface_fluxes* out) {
  const int Nx = cfg->grid.Nx, Ny = cfg->grid.Ny, kN = cfg->grid.Nz - 1, r = cfg->grid.ring;
  const coflux_flux_params* P = &cfg->atmosphere_ocean;
  SUF(thermo_consts) c = SUF(make_thermo)(&cfg->atmosphere.thermodynamics);
#pragma omp parallel for schedule(dynamic, 4)
  for (int j = -r; j < Ny + r; ++j)
    for (int i = -r; i < Nx + r; ++i) {
      SUF(cell_in) in;
      in.ua = SUF(ld)(&atmos->u, i, j, 0, 0); in.va = SUF(ld)(&atmos->v, i, j, 0, 0);
      in.Ta = SUF(ld)(&atmos->T, i, j, 0, 0); in.pa = SUF(ld)(&atmos->p, i, j, 0, 0);
      in.qa = SUF(ld)(&atmos->q, i, j, 0, 0);
      in.Qs = SUF(ld)(&atmos->Qs, i, j, 0, 0); in.Ql = SUF(ld)(&atmos->Ql, i, j, 0, 0);
      in.uo = (SUF(ld)(&ocean->u, i, j, kN, 0) + SUF(ld)(&ocean->u, i + 1, j, kN, 0)) * (FT)0.5;
      in.vo = (SUF(ld)(&ocean->v, i, j, kN, 0) + SUF(ld)(&ocean->v, i, j + 1, kN, 0)) * (FT)0.5;
      FT To_units = SUF(ld)(&ocean->T, i, j, kN, 0);
      in.To = SUF(to_kelvin)(&cfg->ocean, To_units);
      in.So = SUF(ld)(&ocean->S, i, j, kN, 0);
      in.h_ice = in.S_ice = in.albedo = (FT)0;
      FT Qv = 0, Qc = 0, Fv = 0, rtx = 0, rty = 0, Tsout = To_units, us = 0, ts = 0, qs = 0; int its = 0;
      if (SUF(active)(&ocean->mask, i, j)) {
        SUF(thermo_state) atm; FT du, dv;
        SUF(scales) s = SUF(solve_cell)(P, cfg, &c, &in, 0, &atm, &du, &dv);
        FT dU = SQRT(du * du + dv * dv);
        FT taux = (dU == (FT)0) ? dU : -s.ustar * s.ustar * du / dU;
        FT tauy = (dU == (FT)0) ? dU : -s.ustar * s.ustar * dv / dU;
        FT Lv = SUF(latent_heat_vapor)(&c, atm.T);
        Qv = -atm.rho * s.ustar * s.qstar * Lv;
        Qc = -atm.rho * atm.cp_m * s.ustar * s.tstar;
        Fv = -atm.rho * s.ustar * s.qstar;
        rtx = atm.rho * taux; rty = atm.rho * tauy;
        Tsout = SUF(from_kelvin)(&cfg->ocean, s.Ts);
        us = s.ustar; ts = s.tstar; qs = s.qstar; its = s.iterations;
      }
      SUF(st)(&out->latent_heat, i, j, 0, Qv); SUF(st)(&out->sensible_heat, i, j, 0, Qc);
      SUF(st)(&out->water_vapor, i, j, 0, Fv); SUF(st)(&out->x_momentum, i, j, 0, rtx);
      SUF(st)(&out->y_momentum, i, j, 0, rty); SUF(st)(&out->interface_temperature, i, j, 0, Tsout);
      SUF(st)(&out->friction_velocity, i, j, 0, us); SUF(st)(&out->temperature_scale, i, j, 0, ts);
      SUF(st)(&out->humidity_scale, i, j, 0, qs);
      if (out->iterations.ptr) ((int32_t*)out->iterations.ptr)[SUF(idx)(&out->iterations, i, j, 0, 0)] = its;
    }
  return 0;
}

/* a7: atmosphere–sea-ice interface (skin temperature inside the iteration).  Cells with no ice
 * (ℵ == 0 or h == 0) produce zero fluxes and keep T_top. */
int SUF(oracle_atmosphere_sea_ice_fluxes)(const coflux_config* cfg, const coflux_exchange_state* atmos,
                                          const coflux_ocean_surface* ocean, coflux_sea_ice_state* ice,
                                          coflux_interface_fluxes* out) {
  const int Nx = cfg->grid.Nx, Ny = cfg->grid.Ny, r = cfg->grid.ring;
  const coflux_flux_params* P = &cfg->atmosphere_sea_ice;
  SUF(thermo_consts) c = SUF(make_thermo)(&cfg->atmosphere.thermodynamics);
#pragma omp parallel for schedule(dynamic, 4)
  for (int j = -r; j < Ny + r; ++j)
    for (int i = -r; i < Nx + r; ++i) {
      SUF(cell_in) in;
      in.ua = SUF(ld)(&atmos->u, i, j, 0, 0); in.va = SUF(ld)(&atmos->v, i, j, 0, 0);
      in.Ta = SUF(ld)(&atmos->T, i, j, 0, 0); in.pa = SUF(ld)(&atmos->p, i, j, 0, 0);
      in.qa = SUF(ld)(&atmos->q, i, j, 0, 0);
      in.Qs = SUF(ld)(&atmos->Qs, i, j, 0, 0); in.Ql = SUF(ld)(&atmos->Ql, i, j, 0, 0);
      in.uo = (SUF(ld)(&ice->u, i, j, 0, 0) + SUF(ld)(&ice->u, i + 1, j, 0, 0)) * (FT)0.5;
      in.vo = (SUF(ld)(&ice->v, i, j, 0, 0) + SUF(ld)(&ice->v, i, j + 1, 0, 0)) * (FT)0.5;
      FT Ttop_units = SUF(ld)(&ice->top_temperature, i, j, 0, 0);
      in.To = SUF(to_kelvin)(&cfg->ocean, Ttop_units);
      in.So = (FT)0;
      in.h_ice = SUF(ld)(&ice->thickness, i, j, 0, 0);
      in.S_ice = SUF(ld)(&ice->salinity, i, j, 0, 0);
      in.albedo = SUF(sea_ice_albedo)(cfg, ice, i, j, in.To);
      FT conc = SUF(ld)(&ice->concentration, i, j, 0, 0);
      FT Qv = 0, Qc = 0, Fv = 0, rtx = 0, rty = 0, Tsout = Ttop_units, us = 0, ts = 0, qs = 0; int its = 0;
      if (SUF(active)(&ocean->mask, i, j) && conc > (FT)0 && in.h_ice > (FT)0) {
        SUF(thermo_state) atm; FT du, dv;
        SUF(scales) s = SUF(solve_cell)(P, cfg, &c, &in, 1, &atm, &du, &dv);
        FT dU = SQRT(du * du + dv * dv);
        FT taux = (dU == (FT)0) ? dU : -s.ustar * s.ustar * du / dU;
        FT tauy = (dU == (FT)0) ? dU : -s.ustar * s.ustar * dv / dU;
        FT Ls = SUF(latent_heat_sublimation)(&c, atm.T);
        Qv = -atm.rho * s.ustar * s.qstar * Ls;
        Qc = -atm.rho * atm.cp_m * s.ustar * s.tstar;
        Fv = -atm.rho * s.ustar * s.qstar;
        rtx = atm.rho * taux; rty = atm.rho * tauy;
        Tsout = SUF(from_kelvin)(&cfg->ocean, s.Ts);
        us = s.ustar; ts = s.tstar; qs = s.qstar; its = s.iterations;
      }
      SUF(st)(&out->latent_heat, i, j, 0, Qv); SUF(st)(&out->sensible_heat, i, j, 0, Qc);
      SUF(st)(&out->water_vapor, i, j, 0, Fv); SUF(st)(&out->x_momentum, i, j, 0, rtx);
      SUF(st)(&out->y_momentum, i, j, 0, rty); SUF(st)(&out->interface_temperature, i, j, 0, Tsout);
      SUF(st)(&ice->top_temperature, i, j, 0, Tsout);
      SUF(st)(&out->friction_velocity, i, j, 0, us); SUF(st)(&out->temperature_scale, i, j, 0, ts);
      SUF(st)(&out->humidity_scale, i, j, 0, qs);
      if (out->iterations.ptr) ((int32_t*)out->iterations.ptr)[SUF(idx)(&out->iterations, i, j, 0, 0)] = its;
    }
  return 0;
}

/* A10: sea-ice–ocean fluxes: frazil sweep, interface heat, salt, quadratic stress */
static FT SUF(dz_at)(const coflux_array* dz, int i, int j, int k) {
  if (dz->stride_i == 0 && dz->stride_j == 0) return ((const FT*)dz->ptr)[(int64_t)(k + dz->off_k) * dz->stride_k];
  return SUF(ld)(dz, i, j, k, 0);
}
int SUF(oracle_sea_ice_ocean_fluxes)(const coflux_config* cfg, coflux_ocean_columns* oc, coflux_sea_ice_state* ice,
                                     double dt_, coflux_ice_ocean_fluxes* out) {
  const int Nx = cfg->grid.Nx, Ny = cfg->grid.Ny, Nz = cfg->grid.Nz;
  const coflux_ice_ocean_params* I = &cfg->ice_ocean;
  FT rho0 = (FT)cfg->ocean.reference_density, c0 = (FT)cfg->ocean.heat_capacity;
  FT T0 = (FT)I->liquidus_freshwater_melting_temperature, m = (FT)I->liquidus_slope;
  FT dt = (FT)dt_;
  /* stresses first (they feed the momentum-based friction velocity); interior only, needs i-1 / j-1 halos */
#pragma omp parallel for schedule(static)
  for (int j = 0; j < Ny; ++j)
    for (int i = 0; i < Nx; ++i) {
      FT Cd = (FT)I->ice_ocean_drag_coefficient;
      /* (Face, Center) */
      FT dux = SUF(ld)(&ice->u, i, j, 0, 0) - SUF(ld)(&oc->u, i, j, Nz - 1, 0);
      FT viF = (FT)0.25 * (SUF(ld)(&ice->v, i - 1, j, 0, 0) + SUF(ld)(&ice->v, i, j, 0, 0) +
                           SUF(ld)(&ice->v, i - 1, j + 1, 0, 0) + SUF(ld)(&ice->v, i, j + 1, 0, 0));
      FT voF = (FT)0.25 * (SUF(ld)(&oc->v, i - 1, j, Nz - 1, 0) + SUF(ld)(&oc->v, i, j, Nz - 1, 0) +
                           SUF(ld)(&oc->v, i - 1, j + 1, Nz - 1, 0) + SUF(ld)(&oc->v, i, j + 1, Nz - 1, 0));
      FT dvx = viF - voF;
      FT taux = rho0 * Cd * SQRT(dux * dux + dvx * dvx) * dux;
      /* (Center, Face) */
      FT dvy = SUF(ld)(&ice->v, i, j, 0, 0) - SUF(ld)(&oc->v, i, j, Nz - 1, 0);
      FT uiF = (FT)0.25 * (SUF(ld)(&ice->u, i, j - 1, 0, 0) + SUF(ld)(&ice->u, i, j, 0, 0) +
                           SUF(ld)(&ice->u, i + 1, j - 1, 0, 0) + SUF(ld)(&ice->u, i + 1, j, 0, 0));
      FT uoF = (FT)0.25 * (SUF(ld)(&oc->u, i, j - 1, Nz - 1, 0) + SUF(ld)(&oc->u, i, j, Nz - 1, 0) +
                           SUF(ld)(&oc->u, i + 1, j - 1, Nz - 1, 0) + SUF(ld)(&oc->u, i + 1, j, Nz - 1, 0));
      FT duy = uiF - uoF;
      FT tauy = rho0 * Cd * SQRT(duy * duy + dvy * dvy) * dvy;
      SUF(st)(&out->x_momentum, i, j, 0, taux);
      SUF(st)(&out->y_momentum, i, j, 0, tauy);
    }
#pragma omp parallel for schedule(static)
  for (int j = 0; j < Ny; ++j)
    for (int i = 0; i < Nx; ++i) {
      FT dQ_frazil = (FT)0;
      for (int k = Nz - 1; k >= 0; --k) {
        FT dz = SUF(dz_at)(&oc->dz, i, j, k);
        FT Tk = SUF(ld)(&oc->T, i, j, k, 0), Sk = SUF(ld)(&oc->S, i, j, k, 0);
        FT Tm = T0 - m * Sk;
        int freezing = Tk < Tm;
        FT dE = rho0 * c0 * (Tm - Tk);
        if (freezing) {
          ((FT*)oc->T.ptr)[SUF(idx)(&oc->T, i, j, k, 0)] = Tm;
          dQ_frazil -= dE * dz / dt;
        }
      }
      FT TN = SUF(ld)(&oc->T, i, j, Nz - 1, 0), SN = SUF(ld)(&oc->S, i, j, Nz - 1, 0);
      FT conc = SUF(ld)(&ice->concentration, i, j, 0, 0);
      FT Si = SUF(ld)(&ice->salinity, i, j, 0, 0);
      FT Tm = T0 - m * SN;
      FT Qio;
      if (I->heat_flux == COFLUX_ICE_OCEAN_THREE_EQUATION) {
        FT ustar;
        if (I->friction_velocity == COFLUX_FRICTION_VELOCITY_MOMENTUM_BASED) {
          /* |τ| at the cell centre from the face stresses just computed */
          FT tx = (FT)0.5 * (SUF(ld)(&out->x_momentum, i, j, 0, 0) + ((i + 1 < Nx) ? SUF(ld)(&out->x_momentum, i + 1, j, 0, 0) : SUF(ld)(&out->x_momentum, i, j, 0, 0)));
          FT ty = (FT)0.5 * (SUF(ld)(&out->y_momentum, i, j, 0, 0) + ((j + 1 < Ny) ? SUF(ld)(&out->y_momentum, i, j + 1, 0, 0) : SUF(ld)(&out->y_momentum, i, j, 0, 0)));
          ustar = SQRT(SQRT(tx * tx + ty * ty) / rho0);
          ustar = FMAX(ustar, (FT)I->minimum_friction_velocity);
        } else {
          ustar = (FT)I->constant_friction_velocity;
        }
        FT gT = (FT)I->heat_transfer_coefficient * ustar, gS = (FT)I->salt_transfer_coefficient * ustar;
        FT A = rho0 * c0 * gT / ((FT)I->ice_density * (FT)I->ice_latent_heat);
        /* A m S_b² + [A (T_o − T₀) − A m S_i + γ_S] S_b − [A (T_o − T₀) S_i + γ_S S_o] = 0 */
        FT qa = A * m;
        FT qb = A * (TN - T0) - A * m * Si + gS;
        FT qc = -(A * (TN - T0) * Si + gS * SN);
        FT disc = qb * qb - (FT)4 * qa * qc;
        FT Sb = (-qb + SQRT(FMAX(disc, (FT)0))) / ((FT)2 * qa);
        FT Tb = T0 - m * Sb;
        Qio = rho0 * c0 * gT * (TN - Tb) * conc;
      } else {
        FT dE = rho0 * c0 * (Tm - TN);
        Qio = -dE * (FT)I->characteristic_melting_speed * conc;
      }
      FT h = SUF(ld)(&ice->thickness, i, j, 0, 0), hm = SUF(ld)(&ice->previous_thickness, i, j, 0, 0);
      FT Js = (h - hm) / dt * (Si - SN);
      SUF(st)(&out->frazil_heat, i, j, 0, dQ_frazil);
      SUF(st)(&out->interface_heat, i, j, 0, Qio);
      SUF(st)(&out->salt, i, j, 0, Js);
      SUF(st)(&ice->previous_thickness, i, j, 0, h);
    }
  return 0;
}

/* A9: net ocean flux assembly */
int SUF(oracle_assemble_net_ocean_fluxes)(const coflux_config* cfg, const coflux_exchange_state* atmos,
                                          const coflux_ocean_surface* ocean, const coflux_interface_fluxes* ao,
                                          const coflux_sea_ice_state* ice, const coflux_ice_ocean_fluxes* io,
                                          coflux_net_ocean_fluxes* out) {
  const int Nx = cfg->grid.Nx, Ny = cfg->grid.Ny, kN = cfg->grid.Nz - 1, r = cfg->grid.ring;
  const coflux_radiation_properties* R = &cfg->radiation;
  FT sigma = (FT)R->stefan_boltzmann_constant, alpha = (FT)R->ocean_albedo, emis = (FT)R->ocean_emissivity;
  FT rho0inv = (FT)1 / (FT)cfg->ocean.reference_density, c0 = (FT)cfg->ocean.heat_capacity;
  FT rhofinv = (FT)1 / (FT)cfg->ocean.freshwater_density;
  FT Smin = (FT)cfg->ocean.minimum_salinity;
  const int have_ice = ice && ice->concentration.ptr;
  const int px = (r == 0 && cfg->grid.periodic_x);
#pragma omp parallel for schedule(static)
  for (int j = 0; j < Ny; ++j)
    for (int i = 0; i < Nx; ++i) {
      int act = SUF(active)(&ocean->mask, i, j);
      FT conc = have_ice ? SUF(ld)(&ice->concentration, i, j, 0, 0) : (FT)0;
      FT So = SUF(ld)(&ocean->S, i, j, kN, 0);
      FT Ts = SUF(to_kelvin)(&cfg->ocean, SUF(ld)(&ao->interface_temperature, i, j, 0, 0));
      FT Qs = SUF(ld)(&atmos->Qs, i, j, 0, 0), Ql = SUF(ld)(&atmos->Ql, i, j, 0, 0), Mp = SUF(ld)(&atmos->Mp, i, j, 0, 0);
      FT Qc = SUF(ld)(&ao->sensible_heat, i, j, 0, 0), Qv = SUF(ld)(&ao->latent_heat, i, j, 0, 0);
      FT Mv = SUF(ld)(&ao->water_vapor, i, j, 0, 0);
      FT Qu = emis * sigma * Ts * Ts * Ts * Ts;
      FT Qal = -emis * Ql;
      FT Qts = -((FT)1 - alpha) * Qs;
      FT Qss = R->shortwave_penetrates ? (FT)0 : Qts;
      FT SQ = Qu + Qc + Qv + Qal + Qss;
      FT SF = -Mp * rhofinv;
      SF += Mv * rhofinv;
      FT JTao = SQ * rho0inv / c0;
      FT JSao = -So * SF;
      if (So < Smin && JSao > (FT)0) JSao = (FT)0;
      FT Qio = (io && io->interface_heat.ptr) ? SUF(ld)(&io->interface_heat, i, j, 0, 0) : (FT)0;
      FT JSio = (io && io->salt.ptr) ? SUF(ld)(&io->salt, i, j, 0, 0) * conc : (FT)0;
      FT JT = ((FT)1 - conc) * JTao + Qio * rho0inv / c0;
      FT JS = ((FT)1 - conc) * JSao + JSio;
      FT J0 = ((FT)1 - conc) * Qts * rho0inv / c0;
      /* stresses: centre → face */
      int iw = (px && i == 0) ? Nx - 1 : i - 1;
      FT rtx_c = SUF(ld)(&ao->x_momentum, i, j, 0, 0), rtx_w = SUF(ld)(&ao->x_momentum, iw, j, 0, 0);
      FT rty_c = SUF(ld)(&ao->y_momentum, i, j, 0, 0), rty_s = SUF(ld)(&ao->y_momentum, i, j - 1, 0, 0);
      FT cx = have_ice ? (FT)0.5 * (SUF(ld)(&ice->concentration, iw, j, 0, 0) + conc) : (FT)0;
      FT cy = have_ice ? (FT)0.5 * (SUF(ld)(&ice->concentration, i, j - 1, 0, 0) + conc) : (FT)0;
      FT txao = (rtx_w + rtx_c) * (FT)0.5 * rho0inv;
      FT tyao = (rty_s + rty_c) * (FT)0.5 * rho0inv;
      FT txio = (io && io->x_momentum.ptr) ? SUF(ld)(&io->x_momentum, i, j, 0, 0) * rho0inv * cx : (FT)0;
      FT tyio = (io && io->y_momentum.ptr) ? SUF(ld)(&io->y_momentum, i, j, 0, 0) * rho0inv * cy : (FT)0;
      FT tx = ((FT)1 - cx) * txao + txio;
      FT ty = ((FT)1 - cy) * tyao + tyio;
      int act_w = SUF(active)(&ocean->mask, iw, j), act_s = SUF(active)(&ocean->mask, i, j - 1);
      if (!act || !act_w) tx = (FT)0;
      if (!act || !act_s) ty = (FT)0;
      if (!act) { JT = JS = J0 = (FT)0; Qu = Qal = Qts = (FT)0; }
      SUF(st)(&out->u, i, j, 0, tx); SUF(st)(&out->v, i, j, 0, ty);
      SUF(st)(&out->T, i, j, 0, JT); SUF(st)(&out->S, i, j, 0, JS);
      SUF(st)(&out->upwelling_longwave, i, j, 0, Qu); SUF(st)(&out->downwelling_longwave, i, j, 0, Qal);
      SUF(st)(&out->downwelling_shortwave, i, j, 0, Qts); SUF(st)(&out->penetrating_shortwave, i, j, 0, J0);
    }
  return 0;
}

/* a2: update_state! in the reference's order, un-fused */
int SUF(oracle_update_state)(const coflux_config* cfg, const coflux_update_inputs* in, coflux_update_outputs* out,
                             double time) {
  int rc = SUF(oracle_interpolate_atmosphere)(cfg, in->atmosphere, time, out->exchange);
  if (rc) return rc;
  if (in->land) { rc = SUF(oracle_interpolate_land)(cfg, in->land, time, out->exchange); if (rc) return rc; }
  rc = SUF(oracle_atmosphere_ocean_fluxes)(cfg, out->exchange, in->ocean, out->atmosphere_ocean);
  if (rc) return rc;
  return SUF(oracle_assemble_net_ocean_fluxes)(cfg, out->exchange, in->ocean, out->atmosphere_ocean, in->sea_ice,
                                               in->ice_ocean, out->net_ocean);
}

/* compute_net_sea_ice_fluxes!: what the sea-ice model receives.
 *   top    = (Q_d + Q_u + Q_c + Q_v)·[ℵ > 0],  Q_u = ε σ T_s⁴ (ice top temperature), Q_d = −(1 − α) Q_s − ε Q_ℓ
 *   bottom = Q_frazil + Q_interface
 * land cells 0; optional: atmosphere–ice stress ρτ averaged to the velocity points. */
int SUF(oracle_assemble_net_sea_ice_fluxes)(const coflux_config* cfg, const coflux_exchange_state* atmos, const coflux_ocean_surface* ocean,
                                            const coflux_sea_ice_state* ice, const coflux_interface_fluxes* ai,
                                            const coflux_ice_ocean_fluxes* io, coflux_net_sea_ice_fluxes* out) {
  const int Nx = cfg->grid.Nx, Ny = cfg->grid.Ny, r = cfg->grid.ring;
  const coflux_radiation_properties* R = &cfg->radiation;
  FT sigma = (FT)R->stefan_boltzmann_constant, emis = (FT)R->sea_ice_emissivity;
  const int px = (r == 0 && cfg->grid.periodic_x);
  coflux_array nomask; nomask.ptr = 0;
  const coflux_array* mask = ocean ? &ocean->mask : &nomask;
#pragma omp parallel for schedule(static)
  for (int j = 0; j < Ny; ++j)
    for (int i = 0; i < Nx; ++i) {
      int act = SUF(active)(mask, i, j);
      FT TsK = SUF(to_kelvin)(&cfg->ocean, SUF(ld)(&ice->top_temperature, i, j, 0, 0));
      FT conc = SUF(ld)(&ice->concentration, i, j, 0, 0);
      FT alpha = SUF(sea_ice_albedo)(cfg, ice, i, j, TsK);
      FT Qu = emis * sigma * TsK * TsK * TsK * TsK;
      FT Qd = -((FT)1 - alpha) * SUF(ld)(&atmos->Qs, i, j, 0, 0) - emis * SUF(ld)(&atmos->Ql, i, j, 0, 0);
      FT SQt = (conc > (FT)0) ? (Qd + Qu + SUF(ld)(&ai->sensible_heat, i, j, 0, 0) + SUF(ld)(&ai->latent_heat, i, j, 0, 0)) : (FT)0;
      FT Qf = (io && io->frazil_heat.ptr) ? SUF(ld)(&io->frazil_heat, i, j, 0, 0) : (FT)0;
      FT Qi = (io && io->interface_heat.ptr) ? SUF(ld)(&io->interface_heat, i, j, 0, 0) : (FT)0;
      SUF(st)(&out->top_heat, i, j, 0, act ? SQt : (FT)0);
      SUF(st)(&out->bottom_heat, i, j, 0, act ? (Qf + Qi) : (FT)0);
      if (out->top_u.ptr) {
        int iw = (px && i == 0) ? Nx - 1 : i - 1;
        FT v = (SUF(ld)(&ai->x_momentum, iw, j, 0, 0) + SUF(ld)(&ai->x_momentum, i, j, 0, 0)) * (FT)0.5;
        SUF(st)(&out->top_u, i, j, 0, (act && SUF(active)(mask, iw, j)) ? v : (FT)0);
      }
      if (out->top_v.ptr) {
        FT v = (SUF(ld)(&ai->y_momentum, i, j - 1, 0, 0) + SUF(ld)(&ai->y_momentum, i, j, 0, 0)) * (FT)0.5;
        SUF(st)(&out->top_v, i, j, 0, (act && SUF(active)(mask, i, j - 1)) ? v : (FT)0);
      }
    }
  return 0;
}

/* Time-averaged flux diagnostics (§8f row 4): the reference writes the flux fields under
 * `schedule = AveragedTimeInterval(...)` (/root/reference/src/OMIPConfigurations/omip_diagnostics.jl:152-158), i.e. Oceananigans'
 * WindowedTimeAverage, which accumulates  result ← (result·T + field·Δt) / (T + Δt)  at every iteration inside the window
 * (T: time already accumulated).  Fields: omip_diagnostics.jl:77-89, 125-148. */
static void SUF(avg_update)(const coflux_array* a, int i, int j, FT x, FT T, FT dt) {
  if (!a->ptr) return;
  FT* p = (FT*)a->ptr + SUF(idx)(a, i, j, 0, 0);
  *p = (*p * T + x * dt) / (T + dt);
}
int SUF(oracle_accumulate_flux_averages)(const coflux_config* cfg, const coflux_net_ocean_fluxes* net, const coflux_interface_fluxes* ao,
                                         const coflux_sea_ice_state* ice, const coflux_ice_ocean_fluxes* io, const coflux_flux_averages* v) {
  const int Nx = cfg->grid.Nx, Ny = cfg->grid.Ny;
  FT T = (FT)v->previous_interval, dt = (FT)v->dt;
  FT rho0 = (FT)cfg->ocean.reference_density, c0 = (FT)cfg->ocean.heat_capacity;
  FT rho0inv = (FT)1 / rho0;
  for (int j = 0; j < Ny; ++j)
    for (int i = 0; i < Nx; ++i) {
      SUF(avg_update)(&v->tau_x, i, j, SUF(ld)(&net->u, i, j, 0, 0), T, dt);
      SUF(avg_update)(&v->tau_y, i, j, SUF(ld)(&net->v, i, j, 0, 0), T, dt);
      FT JT = SUF(ld)(&net->T, i, j, 0, 0);
      SUF(avg_update)(&v->JT, i, j, JT, T, dt);
      SUF(avg_update)(&v->JS, i, j, SUF(ld)(&net->S, i, j, 0, 0), T, dt);
      if (ao) {
        SUF(avg_update)(&v->Qc, i, j, SUF(ld)(&ao->sensible_heat, i, j, 0, 0), T, dt);
        SUF(avg_update)(&v->Qv, i, j, SUF(ld)(&ao->latent_heat, i, j, 0, 0), T, dt);
      }
      FT JTio = (io && io->interface_heat.ptr) ? SUF(ld)(&io->interface_heat, i, j, 0, 0) * rho0inv / c0 : (FT)0;
      SUF(avg_update)(&v->JT_ice_ocean, i, j, JTio, T, dt);
      SUF(avg_update)(&v->JT_atmosphere_ocean, i, j, JT - JTio, T, dt);
      FT conc = (ice && ice->concentration.ptr) ? SUF(ld)(&ice->concentration, i, j, 0, 0) : (FT)0;
      SUF(avg_update)(&v->JS_ice_ocean, i, j, (io && io->salt.ptr) ? SUF(ld)(&io->salt, i, j, 0, 0) * conc : (FT)0, T, dt);
      if (io && io->frazil_heat.ptr) SUF(avg_update)(&v->JT_frazil, i, j, SUF(ld)(&io->frazil_heat, i, j, 0, 0) / rho0 / c0, T, dt);
    }
  return 0;
}

/* scalar probes for unit tests */
/* Closure surface-forcing front ends — literal restatement of the in-tree consumers of the net fluxes:
 *   /root/reference/src/OMIPConfigurations/KPP/kpp_surface_forcing.jl:18-22    u★  = max(sqrt(sqrt(τx^2 + τy^2)), minimum_friction_velocity)
 *   …/KPP/kpp_surface_forcing.jl:28-29                                         Bo  = − top_buoyancy_flux = −g (α Jᵀ − β Jˢ)
 *   …/NEMOTKE/nemo_tke_surface_forcing.jl:14-22                                u★² = sqrt(τx^2 + τy^2);  e = max(minimum_surface_TKE, Cᵇ u★²)
 * τx, τy are taken at the cell's own (i, j), as the reference does. */
int SUF(oracle_closure_surface_forcing)(const coflux_config* cfg, const coflux_net_ocean_fluxes* net, const coflux_closure_forcing* f) {
  const int Nx = cfg->grid.Nx, Ny = cfg->grid.Ny;
  for (int j = 0; j < Ny; ++j)
    for (int i = 0; i < Nx; ++i) {
      const FT tx = SUF(ld)(&net->u, i, j, 0, 0), ty = SUF(ld)(&net->v, i, j, 0, 0);
      const FT ustar2 = SQRT(tx * tx + ty * ty);
      SUF(st)(&f->friction_velocity_squared, i, j, 0, ustar2);
      SUF(st)(&f->friction_velocity, i, j, 0, FMAX(SQRT(SQRT(tx * tx + ty * ty)), (FT)f->minimum_friction_velocity));
      SUF(st)(&f->surface_tke, i, j, 0, FMAX((FT)f->minimum_surface_tke, (FT)f->Cb * ustar2));
      if (f->buoyancy_flux.ptr && f->thermal_expansion.ptr && f->haline_contraction.ptr) {
        const FT JT = SUF(ld)(&net->T, i, j, 0, 0), JS = SUF(ld)(&net->S, i, j, 0, 0);
        const FT top = (FT)f->gravitational_acceleration *
                       (SUF(ld)(&f->thermal_expansion, i, j, 0, 0) * JT - SUF(ld)(&f->haline_contraction, i, j, 0, 0) * JS);
        SUF(st)(&f->buoyancy_flux, i, j, 0, -top);
      }
    }
  return 0;
}

/* NormalizeSalinity — restatement of /root/reference/src/OMIPConfigurations/omip_simulation.jl:187-220:
 *   compute!(mean_total)                        mean_total = Field(Average(flux_field [+ additional_buffer], dims=(1,2)))
 *   parent(n.flux_field) .-= n.mean_total       (whole parent, halos included)
 * Oceananigans' Average over (1,2) of a (Center, Center, Nothing) field is ∫ f dA / ∫ dA over the wet interior.
 * Sums in long double, serial order (the checker does not care about speed here).  sums_out (may be NULL)
 * receives {Σ f·Az, Σ Az}; when `mean_in` is non-NULL that mean is subtracted instead of the local one
 * (multi-slab case: the caller combines the slabs' sums). */
int SUF(oracle_normalize_salinity_flux)(const coflux_config* cfg, const coflux_salinity_normalization* n, double* sums_out,
                                        const double* mean_in) {
  const int Nx = cfg->grid.Nx, Ny = cfg->grid.Ny;
  long double num = 0.0L, den = 0.0L;
  for (int j = 0; j < Ny; ++j)
    for (int i = 0; i < Nx; ++i) {
      if (!SUF(active)(&n->mask, i, j)) continue;
      long double f = (long double)SUF(ld)(&n->flux, i, j, 0, 0);
      if (n->additional.ptr) f += (long double)SUF(ld)(&n->additional, i, j, 0, 0);
      const long double A = (long double)SUF(ld)(&n->area, i, j, 0, 0);
      num += f * A;
      den += A;
    }
  if (sums_out) { sums_out[0] = (double)num; sums_out[1] = (double)den; }
  const FT mean = mean_in ? (FT)*mean_in : ((den != 0.0L) ? (FT)(double)(num / den) : (FT)0);
  const int Hi = n->flux.off_i, Hj = n->flux.off_j;
  for (int j = -Hj; j < Ny + Hj; ++j)
    for (int i = -Hi; i < Nx + Hi; ++i) SUF(st)(&n->flux, i, j, 0, SUF(ld)(&n->flux, i, j, 0, 0) - mean);
  return 0;
}

void SUF(oracle_probe_psi)(int kind, int n, const FT* zeta, FT* psi_m, FT* psi_s) {
  for (int k = 0; k < n; ++k) { psi_m[k] = SUF(psi_momentum)(kind, zeta[k]); psi_s[k] = SUF(psi_scalar)(kind, zeta[k]); }
}
void SUF(oracle_probe_saturation)(const coflux_config* cfg, int n, const FT* T, FT* p_liq, FT* p_ice) {
  SUF(thermo_consts) c = SUF(make_thermo)(&cfg->atmosphere.thermodynamics);
  for (int k = 0; k < n; ++k) {
    p_liq[k] = SUF(saturation_vapor_pressure_liquid)(&c, T[k]);
    p_ice[k] = SUF(saturation_vapor_pressure_ice)(&c, T[k]);
  }
}
void SUF(oracle_probe_thermo)(const coflux_config* cfg, FT p, FT T, FT q, FT* out7) {
  SUF(thermo_consts) c = SUF(make_thermo)(&cfg->atmosphere.thermodynamics);
  SUF(thermo_state) s = SUF(phase_equil_pTq)(&c, p, T, q);
  out7[0] = s.rho; out7[1] = s.cp_m; out7[2] = s.q_vap; out7[3] = s.T_v; out7[4] = s.q_liq; out7[5] = s.q_ice;
  out7[6] = SUF(latent_heat_vapor)(&c, T);
}
/* solve one atmosphere–ocean cell: in8 = ua va Ta pa qa uo vo To[K], So ; out = u★ θ★ q★ iterations */
void SUF(oracle_probe_solve)(const coflux_config* cfg, const FT* in9, FT* out4) {
  SUF(thermo_consts) c = SUF(make_thermo)(&cfg->atmosphere.thermodynamics);
  SUF(cell_in) in;
  in.ua = in9[0]; in.va = in9[1]; in.Ta = in9[2]; in.pa = in9[3]; in.qa = in9[4];
  in.uo = in9[5]; in.vo = in9[6]; in.To = in9[7]; in.So = in9[8];
  in.Qs = in.Ql = in.h_ice = in.S_ice = in.albedo = (FT)0;
  SUF(thermo_state) atm; FT du, dv;
  SUF(scales) s = SUF(solve_cell)(&cfg->atmosphere_ocean, cfg, &c, &in, 0, &atm, &du, &dv);
  out4[0] = s.ustar; out4[1] = s.tstar; out4[2] = s.qstar; out4[3] = (FT)s.iterations;
}
/* solve one atmosphere–sea-ice cell (skin temperature): in13 = ua va Ta pa qa ui vi Ttop[K] Qs Ql h_ice S_ice albedo;
 * out5 = u★ θ★ q★ T_s[K] iterations */
void SUF(oracle_probe_solve_ice)(const coflux_config* cfg, const FT* in13, FT* out5) {
  SUF(thermo_consts) c = SUF(make_thermo)(&cfg->atmosphere.thermodynamics);
  SUF(cell_in) in;
  in.ua = in13[0]; in.va = in13[1]; in.Ta = in13[2]; in.pa = in13[3]; in.qa = in13[4];
  in.uo = in13[5]; in.vo = in13[6]; in.To = in13[7]; in.So = (FT)0;
  in.Qs = in13[8]; in.Ql = in13[9]; in.h_ice = in13[10]; in.S_ice = in13[11]; in.albedo = in13[12];
  SUF(thermo_state) atm; FT du, dv;
  SUF(scales) s = SUF(solve_cell)(&cfg->atmosphere_sea_ice, cfg, &c, &in, 1, &atm, &du, &dv);
  out5[0] = s.ustar; out5[1] = s.tstar; out5[2] = s.qstar; out5[3] = s.Ts; out5[4] = (FT)s.iterations;
}

"""Piecewise-polynomial tables of stability functions for the CUDA kernels:

  COFLUX_PSI_TABLE_*       unstable Edson et al. (2013) ψ_u(ζ), ψ_θ(ζ), ζ < 0     (hot loop, coflux_solve_tile.cuh)
  COFLUX_PSI_PAULSON_*     unstable Paulson (1970) ψ_m, ψ_h with γ = 16, ζ < 0     (sea-ice / Large–Yeager solves)
  COFLUX_PSI_SHEBA_*       stable Grachev et al. (2007) SHEBA ψ_m, ψ_h, ζ > 0      (sea-ice solve)

Domain: −ζ ∈ [2^KMIN, 2^KMAX).  Each binade [2^k, 2^(k+1)) is split into NS equal sub-intervals; on each a
degree-DEG polynomial in the local variable t ∈ [−1, 1) interpolates the function at Chebyshev nodes
(computed in 50-digit arithmetic, coefficients rounded to double).  The script verifies the tables
against the exact formulas at 40 points per interval, evaluating the polynomial in double precision
exactly as the device does (Horner with fma), and writes climaocean.jl_b200/csrc/coflux_psi_table.h.

    python tools/gen_psi_table.py            # ≈ 1 min
"""
import os
import sys

import mpmath as mp
import numpy as np

mp.mp.dps = 50
KMIN, KMAX, NS, DEG = -30, 13, 16, 7
KMAX_ICE = 26      # the Paulson / SHEBA tables of the sea-ice solve reach |ζ| < 2²⁶: calm, strongly stratified cells get there


def conv(y):
    r3 = mp.sqrt(3)
    return mp.mpf("1.5") * mp.log((1 + y + y * y) / 3) - r3 * mp.atan((1 + 2 * y) / r3) + mp.pi / r3


def psi_u(z):      # z = ζ < 0
    x = (1 - 15 * z) ** mp.mpf("0.25")
    pk = 2 * mp.log((1 + x) / 2) + mp.log((1 + x * x) / 2) - 2 * mp.atan(x) + mp.pi / 2
    pc = conv(mp.cbrt(1 - mp.mpf("10.15") * z))
    f = z * z / (1 + z * z)
    return (1 - f) * pk + f * pc


def psi_t(z):
    x = mp.sqrt(1 - 15 * z)
    pk = 2 * mp.log((1 + x) / 2)
    pc = conv(mp.cbrt(1 - mp.mpf("34.15") * z))
    f = z * z / (1 + z * z)
    return (1 - f) * pk + f * pc


def paulson_m(z):   # z = ζ < 0
    x = (1 - 16 * z) ** mp.mpf("0.25")
    return 2 * mp.log((1 + x) / 2) + mp.log((1 + x * x) / 2) - 2 * mp.atan(x) + mp.pi / 2


def paulson_h(z):
    return 2 * mp.log((1 + mp.sqrt(1 - 16 * z)) / 2)


def sheba_m(mz):    # called with mz = −ζ (the tables are indexed by a positive argument): ζ = −mz > 0
    z = -mz
    a, b = mp.mpf(5), mp.mpf(5) / mp.mpf("6.5")
    r3 = mp.sqrt(3)
    x = mp.cbrt(1 + z)
    B = mp.cbrt((1 - b) / b)
    p1 = -3 * a * (x - 1) / b
    p2 = a * B / (2 * b) * (2 * mp.log((x + B) / (1 + B)) - mp.log((x * x - B * x + B * B) / (1 - B + B * B))
                            + 2 * r3 * (mp.atan((2 * x - B) / (r3 * B)) - mp.atan((2 - B) / (r3 * B))))
    return p1 + p2


def sheba_h(mz):
    z = -mz
    a, b, c = mp.mpf(5), mp.mpf(5), mp.mpf(3)
    B = mp.sqrt(c * c - 4)
    p1 = -b / 2 * mp.log(1 + c * z + z * z)
    p2 = (-a / B + b * c / (2 * B)) * (mp.log((2 * z + c - B) / (2 * z + c + B)) - mp.log((c - B) / (c + B)))
    return p1 + p2


def cheb_fit(f, a, b, deg):
    """Monomial coefficients c[0..deg] in t ∈ [−1,1] of the Chebyshev interpolant of f on [a,b] (a,b are −ζ)."""
    n = deg + 1
    nodes = [mp.cos(mp.pi * (2 * k + 1) / (2 * n)) for k in range(n)]
    vals = [f(-((a + b) / 2 + (b - a) / 2 * t)) for t in nodes]
    # Chebyshev coefficients
    ck = []
    for j in range(n):
        s = mp.fsum(vals[k] * mp.cos(mp.pi * j * (2 * k + 1) / (2 * n)) for k in range(n))
        ck.append(s * (2 if j else 1) / n)
    # convert Σ ck T_k(t) to monomials
    T = [[mp.mpf(1)], [mp.mpf(0), mp.mpf(1)]]
    for k in range(2, n):
        Tk = [mp.mpf(0)] + [2 * c for c in T[k - 1]]
        for i, c in enumerate(T[k - 2]):
            Tk[i] -= c
        T.append(Tk)
    mono = [mp.mpf(0)] * n
    for j in range(n):
        for i, c in enumerate(T[j]):
            mono[i] += ck[j] * c
    return [float(c) for c in mono]


def horner(c, t):
    acc = c[-1]
    for v in reversed(c[:-1]):
        acc = np.float64(acc) * np.float64(t) + np.float64(v)      # numpy double ops (fma differs by ≤ 1 ulp of each step)
    return float(acc)


def build(fu, ft, label, kmax=None):
    rows, worst_u, worst_t = [], 0.0, 0.0
    for k in range(KMIN, KMAX if kmax is None else kmax):
        for j in range(NS):
            a = mp.mpf(2) ** k * (1 + mp.mpf(j) / NS)
            b = mp.mpf(2) ** k * (1 + mp.mpf(j + 1) / NS)
            cu, ct = cheb_fit(fu, a, b, DEG), cheb_fit(ft, a, b, DEG)
            rows.append((cu, ct))
            for m in range(40):
                t = -1 + 2 * (m + 0.37) / 40
                z = -((a + b) / 2 + (b - a) / 2 * mp.mpf(t))
                eu = abs(mp.mpf(horner(cu, t)) - fu(z)) / max(1, abs(fu(z)))
                et = abs(mp.mpf(horner(ct, t)) - ft(z)) / max(1, abs(ft(z)))
                worst_u, worst_t = max(worst_u, float(eu)), max(worst_t, float(et))
        print(f"{label} binade 2^{k}: worst so far {worst_u:.2e} {worst_t:.2e}", file=sys.stderr)
    return rows, worst_u, worst_t


def write_table(fh, rows, base):
    for ctype, name, suffix in (("double", base + "_F64", ""), ("float", base + "_F32", "f")):
        fh.write("__device__ __align__(128) const %s %s[%d][2][%d] = {\n" % (ctype, name, len(rows), DEG + 1))
        for cu, ct in rows:
            fmt = (lambda v: repr(v)) if ctype == "double" else (lambda v: repr(float(np.float32(v))) + "f")
            fh.write("  {{" + ", ".join(fmt(v) for v in cu) + "},\n   {" + ", ".join(fmt(v) for v in ct) + "}},\n")
        fh.write("};\n")


def main():
    tables = [(psi_u, psi_t, "COFLUX_PSI_TABLE", "unstable Edson et al. (2013) ψ_u, ψ_θ on −ζ", KMAX),
              (paulson_m, paulson_h, "COFLUX_PSI_PAULSON", "unstable Paulson (1970, γ = 16) ψ_m, ψ_h on −ζ, up to 2^%d" % KMAX_ICE, KMAX_ICE),
              (sheba_m, sheba_h, "COFLUX_PSI_SHEBA", "stable Grachev et al. (2007) SHEBA ψ_m, ψ_h on ζ, up to 2^%d" % KMAX_ICE, KMAX_ICE)]
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "climaocean.jl_b200", "csrc", "coflux_psi_table.h")
    built = [(build(fu, ft, base, kmax), base, what) for fu, ft, base, what, kmax in tables]
    with open(out, "w") as fh:
        fh.write("// GENERATED by tools/gen_psi_table.py — do not edit.\n")
        fh.write("// Piecewise degree-%d polynomials on [2^%d, 2^%d), %d sub-intervals per binade; max error relative to max(1,|ψ|)\n" % (DEG, KMIN, KMAX, NS))
        fh.write("// (40 points per interval, double Horner):\n")
        for (rows, wu, wt), base, what in built:
            fh.write("//   %-20s %s: %.2e, %.2e\n" % (base, what, wu, wt))
        fh.write("#pragma once\n#define COFLUX_PSI_KMIN (%d)\n#define COFLUX_PSI_KMAX (%d)\n#define COFLUX_PSI_NS %d\n#define COFLUX_PSI_DEG %d\n#define COFLUX_PSI_ICE_KMAX (%d)\n"
                 % (KMIN, KMAX, NS, DEG, KMAX_ICE))
        for (rows, wu, wt), base, what in built:
            write_table(fh, rows, base)
    for (rows, wu, wt), base, what in built:
        print(f"{base}: {len(rows)} intervals, worst relative error {wu:.2e}, {wt:.2e}")
    print("wrote", out)


if __name__ == "__main__":
    main()

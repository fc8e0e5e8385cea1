"""Generate Taylor coefficients (about 0) of the Edson et al. (2013) stability functions on each
branch, for the small-argument evaluation of ψ(ℓ/L) in the CUDA solve (coflux_solve_tile.cuh).
Prints C arrays and the worst absolute error on |x| ≤ X0 against 60-digit mpmath."""
import mpmath as mp
mp.mp.dps = 60
X0 = mp.mpf(2) ** -9

def conv(y):
    r3 = mp.sqrt(3)
    return mp.mpf("1.5") * mp.log((1 + y + y * y) / 3) - r3 * mp.atan((1 + 2 * y) / r3) + mp.pi / r3
def pm_unst(z):
    x = (1 - 15 * z) ** mp.mpf("0.25")
    pk = 2 * mp.log((1 + x) / 2) + mp.log((1 + x * x) / 2) - 2 * mp.atan(x) + mp.pi / 2
    pc = conv(mp.cbrt(1 - mp.mpf("10.15") * z)); f = z * z / (1 + z * z)
    return (1 - f) * pk + f * pc
def pm_stab(z):
    return -(mp.mpf("0.7") * z + mp.mpf("0.75") * (z - 5 / mp.mpf("0.35")) * mp.exp(-mp.mpf("0.35") * z) + mp.mpf("0.75") * 5 / mp.mpf("0.35"))
def ps_unst(z):
    x = mp.sqrt(1 - 15 * z); pk = 2 * mp.log((1 + x) / 2)
    pc = conv(mp.cbrt(1 - mp.mpf("34.15") * z)); f = z * z / (1 + z * z)
    return (1 - f) * pk + f * pc
def ps_stab(z):
    return -((1 + mp.mpf(2) / 3 * z) ** mp.mpf("1.5") + mp.mpf(2) / 3 * (z - mp.mpf("14.28")) * mp.exp(-mp.mpf("0.35") * z) + mp.mpf("8.525"))

def gen(name, f, deg, sign):
    c = mp.taylor(f, 0, deg)
    cd = [float(v) for v in c]
    worst = 0
    for k in range(0, 401):
        x = sign * X0 * k / 400
        xd = float(x)
        acc = 0.0
        for v in reversed(cd):
            acc = acc * xd + v           # double Horner (python floats)
        worst = max(worst, abs(mp.mpf(acc) - f(mp.mpf(xd))))
    print(f"// {name}: degree {deg}, max abs err on |x|<=2^-9: {float(worst):.2e}")
    print(f"static __device__ const double {name}[{deg+1}] = {{" + ", ".join(f"{v!r}" for v in cd) + "};")

gen("PSI_M_UNST", pm_unst, 13, -1)
gen("PSI_M_STAB", pm_stab, 7, +1)
gen("PSI_S_UNST", ps_unst, 15, -1)
gen("PSI_S_STAB", ps_stab, 7, +1)

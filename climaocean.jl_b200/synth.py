"""Synthetic inputs for the surface-flux path.

Every value is produced by integer hashing (SplitMix64) and a fixed sequence of IEEE-754 basic
operations (+ − × ÷ on float64; no libm), so that C++, Python and Julia produce bit-identical
inputs.  value = lo + (hi − lo) · (½·u01 + ½·smooth), u01 = (splitmix64(seed + k) >> 11) · 2⁻⁵³,
smooth = a triangle-wave pattern of the integer cell indices.  Longitude indices are wrapped
periodically with the GLOBAL Nx so that halos and longitude slabs are consistent.
"""
import numpy as np

MASK64 = np.uint64(0xFFFFFFFFFFFFFFFF)
SEED_BASE = 0xC0F10000


def splitmix64(x):
    """Vectorised SplitMix64 finaliser of uint64 counters x (already seed + k·γ)."""
    with np.errstate(over="ignore"):
        z = x.astype(np.uint64)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z


def u01(field_id, k):
    """Uniform [0,1) doubles for integer counters k (any shape, int64 ≥ 0)."""
    with np.errstate(over="ignore"):
        seed = np.uint64(SEED_BASE + field_id)
        x = seed + (k.astype(np.uint64) + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
    return (splitmix64(x) >> np.uint64(11)).astype(np.float64) * (2.0 ** -53)


def tri(m, N):
    """Triangle wave |2·(m mod N)/N − 1| ∈ [0,1] from integers (exact IEEE ops)."""
    return np.abs(2.0 * (np.mod(m, N).astype(np.float64) / float(N)) - 1.0)


def pattern(field_id, lo, hi, ig, jg, Nxg, Nyg, level=0, random_weight=0.5):
    """2-D synthetic plane on global integer indices ig (periodic, shape (ni,)) × jg (shape (nj,))."""
    I = np.mod(ig, Nxg).astype(np.int64)[None, :]
    J = jg.astype(np.int64)[:, None]
    k = (J + 64) * np.int64(Nxg) + I + np.int64(level) * np.int64(Nxg) * np.int64(Nyg + 128)
    r = u01(field_id, k)
    p1, p2 = 1 + (field_id % 3), 1 + ((field_id // 3) % 2)
    s = 0.5 * tri(I * p1 + (field_id * 37) % Nxg, Nxg) + 0.5 * tri((J + 64) * p2 + (field_id * 11) % Nyg, Nyg)
    s = np.broadcast_to(s, r.shape)
    w = random_weight
    return lo + (hi - lo) * (w * r + (1.0 - w) * s)


# field ids
F_OCEAN_U, F_OCEAN_V, F_OCEAN_T, F_OCEAN_S = 1, 2, 3, 4
F_ATM = {"u": 10, "v": 11, "T": 12, "q": 13, "p": 14, "Qs": 15, "Ql": 16, "rain": 17, "snow": 18}
F_ICE_CONC, F_ICE_H, F_ICE_S, F_ICE_U, F_ICE_V, F_ICE_T, F_ICE_HPREV, F_MASK = 30, 31, 32, 33, 34, 35, 36, 40
F_ICE_SNOW = 37
F_LAND = {"rivers": 20, "icebergs": 21}
LAND_RANGES = {"rivers": (-3e-4, 1.5e-4), "icebergs": (-6e-4, 1e-4)}    # clipped at 0: runoff / calving are non-zero in patches only

ATM_RANGES = {"u": (-25.0, 25.0), "v": (-25.0, 25.0), "T": (250.0, 305.0), "q": (1e-4, 2e-2), "p": (9.6e4, 1.04e5),
              "Qs": (0.0, 1000.0), "Ql": (100.0, 450.0), "rain": (0.0, 3e-4), "snow": (0.0, 3e-4)}


def ocean_state(grid, dtype=None, frazil=False):
    """Ocean u, v, T, S parents, shape (Nz+2Hz, Ny+2Hy, Nx+2Hx).  T ∈ [−1.8, 30] °C at the surface,
    cooling with depth; with frazil=True a band of cells sits below the local freezing point."""
    dtype = dtype or grid.dtype
    Hx, Hy, Hz = grid.halo
    Nxg = grid.global_Nx or grid.Nx
    ig = np.arange(-Hx, grid.Nx + Hx) + grid.i_offset
    jg = np.arange(-Hy, grid.Ny + Hy)
    nk = grid.Nz + 2 * Hz
    out = {}
    for name, fid, lo, hi in (("u", F_OCEAN_U, -1.0, 1.0), ("v", F_OCEAN_V, -1.0, 1.0), ("T", F_OCEAN_T, -1.8, 30.0),
                              ("S", F_OCEAN_S, 30.0, 38.0)):
        a = np.empty((nk, jg.size, ig.size), dtype=np.float64)
        for kk in range(nk):
            k = kk - Hz
            depth_frac = 0.0 if k >= grid.Nz - 1 else (grid.Nz - 1 - max(k, 0)) / float(max(grid.Nz - 1, 1))
            plane = pattern(fid, lo, hi, ig, jg, Nxg, grid.Ny, level=max(min(k, grid.Nz - 1), 0))
            if name == "T":
                plane = plane - depth_frac * (plane + 1.0) * 0.9      # relax towards −1 °C at depth
                if frazil:
                    cold = tri(jg[:, None] + 64 + 3 * max(k, 0), 16) > 0.75   # bands of super-cooled water
                    plane = np.where(cold, -2.6 - 0.2 * depth_frac, plane)
            elif name in ("u", "v"):
                plane = plane * (1.0 - 0.8 * depth_frac)
            a[kk] = plane
        out[name] = a.astype(dtype)
    return out


def atmosphere_series(Nxa=640, Nya=320, Nt=8, halo=3, dtype=np.float64, dt_hours=3.0):
    """JRA55-like prescribed atmosphere on a regular Nxa×Nya source grid, Nt levels 3 h apart.
    Parents have shape (Nt, 1, Nya+2H, Nxa+2H); λ halos periodic, φ halos by index continuation."""
    ig = np.arange(-halo, Nxa + halo)
    jg = np.arange(-halo, Nya + halo)
    times = np.arange(Nt, dtype=np.float64) * dt_hours * 3600.0
    out = {}
    for name, fid in F_ATM.items():
        lo, hi = ATM_RANGES[name]
        a = np.empty((Nt, 1, jg.size, ig.size), dtype=np.float64)
        for n in range(Nt):
            a[n, 0] = pattern(fid, lo, hi, ig, jg, Nxa, Nya, level=n)
        out[name] = a.astype(dtype)
    return out, times


def land_series(Nxl=360, Nyl=180, Nt=4, halo=2, dtype=np.float64, dt_hours=24.0):
    """JRA55-do-like land freshwater forcing (friver, licalvf: jra55_data_staging.jl:8) on its own regular source grid and
    (daily) time axis.  Parents (Nt, 1, Nyl+2H, Nxl+2H), kg m⁻² s⁻¹, zero over most of the globe."""
    ig = np.arange(-halo, Nxl + halo)
    jg = np.arange(-halo, Nyl + halo)
    times = np.arange(Nt, dtype=np.float64) * dt_hours * 3600.0
    out = {}
    for name, fid in F_LAND.items():
        lo, hi = LAND_RANGES[name]
        a = np.empty((Nt, 1, jg.size, ig.size), dtype=np.float64)
        for n in range(Nt):
            a[n, 0] = np.maximum(pattern(fid, lo, hi, ig, jg, Nxl, Nyl, level=n), 0.0)
        out[name] = a.astype(dtype)
    return out, times


def sea_ice_state(grid, dtype=None):
    """ℵ ∈ {0, (0,1), 1}, h ∈ [0,3] m (0 where ℵ = 0), S_i, ice velocities, top temperature (°C)."""
    dtype = dtype or grid.dtype
    Hx, Hy, _ = grid.halo
    Nxg = grid.global_Nx or grid.Nx
    ig = np.arange(-Hx, grid.Nx + Hx) + grid.i_offset
    jg = np.arange(-Hy, grid.Ny + Hy)
    c = pattern(F_ICE_CONC, -0.5, 1.5, ig, jg, Nxg, grid.Ny)
    conc = np.clip(c, 0.0, 1.0)
    h = pattern(F_ICE_H, 0.0, 3.0, ig, jg, Nxg, grid.Ny)
    h = np.where(conc > 0.0, h, 0.0)
    hprev = np.where(conc > 0.0, pattern(F_ICE_HPREV, 0.0, 3.0, ig, jg, Nxg, grid.Ny), 0.0)
    out = {"concentration": conc, "thickness": h, "previous_thickness": hprev,
           "salinity": pattern(F_ICE_S, 2.0, 8.0, ig, jg, Nxg, grid.Ny),
           "u": pattern(F_ICE_U, -0.3, 0.3, ig, jg, Nxg, grid.Ny), "v": pattern(F_ICE_V, -0.3, 0.3, ig, jg, Nxg, grid.Ny),
           "top_temperature": pattern(F_ICE_T, -30.0, -0.5, ig, jg, Nxg, grid.Ny),
           "snow_thickness": np.where(conc > 0.0, np.maximum(pattern(F_ICE_SNOW, -0.2, 0.5, ig, jg, Nxg, grid.Ny), 0.0), 0.0)}
    return {k: v[None].astype(dtype) for k, v in out.items()}


def land_mask(grid, land_fraction):
    """uint8 parent (1, Ny+2Hy, Nx+2Hx): 1 = wet.  land_fraction = 0 → all wet."""
    Hx, Hy, _ = grid.halo
    Nxg = grid.global_Nx or grid.Nx
    ig = np.arange(-Hx, grid.Nx + Hx) + grid.i_offset
    jg = np.arange(-Hy, grid.Ny + Hy)
    r = pattern(F_MASK, 0.0, 1.0, ig, jg, Nxg, grid.Ny, random_weight=1.0)
    return (r >= land_fraction).astype(np.uint8)[None]

"""Oceananigans-style halo-padded fields and grids (host-side mirror; plumbing only).

A `Field` wraps either a numpy array (host; used to feed the CPU oracle in tests) or a torch tensor
(device memory for the CUDA path).  Memory layout equals an Oceananigans "parent" array: size
(Nx+2Hx, Ny+2Hy, Nz+2Hz) column-major, i fastest — i.e. a C-ordered array of shape (nk, nj, ni).
A `FieldTimeSeries` adds a leading time axis: (Nt, nk, nj, ni).
"""
import numpy as np

from . import _abi


def _is_torch(x):
    return type(x).__module__.startswith("torch")


class Field:
    def __init__(self, data, halo=(0, 0, 0), name=""):
        assert data.ndim == 3, "Field data must have shape (nk, nj, ni)"
        self.data = data
        self.halo = tuple(int(h) for h in halo)
        self.name = name

    # ---- construction ----
    @classmethod
    def zeros(cls, size, halo, dtype, device=None, name="", fill=0.0):
        Nx, Ny, Nz = size
        Hx, Hy, Hz = halo
        shape = (Nz + 2 * Hz, Ny + 2 * Hy, Nx + 2 * Hx)
        if device is None:
            data = np.full(shape, fill, dtype=dtype)
        else:
            import torch
            tdt = {np.float64: torch.float64, np.float32: torch.float32, np.uint8: torch.uint8,
                   np.int32: torch.int32}[np.dtype(dtype).type]
            data = torch.full(shape, fill, dtype=tdt, device=device)
        return cls(data, halo, name)

    @classmethod
    def from_numpy(cls, arr, halo, device=None, name=""):
        """arr: numpy (nk, nj, ni) parent INCLUDING halos."""
        if device is None:
            return cls(np.ascontiguousarray(arr), halo, name)
        import torch
        return cls(torch.from_numpy(np.ascontiguousarray(arr)).to(device), halo, name)

    # ---- views ----
    @property
    def size(self):
        nk, nj, ni = self.data.shape
        Hx, Hy, Hz = self.halo
        return (ni - 2 * Hx, nj - 2 * Hy, nk - 2 * Hz)

    @property
    def interior(self):
        Hx, Hy, Hz = self.halo
        nk, nj, ni = self.data.shape
        return self.data[Hz:nk - Hz, Hy:nj - Hy, Hx:ni - Hx]

    def numpy(self):
        if _is_torch(self.data):
            return self.data.detach().cpu().numpy()
        return self.data

    def to(self, device):
        return Field.from_numpy(self.numpy().copy(), self.halo, device, self.name)

    def clone(self):
        d = self.data.clone() if _is_torch(self.data) else self.data.copy()
        return Field(d, self.halo, self.name)

    # ---- ABI descriptor ----
    def ptr(self):
        return self.data.data_ptr() if _is_torch(self.data) else self.data.ctypes.data

    def array(self):
        nk, nj, ni = self.data.shape
        Hx, Hy, Hz = self.halo
        if _is_torch(self.data):
            assert self.data.is_contiguous()
        else:
            assert self.data.flags["C_CONTIGUOUS"]
        return _abi.Array(self.ptr(), 1, ni, ni * nj, 0, Hx, Hy, Hz, 0)


class FieldTimeSeries:
    def __init__(self, data, halo, times, name=""):
        assert data.ndim == 4, "series data must have shape (Nt, nk, nj, ni)"
        self.data = data
        self.halo = tuple(int(h) for h in halo)
        self.times = np.ascontiguousarray(times, dtype=np.float64)
        self.name = name

    @classmethod
    def from_numpy(cls, arr, halo, times, device=None, name=""):
        if device is None:
            return cls(np.ascontiguousarray(arr), halo, times, name)
        import torch
        return cls(torch.from_numpy(np.ascontiguousarray(arr)).to(device), halo, times, name)

    def numpy(self):
        return self.data.detach().cpu().numpy() if _is_torch(self.data) else self.data

    def to(self, device):
        return FieldTimeSeries.from_numpy(self.numpy().copy(), self.halo, self.times, device, self.name)

    def array(self):
        nt, nk, nj, ni = self.data.shape
        Hx, Hy, Hz = self.halo
        ptr = self.data.data_ptr() if _is_torch(self.data) else self.data.ctypes.data
        return _abi.Array(ptr, 1, ni, ni * nj, ni * nj * nk, Hx, Hy, Hz, 0)


NULL_ARRAY = _abi.Array(None, 0, 0, 0, 0, 0, 0, 0, 0)


def arr(field):
    """Descriptor of a Field / FieldTimeSeries, or the NULL descriptor for None."""
    return NULL_ARRAY if field is None else field.array()


class LatitudeLongitudeGrid:
    """Regular lat-lon grid: λ periodic over `longitude`, φ bounded over `latitude`
    (the grids of BASELINE.json configs 1, 2, 4)."""

    def __init__(self, size, longitude=(0.0, 360.0), latitude=(-75.0, 75.0), z=(-5000.0, 0.0), halo=(7, 7, 7),
                 dtype=np.float64):
        self.Nx, self.Ny, self.Nz = (int(s) for s in size)
        self.longitude, self.latitude, self.z = longitude, latitude, z
        self.halo = tuple(halo)
        self.dtype = np.dtype(dtype)

    @property
    def size(self):
        return (self.Nx, self.Ny, self.Nz)

    def lambda_centers(self, i):
        # a slab evaluates the GLOBAL grid's expression at its global column index, so that its coordinates — and the
        # fractional source indices built from them — carry the same bits as a one-process solve of the whole grid
        L0, L1 = self.global_longitude or self.longitude
        return L0 + (np.asarray(i, dtype=np.float64) + self.i_offset + 0.5) * ((L1 - L0) / (self.global_Nx or self.Nx))

    def phi_centers(self, j):
        P0, P1 = self.latitude
        return P0 + (np.asarray(j, dtype=np.float64) + 0.5) * ((P1 - P0) / self.Ny)

    def dz(self):
        """Uniform layer thickness Δz(k) as a (Nz+2Hz,)-long 1-D parent."""
        Hz = self.halo[2]
        d = (self.z[1] - self.z[0]) / self.Nz
        return np.full(self.Nz + 2 * Hz, d, dtype=self.dtype)

    def horizontal_areas(self, radius=6371e3):
        """Az(j) of the (Center, Center) cells, R² Δλ (sin φ_{j+½} − sin φ_{j−½}), as a 1-D parent of Ny + 2Hy rows
        (what Oceananigans' `Average(…, dims=(1,2))` weights a surface field with on a LatitudeLongitudeGrid)."""
        Hy = self.halo[1]
        P0, P1 = self.latitude
        dphi = (P1 - P0) / self.Ny
        j = np.arange(-Hy, self.Ny + Hy, dtype=np.float64)
        south, north = np.deg2rad(P0 + j * dphi), np.deg2rad(P0 + (j + 1.0) * dphi)
        L0, L1 = self.global_longitude or self.longitude
        dlam = np.deg2rad((L1 - L0) / (self.global_Nx or self.Nx))
        return (radius ** 2 * dlam * (np.sin(north) - np.sin(south))).astype(self.dtype)

    def slab(self, rank, world_size):
        """Longitude slab of this grid owned by `rank`: Nx/P columns × full Ny."""
        assert self.Nx % world_size == 0, "Nx must divide evenly into longitude slabs"
        nx = self.Nx // world_size
        L0, L1 = self.longitude
        dl = (L1 - L0) / self.Nx
        g = LatitudeLongitudeGrid((nx, self.Ny, self.Nz), (L0 + rank * nx * dl, L0 + (rank + 1) * nx * dl),
                                  self.latitude, self.z, self.halo, self.dtype)
        g.i_offset = rank * nx
        g.global_Nx = self.Nx
        g.global_longitude = self.longitude
        return g

    i_offset = 0
    global_Nx = None
    global_longitude = None


def fractional_indices(grid, source_Nx, source_Ny, ring=1, source_longitude=(0.0, 360.0),
                       source_latitude=(-90.0, 90.0)):
    """Construction-time regridding weights: zero-based fractional indices of
    every cell of the ring-extended ocean surface in a regular lat-lon source grid.  λ is wrapped
    into the source's periodic range so that fi ≥ 0.  Returns numpy parents (1, Ny+2r, Nx+2r)."""
    if hasattr(grid, "fold"):                       # curvilinear (tripolar) grid: 2-D coordinates
        return grid.fractional_indices(source_Nx, source_Ny, ring, source_longitude, source_latitude)
    r = ring
    i = np.arange(-r, grid.Nx + r)
    j = np.arange(-r, grid.Ny + r)
    dl = (source_longitude[1] - source_longitude[0]) / source_Nx
    dp = (source_latitude[1] - source_latitude[0]) / source_Ny
    lam0 = source_longitude[0] + 0.5 * dl
    phi0 = source_latitude[0] + 0.5 * dp
    lam = grid.lambda_centers(i)
    lam = lam0 + np.mod(lam - lam0, source_longitude[1] - source_longitude[0])
    fi = (lam - lam0) / dl
    fj = (grid.phi_centers(j) - phi0) / dp
    FI = np.broadcast_to(fi[None, :], (j.size, i.size)).astype(grid.dtype)[None].copy()
    FJ = np.broadcast_to(fj[:, None], (j.size, i.size)).astype(grid.dtype)[None].copy()
    return FI, FJ


class TripolarGrid:
    """Tripolar grid in the sense of Murray (1996), as `TripolarGrid` of the reference's configurations
    (/root/reference/src/OceanConfigurations/one_degree_tripolar.jl:20-73, examples/one_degree_tripolar_ocean_sea_ice.jl:17-42;
    BASELINE.json config 5): regular latitude–longitude rows south of `north_poles_latitude`, and north of it a conformal
    bipolar cap with its two singularities on that latitude circle (on land in the real grids), so that the top row of the
    index space runs along the line between the two poles and FOLDS onto itself: cell (i, Ny−1) touches cell (Nx−1−i, Ny−1).

    Construction (host-side set-up only; nothing here runs per step): the cap is mapped stereographically to the unit disk,
    z = r e^{iλ'}, r = tan((90°−φ)/2)/tan((90°−φ₀)/2); w = 2 atanh z maps the disk conformally onto the strip |Im w| < π/2 with
    the poles z = ±1 at w = ±∞.  Grid lines are Re w = const (i) and Im w = const (j): on the boundary circle Re w =
    ln|cot(λ'/2)|, which ties column i to the longitude of the regular rows below; Im w runs from ±π/2 at φ₀ to 0 on the fold.
    What the flux path needs from the grid: λ, φ of every cell of the ring-extended surface (→ fractional source indices),
    the angle θ between the i-direction and east (→ cos θ, sin θ: the prescribed winds are rotated into the grid frame), and
    the fold rule for north halos."""

    def __init__(self, size, south_latitude=-80.0, north_poles_latitude=55.0, first_pole_longitude=70.0, z=(-5000.0, 0.0),
                 halo=(7, 7, 7), dtype=np.float64, cap_rows=None):
        self.Nx, self.Ny, self.Nz = (int(s) for s in size)
        assert self.Nx % 2 == 0, "the fold pairs column i with column Nx-1-i"
        self.south, self.phi0, self.lam_pole = float(south_latitude), float(north_poles_latitude), float(first_pole_longitude)
        self.z, self.halo, self.dtype = z, tuple(halo), np.dtype(dtype)
        # rows in the cap: same meridional spacing as below, measured along the boundary-to-fold arc through the geographic pole
        if cap_rows is None:
            dphi = (90.0 - self.south) / self.Ny
            cap_rows = max(2, int(round((90.0 - self.phi0) / dphi)))
        self.j0 = self.Ny - int(cap_rows)          # first cap row
        self.longitude, self.latitude = (0.0, 360.0), (self.south, 90.0)

    i_offset = 0
    global_Nx = None

    @property
    def size(self):
        return (self.Nx, self.Ny, self.Nz)

    def dz(self):
        Hz = self.halo[2]
        return np.full(self.Nz + 2 * Hz, (self.z[1] - self.z[0]) / self.Nz, dtype=self.dtype)

    def fold(self, i, j):
        """Interior (i, j) that a (possibly halo) index pair refers to: periodic in i, folded at the north, clamped at the south."""
        i = np.mod(np.asarray(i), self.Nx)
        j = np.asarray(j)
        over = j >= self.Ny
        jf = np.where(over, 2 * self.Ny - 1 - j, j)
        i_f = np.where(over, self.Nx - 1 - i, i)
        return i_f, np.clip(jf, 0, self.Ny - 1), over

    def coordinates(self, i, j):
        """λ, φ (degrees) of the cell centres at index arrays i, j (broadcast against each other; halos allowed)."""
        i = np.asarray(i, dtype=np.float64)
        j = np.asarray(j, dtype=np.float64)
        I, J = np.broadcast_arrays(i, j)
        ii, jj, _ = self.fold(I.astype(np.int64), J.astype(np.int64))
        ii = ii.astype(np.float64)
        jj = jj.astype(np.float64)
        lam = (ii + 0.5) * (360.0 / self.Nx)
        dphi = (self.phi0 - self.south) / self.j0
        phi = self.south + (np.minimum(jj, self.j0 - 1) + 0.5) * dphi
        cap = jj >= self.j0
        if np.any(cap):
            lp = np.deg2rad(lam[cap])                                   # λ' of the column on the boundary circle
            upper = lp < np.pi
            s = np.log(np.abs(1.0 / np.tan(lp / 2.0)))                  # Re w
            frac = (jj[cap] - self.j0 + 0.5) / (self.Ny - self.j0)      # 0 at φ₀ → 1 on the fold
            t = (np.pi / 2.0) * (1.0 - frac) * np.where(upper, 1.0, -1.0)   # Im w
            zc = np.tanh(0.5 * (s + 1j * t))
            r = np.abs(zc)
            phi[cap] = 90.0 - 2.0 * np.rad2deg(np.arctan(r * np.tan(np.deg2rad(90.0 - self.phi0) / 2.0)))
            lam[cap] = np.rad2deg(np.angle(zc)) % 360.0
        return (lam + self.lam_pole) % 360.0, phi

    def rotation(self, ring=1):
        """cos θ, sin θ parents (1, Ny+2r, Nx+2r): θ is the angle from geographic east to the grid's i-direction, from centred
        differences of the cell-centre coordinates.  A prescribed (eastward, northward) wind becomes
        u_i = u cos θ + v sin θ, v_j = −u sin θ + v cos θ (coflux_atmos_series.cos_theta / sin_theta)."""
        r = ring
        jj, ii = np.meshgrid(np.arange(-r, self.Ny + r), np.arange(-r, self.Nx + r), indexing="ij")
        # neighbours along +i / −i in INDEX space of the (folded) interior cell, so that halo cells carry their partner's frame
        i_f, j_f, over = self.fold(ii, jj)
        lam_e, phi_e = self.coordinates(i_f + 1, j_f)
        lam_w, phi_w = self.coordinates(i_f - 1, j_f)
        _, phi_c = self.coordinates(i_f, j_f)
        dlam = (lam_e - lam_w + 540.0) % 360.0 - 180.0
        dx = np.deg2rad(dlam) * np.cos(np.deg2rad(phi_c))
        dy = np.deg2rad(phi_e - phi_w)
        theta = np.arctan2(dy, dx)
        theta = np.where(over, theta + np.pi, theta)                 # across the fold the i-direction is reversed
        return np.cos(theta).astype(self.dtype)[None].copy(), np.sin(theta).astype(self.dtype)[None].copy()

    def fractional_indices(self, source_Nx, source_Ny, ring=1, source_longitude=(0.0, 360.0), source_latitude=(-90.0, 90.0)):
        """Construction-time regridding weights for this curvilinear grid: 2-D fractional indices."""
        r = ring
        jj, ii = np.meshgrid(np.arange(-r, self.Ny + r), np.arange(-r, self.Nx + r), indexing="ij")
        lam, phi = self.coordinates(ii, jj)
        dl = (source_longitude[1] - source_longitude[0]) / source_Nx
        dp = (source_latitude[1] - source_latitude[0]) / source_Ny
        lam0 = source_longitude[0] + 0.5 * dl
        phi0 = source_latitude[0] + 0.5 * dp
        lamw = lam0 + np.mod(lam - lam0, source_longitude[1] - source_longitude[0])
        fi = (lamw - lam0) / dl
        fj = np.clip((phi - phi0) / dp, 0.0, source_Ny - 1.0)        # the cap reaches the pole: stay inside the source rows
        return fi.astype(self.dtype)[None].copy(), fj.astype(self.dtype)[None].copy()

    def fill_north_fold(self, parent, sign=1.0):
        """Fill the north halo rows of a (nk, Ny+2Hy, Nx+2Hx) parent by the fold rule (cell-centred partner; sign = −1 for
        the components of a vector, whose axes are reversed across the fold), and the east / west halos periodically."""
        Hx, Hy = self.halo[0], self.halo[1]
        Nx, Ny = self.Nx, self.Ny
        a = parent
        a[:, :, :Hx] = a[:, :, Nx:Nx + Hx]
        a[:, :, Nx + Hx:] = a[:, :, Hx:2 * Hx]
        for m in range(Hy):
            src = a[:, Hy + Ny - 1 - m, Hx:Hx + Nx][:, ::-1]
            a[:, Hy + Ny + m, Hx:Hx + Nx] = sign * src
        a[:, Hy + Ny:, :Hx] = a[:, Hy + Ny:, Nx:Nx + Hx]
        a[:, Hy + Ny:, Nx + Hx:] = a[:, Hy + Ny:, Hx:2 * Hx]
        return a

    def horizontal_areas(self, radius=6371e3):
        """Az(j) of the regular rows (the cap rows reuse the last regular row's value: the salinity normaliser of a tripolar
        run needs the true 2-D metric, which belongs to the host ocean model)."""
        Hy = self.halo[1]
        dphi = (self.phi0 - self.south) / self.j0
        j = np.minimum(np.arange(-Hy, self.Ny + Hy, dtype=np.float64), self.j0 - 1)
        south, north = np.deg2rad(self.south + j * dphi), np.deg2rad(self.south + (j + 1.0) * dphi)
        return (radius ** 2 * np.deg2rad(360.0 / self.Nx) * (np.sin(north) - np.sin(south))).astype(self.dtype)

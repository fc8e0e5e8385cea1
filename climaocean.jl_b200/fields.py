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
        L0, L1 = self.longitude
        return L0 + (np.asarray(i, dtype=np.float64) + 0.5) * ((L1 - L0) / self.Nx)

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
        dlam = np.deg2rad((self.longitude[1] - self.longitude[0]) / self.Nx)
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
        return g

    i_offset = 0
    global_Nx = None


def fractional_indices(grid, source_Nx, source_Ny, ring=1, source_longitude=(0.0, 360.0),
                       source_latitude=(-90.0, 90.0)):
    """Construction-time regridding weights: zero-based fractional indices of
    every cell of the ring-extended ocean surface in a regular lat-lon source grid.  λ is wrapped
    into the source's periodic range so that fi ≥ 0.  Returns numpy parents (1, Ny+2r, Nx+2r)."""
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

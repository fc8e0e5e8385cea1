"""All arrays the surface-flux path touches on one (local) grid, as Fields on the host (numpy) or on
a CUDA device (torch), plus builders of the ctypes bundles of include/coflux.h.

Plumbing only: allocation, layout, descriptors.  No flux arithmetic happens in Python.
"""
import ctypes as C

import numpy as np

from . import _abi, synth
from .fields import Field, FieldTimeSeries, arr, fractional_indices

ATM_NAMES = ("u", "v", "T", "q", "p", "Qs", "Ql", "rain", "snow")
XCH_NAMES = ("u", "v", "T", "p", "q", "Qs", "Ql", "Mp")
IFACE_NAMES = ("latent_heat", "sensible_heat", "water_vapor", "x_momentum", "y_momentum", "interface_temperature",
               "friction_velocity", "temperature_scale", "humidity_scale")
NET_NAMES = ("u", "v", "T", "S", "upwelling_longwave", "downwelling_longwave", "downwelling_shortwave",
             "penetrating_shortwave")
IO_NAMES = ("frazil_heat", "interface_heat", "salt", "x_momentum", "y_momentum")
ICE_NAMES = ("thickness", "previous_thickness", "concentration", "salinity", "u", "v", "top_temperature", "snow_thickness")
NET_ICE_NAMES = ("top_heat", "bottom_heat", "top_u", "top_v")
AVG_NAMES = ("tau_x", "tau_y", "JT", "JS", "Qc", "Qv", "JT_atmosphere_ocean", "JT_ice_ocean", "JS_ice_ocean", "JT_frazil")
LAND_NAMES = ("rivers", "icebergs")


class SurfaceFluxData:
    def __init__(self, grid, device=None):
        self.grid = grid
        self.device = device
        self.dtype = grid.dtype
        self.atmos = {}          # name -> FieldTimeSeries
        self.times = None
        self.time_indexing = _abi.TIME_LINEAR
        self.fi = self.fj = None
        self.ocean = {}          # u v T S (3-D)
        self.dz = None           # 1-D Field (nk,1,1)
        self.area = None         # 1-D Field (1,nj,1): horizontal cell areas Az(j) of the latitude–longitude grid
        self.mask = None
        self.ice = None          # dict or None
        self.exchange = {}
        self.ao = {}
        self.ai = {}
        self.io = {}
        self.net = {}
        self.iterations = None
        self.land = None         # dict name -> FieldTimeSeries (rivers, icebergs) or None
        self.land_times = None
        self.land_time_indexing = _abi.TIME_LINEAR
        self.lfi = self.lfj = None
        self.rotation = None     # (cos θ, sin θ) Fields of a curvilinear grid, or None
        self.net_ice = {}
        self.averages = {}
        self.ring_start = self.ring_capacity = 0      # device ring buffer of the atmosphere series (forcing.DeviceForcingWindow)
        self._keep = []

    # ------------------------------------------------------------------------------------------
    @classmethod
    def synthetic(cls, grid, device=None, with_ice=False, land_fraction=0.0, frazil=False, Nt=8, atmos_size=(640, 320),
                  atmos_halo=3, ring=1, with_land=False, land_size=(360, 180), land_Nt=4):
        """Inputs per SURVEY §8d on the host; call .to(device) for the CUDA copy."""
        self = cls(grid, None)
        dt = grid.dtype
        series, times = synth.atmosphere_series(atmos_size[0], atmos_size[1], Nt, atmos_halo, dt)
        self.times = times
        for n in ATM_NAMES:
            self.atmos[n] = FieldTimeSeries(series[n], (atmos_halo, atmos_halo, 0), times, n)
        FI, FJ = fractional_indices(grid, atmos_size[0], atmos_size[1], ring=ring)
        self.fi = Field(FI, (ring, ring, 0), "fi")
        self.fj = Field(FJ, (ring, ring, 0), "fj")
        oc = synth.ocean_state(grid, dt, frazil=frazil)
        for n in ("u", "v", "T", "S"):
            self.ocean[n] = Field(oc[n], grid.halo, "ocean_" + n)
        self.dz = Field(grid.dz().reshape(-1, 1, 1).copy(), (0, 0, grid.halo[2]), "dz")
        self.area = Field(grid.horizontal_areas().reshape(1, -1, 1).copy(), (0, grid.halo[1], 0), "Az")
        if land_fraction > 0:
            self.mask = Field(synth.land_mask(grid, land_fraction), (grid.halo[0], grid.halo[1], 0), "mask")
        if with_ice:
            ice = synth.sea_ice_state(grid, dt)
            self.ice = {n: Field(ice[n], (grid.halo[0], grid.halo[1], 0), "ice_" + n) for n in ICE_NAMES}
        if hasattr(grid, "fold"):
            # curvilinear grid: winds are rotated into the grid frame; north halos of the model state follow the fold
            cs, sn = grid.rotation(ring)
            self.rotation = (Field(cs, (ring, ring, 0), "cos_theta"), Field(sn, (ring, ring, 0), "sin_theta"))
            for n in ("u", "v", "T", "S"):
                grid.fill_north_fold(self.ocean[n].data, -1.0 if n in ("u", "v") else 1.0)
            if self.ice is not None:
                for n, f in self.ice.items():
                    grid.fill_north_fold(f.data, -1.0 if n in ("u", "v") else 1.0)
            if self.mask is not None:
                grid.fill_north_fold(self.mask.data)
        if with_land:
            lh = 2
            lser, ltimes = synth.land_series(land_size[0], land_size[1], land_Nt, lh, dt)
            self.land_times = ltimes
            self.land = {n: FieldTimeSeries(lser[n], (lh, lh, 0), ltimes, n) for n in LAND_NAMES}
            LFI, LFJ = fractional_indices(grid, land_size[0], land_size[1], ring=ring)
            self.lfi, self.lfj = Field(LFI, (ring, ring, 0), "land_fi"), Field(LFJ, (ring, ring, 0), "land_fj")
        self.allocate_outputs()
        return self

    def allocate_outputs(self):
        g = self.grid
        h2 = (g.halo[0], g.halo[1], 0)
        size2 = (g.Nx, g.Ny, 1)
        mk = lambda name: Field.zeros(size2, h2, self.dtype, self.device, name)
        self.exchange = {n: mk("exchange_" + n) for n in XCH_NAMES}
        self.ao = {n: mk("ao_" + n) for n in IFACE_NAMES}
        self.net = {n: mk("net_" + n) for n in NET_NAMES}
        self.iterations = Field.zeros(size2, h2, np.int32, self.device, "iterations")
        if self.ice is not None:
            self.iterations_ai = Field.zeros(size2, h2, np.int32, self.device, "iterations_ai")
            self.ai = {n: mk("ai_" + n) for n in IFACE_NAMES}
            self.io = {n: mk("io_" + n) for n in IO_NAMES}
            self.net_ice = {n: mk("net_ice_" + n) for n in NET_ICE_NAMES}

    def allocate_averages(self):
        """Zeroed running-average Fields (omip_diagnostics.jl:125-158), one per averaged flux."""
        g = self.grid
        self.averages = {n: Field.zeros((g.Nx, g.Ny, 1), (g.halo[0], g.halo[1], 0), self.dtype, self.device, "avg_" + n)
                         for n in AVG_NAMES}
        return self.averages

    def to(self, device):
        o = SurfaceFluxData(self.grid, device)
        o.times, o.time_indexing = self.times, self.time_indexing
        o.atmos = {n: f.to(device) for n, f in self.atmos.items()}
        o.fi, o.fj = self.fi.to(device), self.fj.to(device)
        o.ocean = {n: f.to(device) for n, f in self.ocean.items()}
        o.dz = self.dz.to(device)
        o.area = self.area.to(device) if self.area is not None else None
        o.mask = self.mask.to(device) if self.mask is not None else None
        o.ice = {n: f.to(device) for n, f in self.ice.items()} if self.ice is not None else None
        if self.land is not None:
            o.land = {n: f.to(device) for n, f in self.land.items()}
            o.land_times, o.land_time_indexing = self.land_times, self.land_time_indexing
            o.lfi, o.lfj = self.lfi.to(device), self.lfj.to(device)
        if self.rotation is not None:
            o.rotation = tuple(f.to(device) for f in self.rotation)
        o.allocate_outputs()
        if getattr(self, "eos", None):
            o.eos = {k: f.to(device) for k, f in self.eos.items()}
        return o

    def to_device_columns(self, device, Nz, Hz=7, fill_columns=False):
        """Device copy whose ocean u, v, T, S are full 3-D parents (Nz+2Hz levels) although this host
        object only holds the surface plane (grid.Nz == 1): the plane is placed at k = Nz-1 and, when
        fill_columns is set, T and S are extended downwards on the device (T relaxing to −1 °C with
        bands of super-cooled water for the frazil sweep).  Used by bench.py for the 1/12° grid, where
        generating 75 levels on the host would only cost time — the flux solve reads k = Nz-1 only."""
        import torch
        from .fields import LatitudeLongitudeGrid
        g0 = self.grid
        assert g0.Nz == 1 and g0.halo[2] == 0
        g = LatitudeLongitudeGrid((g0.Nx, g0.Ny, Nz), g0.longitude, g0.latitude, g0.z, (g0.halo[0], g0.halo[1], Hz), g0.dtype)
        g.i_offset, g.global_Nx, g.global_longitude = g0.i_offset, g0.global_Nx, g0.global_longitude
        o = self.to(device)
        o.grid = g
        tdt = torch.float64 if np.dtype(self.dtype) == np.float64 else torch.float32
        nk = Nz + 2 * Hz
        for n in ("u", "v", "T", "S"):
            plane = o.ocean[n].data[0]
            full = torch.zeros((nk,) + tuple(plane.shape), dtype=tdt, device=device)
            if fill_columns and n in ("T", "S"):
                k = torch.arange(nk, device=device, dtype=tdt).view(-1, 1, 1)
                depth = ((Nz - 1 + Hz) - k).clamp(min=0) / max(Nz - 1, 1)
                if n == "T":
                    full[...] = plane[None] - depth * (plane[None] + 1.0) * 0.9
                    jj = torch.arange(plane.shape[0], device=device).view(1, -1, 1)
                    cold = ((jj + 3 * k.long()) % 16) >= 14
                    full[...] = torch.where(cold, torch.full_like(full, -2.6) - 0.2 * depth, full)
                else:
                    full[...] = plane[None] + 0.5 * depth
            else:
                full[Nz - 1 + Hz] = plane
            o.ocean[n] = Field(full, (g0.halo[0], g0.halo[1], Hz), "ocean_" + n)
        o.dz = Field.from_numpy(g.dz().reshape(-1, 1, 1).copy(), (0, 0, Hz), device, "dz")
        return o

    # ------------------------------------------------------------------------------------------
    # ctypes bundles (the returned struct keeps python references alive through self._keep)
    # ------------------------------------------------------------------------------------------
    def atmos_series(self):
        s = _abi.AtmosSeries()
        for n in ATM_NAMES:
            setattr(s, n, arr(self.atmos.get(n)))
        self._times_c = (C.c_double * len(self.times))(*self.times)
        s.times = C.cast(self._times_c, C.POINTER(C.c_double))
        s.Nt = len(self.times)
        s.time_indexing = self.time_indexing
        s.cycle_period = 0.0
        s.fi, s.fj = arr(self.fi), arr(self.fj)
        if self.rotation is not None:
            s.cos_theta, s.sin_theta = arr(self.rotation[0]), arr(self.rotation[1])
        else:
            s.cos_theta, s.sin_theta = arr(None), arr(None)
        s.ring_start, s.ring_capacity = int(self.ring_start), int(self.ring_capacity)
        return s

    def land_series(self):
        """coflux_land_series (JRA55PrescribedLand: friver + licalvf, atmosphere.jl:46), or None without land."""
        if self.land is None:
            return None
        s = _abi.LandSeries()
        s.rivers, s.icebergs = arr(self.land.get("rivers")), arr(self.land.get("icebergs"))
        self._land_times_c = (C.c_double * len(self.land_times))(*self.land_times)
        s.times = C.cast(self._land_times_c, C.POINTER(C.c_double))
        s.Nt = len(self.land_times)
        s.time_indexing = self.land_time_indexing
        s.cycle_period = 0.0
        s.fi, s.fj = arr(self.lfi), arr(self.lfj)
        s.ring_start = s.ring_capacity = 0
        return s

    def exchange_state(self):
        s = _abi.ExchangeState()
        for n in XCH_NAMES:
            setattr(s, n, arr(self.exchange[n]))
        return s

    def ocean_surface(self):
        s = _abi.OceanSurface()
        for n in ("u", "v", "T", "S"):
            setattr(s, n, arr(self.ocean[n]))
        s.mask = arr(self.mask)
        return s

    def interface_fluxes(self, which="ao", with_iterations=True):
        d = self.ao if which == "ao" else self.ai
        s = _abi.InterfaceFluxes()
        for n in IFACE_NAMES:
            setattr(s, n, arr(d[n]))
        s.iterations = arr(self.iterations if which == "ao" else self.iterations_ai) if with_iterations else arr(None)
        return s

    def sea_ice_state(self):
        s = _abi.SeaIceState()
        for n in ICE_NAMES:
            setattr(s, n, arr(self.ice.get(n)))
        s.albedo = arr(self.ice.get("albedo"))
        return s

    def net_sea_ice_fluxes(self, with_stress=True):
        s = _abi.NetSeaIceFluxes()
        s.top_heat, s.bottom_heat = arr(self.net_ice["top_heat"]), arr(self.net_ice["bottom_heat"])
        s.top_u = arr(self.net_ice["top_u"]) if with_stress else arr(None)
        s.top_v = arr(self.net_ice["top_v"]) if with_stress else arr(None)
        return s

    def flux_averages(self, previous_interval, dt, names=AVG_NAMES):
        """coflux_flux_averages over self.averages (allocate_averages() first)."""
        s = _abi.FluxAverages()
        for n in AVG_NAMES:
            setattr(s, n, arr(self.averages[n]) if n in names else arr(None))
        s.previous_interval, s.dt = float(previous_interval), float(dt)
        return s

    def ocean_columns(self):
        s = _abi.OceanColumns()
        s.T, s.S = arr(self.ocean["T"]), arr(self.ocean["S"])
        dz = self.dz.array()
        dz.stride_i = 0
        dz.stride_j = 0
        dz.stride_k = 1
        s.dz = dz
        s.u, s.v = arr(self.ocean["u"]), arr(self.ocean["v"])
        return s

    def ice_ocean_fluxes(self):
        s = _abi.IceOceanFluxes()
        for n in IO_NAMES:
            setattr(s, n, arr(self.io[n]))
        return s

    def net_ocean_fluxes(self):
        s = _abi.NetOceanFluxes()
        for n in NET_NAMES:
            setattr(s, n, arr(self.net[n]))
        return s

    def closure_forcing(self, minimum_friction_velocity=1e-6, minimum_surface_tke=1e-4, Cb=3.75, gravitational_acceleration=9.80665,
                        with_buoyancy=True):
        """coflux_closure_forcing with freshly allocated output Fields (self.closure) and synthetic α, β of the surface cell
        (defaults: kpp_parameters.jl:98, nemo_tke_parameters.jl:45,54)."""
        g = self.grid
        h2, size2 = (g.halo[0], g.halo[1], 0), (g.Nx, g.Ny, 1)
        if not getattr(self, "closure", None):
            self.closure = {n: Field.zeros(size2, h2, self.dtype, self.device, "closure_" + n)
                            for n in ("friction_velocity", "friction_velocity_squared", "surface_tke", "buoyancy_flux")}
        if with_buoyancy and not getattr(self, "eos", None):
            T = self.ocean["T"].numpy()[-1 - self.ocean["T"].halo[2]] if self.ocean["T"].data.shape[0] > 1 else self.ocean["T"].numpy()[0]
            # smooth stand-ins for α(T), β: the EOS itself belongs to the host ocean model
            alpha = (5e-5 + 1.1e-5 * np.clip(T, -2.0, 32.0)).astype(self.dtype)[None]
            beta = np.full_like(alpha, 7.6e-4)
            self.eos = {"alpha": Field.from_numpy(alpha, (self.ocean["T"].halo[0], self.ocean["T"].halo[1], 0), self.device, "alpha"),
                        "beta": Field.from_numpy(beta, (self.ocean["T"].halo[0], self.ocean["T"].halo[1], 0), self.device, "beta")}
        f = _abi.ClosureForcing()
        f.thermal_expansion = arr(self.eos["alpha"]) if with_buoyancy else arr(None)
        f.haline_contraction = arr(self.eos["beta"]) if with_buoyancy else arr(None)
        f.friction_velocity = arr(self.closure["friction_velocity"])
        f.friction_velocity_squared = arr(self.closure["friction_velocity_squared"])
        f.surface_tke = arr(self.closure["surface_tke"])
        f.buoyancy_flux = arr(self.closure["buoyancy_flux"]) if with_buoyancy else arr(None)
        f.minimum_friction_velocity, f.minimum_surface_tke = minimum_friction_velocity, minimum_surface_tke
        f.Cb, f.gravitational_acceleration = Cb, gravitational_acceleration
        return f

    def salinity_normalization(self, additional=None):
        """coflux_salinity_normalization for NormalizeSalinity (omip_simulation.jl:187-220): the bulk salinity flux
        net.S, an optional materialised additional flux, Az(j) (stride_i = 0) and the wet mask."""
        s = _abi.SalinityNormalization()
        s.flux = arr(self.net["S"])
        s.additional = arr(additional)
        a = self.area.array()
        a.stride_i = 0
        a.stride_j = 1
        a.stride_k = 0
        a.off_i = 0
        s.area = a
        s.mask = arr(self.mask)
        self._keep_norm = (s, additional)
        return s

    def update_bundles(self, with_ice_terms=False, with_land=True):
        """(UpdateInputs, UpdateOutputs) for coflux_update_state; keeps the sub-structs alive."""
        a, o = self.atmos_series(), self.ocean_surface()
        x, f, n = self.exchange_state(), self.interface_fluxes("ao"), self.net_ocean_fluxes()
        inp = _abi.UpdateInputs(C.pointer(a), C.pointer(o), None, None, None)
        keep = [a, o, x, f, n]
        land = self.land_series() if with_land else None
        if land is not None:
            inp.land = C.pointer(land)
            keep.append(land)
        if with_ice_terms and self.ice is not None:
            ice, io = self.sea_ice_state(), self.ice_ocean_fluxes()
            inp.sea_ice, inp.ice_ocean = C.pointer(ice), C.pointer(io)
            keep += [ice, io]
        out = _abi.UpdateOutputs(C.pointer(x), C.pointer(f), C.pointer(n))
        self._keep = keep
        return inp, out

    def outputs(self):
        """name -> numpy interior (Ny, Nx) of every output field (copied to host)."""
        res = {}
        for grp, d in (("exchange", self.exchange), ("ao", self.ao), ("ai", self.ai), ("io", self.io), ("net", self.net),
                       ("net_ice", self.net_ice), ("avg", self.averages)):
            for n, f in d.items():
                a = f.numpy()
                Hx, Hy, _ = f.halo
                res[f"{grp}.{n}"] = a[0, Hy:a.shape[1] - Hy, Hx:a.shape[2] - Hx].copy()
        return res

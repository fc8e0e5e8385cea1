"""Thin object wrapper over the C ABI (one `coflux_ctx` per device/grid/parameter set).

Every method enqueues CUDA work through libcoflux.so on the given stream; there is no Python
or CPU implementation of any of these operations."""
import ctypes as C

from . import _abi


def _stream_handle(stream):
    if stream is None:
        return None
    if isinstance(stream, int):
        return C.c_void_p(stream)
    return C.c_void_p(stream.cuda_stream)  # torch.cuda.Stream


class Engine:
    def __init__(self, cfg):
        self.lib = _abi.load_library()
        self.cfg = cfg
        self._ctx = C.c_void_p()
        _abi.check(self.lib.coflux_create(C.byref(self._ctx), C.byref(cfg)), self.lib)

    def close(self):
        if self._ctx:
            self.lib.coflux_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launches(self):
        n = C.c_int64()
        _abi.check(self.lib.coflux_launch_count(self._ctx, C.byref(n)), self.lib)
        return n.value

    def profile(self, enable=True):
        _abi.check(self.lib.coflux_profile_enable(self._ctx, 1 if enable else 0), self.lib)

    def profile_read(self):
        """(flux_kernel_ms, stress_kernel_ms, calls) accumulated since the last read."""
        a, b, n = C.c_double(), C.c_double(), C.c_int64()
        _abi.check(self.lib.coflux_profile_read(self._ctx, C.byref(a), C.byref(b), C.byref(n)), self.lib)
        return a.value, b.value, n.value

    # --- multi-GPU seam (mode B) ---
    def seam_export(self):
        buf = C.create_string_buffer(_abi.SEAM_HANDLE_BYTES)
        _abi.check(self.lib.coflux_seam_export(self._ctx, buf), self.lib)
        return bytes(buf.raw)

    def seam_attach(self, west_handle, east_handle, rank, world):
        w = C.create_string_buffer(west_handle, _abi.SEAM_HANDLE_BYTES)
        e = C.create_string_buffer(east_handle, _abi.SEAM_HANDLE_BYTES)
        _abi.check(self.lib.coflux_seam_attach(self._ctx, w, e, int(rank), int(world)), self.lib)

    def seam_detach(self):
        _abi.check(self.lib.coflux_seam_detach(self._ctx), self.lib)

    # --- entry points (names follow the reference's generic functions) ---
    def interpolate_atmosphere_state(self, series, time, exchange, stream=None):
        _abi.check(self.lib.coflux_interpolate_atmosphere(self._ctx, C.byref(series), float(time), C.byref(exchange),
                                                          _stream_handle(stream)), self.lib)

    def compute_atmosphere_ocean_fluxes(self, exchange, ocean, fluxes, stream=None):
        _abi.check(self.lib.coflux_atmosphere_ocean_fluxes(self._ctx, C.byref(exchange), C.byref(ocean),
                                                           C.byref(fluxes), _stream_handle(stream)), self.lib)

    def compute_atmosphere_sea_ice_fluxes(self, exchange, ocean, ice, fluxes, stream=None):
        _abi.check(self.lib.coflux_atmosphere_sea_ice_fluxes(self._ctx, C.byref(exchange), C.byref(ocean),
                                                             C.byref(ice), C.byref(fluxes), _stream_handle(stream)),
                   self.lib)

    def compute_sea_ice_ocean_fluxes(self, columns, ice, dt, fluxes, stream=None):
        _abi.check(self.lib.coflux_sea_ice_ocean_fluxes(self._ctx, C.byref(columns), C.byref(ice), float(dt),
                                                        C.byref(fluxes), _stream_handle(stream)), self.lib)

    def compute_net_ocean_fluxes(self, exchange, ocean, ao, ice, io, net, stream=None):
        _abi.check(self.lib.coflux_assemble_net_ocean_fluxes(
            self._ctx, C.byref(exchange), C.byref(ocean), C.byref(ao), C.byref(ice) if ice is not None else None,
            C.byref(io) if io is not None else None, C.byref(net), _stream_handle(stream)), self.lib)

    def update_state(self, inputs, outputs, time, stream=None):
        _abi.check(self.lib.coflux_update_state(self._ctx, C.byref(inputs), C.byref(outputs), float(time),
                                                _stream_handle(stream)), self.lib)

    def interpolate_land(self, land, time, exchange, stream=None):
        """Exchange Mp += rivers + icebergs (JRA55PrescribedLand, atmosphere.jl:46)."""
        _abi.check(self.lib.coflux_interpolate_land(self._ctx, C.byref(land), float(time), C.byref(exchange), _stream_handle(stream)), self.lib)

    def compute_net_sea_ice_fluxes(self, exchange, ocean, ice, ai, io, out, stream=None):
        _abi.check(self.lib.coflux_assemble_net_sea_ice_fluxes(
            self._ctx, C.byref(exchange), C.byref(ocean) if ocean is not None else None, C.byref(ice), C.byref(ai),
            C.byref(io) if io is not None else None, C.byref(out), _stream_handle(stream)), self.lib)

    # --- time-averaged flux diagnostics (omip_diagnostics.jl:125-158) ---
    def attach_flux_averages(self, averages):
        """Every following update_state / compute_sea_ice_ocean_fluxes updates the running averages; None detaches."""
        self._avg_keep = averages
        _abi.check(self.lib.coflux_attach_flux_averages(self._ctx, C.byref(averages) if averages is not None else None), self.lib)

    def accumulate_flux_averages(self, net, ao, ice, io, averages, stream=None):
        _abi.check(self.lib.coflux_accumulate_flux_averages(
            self._ctx, C.byref(net), C.byref(ao) if ao is not None else None, C.byref(ice) if ice is not None else None,
            C.byref(io) if io is not None else None, C.byref(averages), _stream_handle(stream)), self.lib)

    # --- closure surface-forcing front ends (KPP/kpp_surface_forcing.jl, NEMOTKE/nemo_tke_surface_forcing.jl) ---
    def closure_surface_forcing(self, net, forcing, stream=None):
        _abi.check(self.lib.coflux_closure_surface_forcing(self._ctx, C.byref(net), C.byref(forcing), _stream_handle(stream)), self.lib)

    def attach_closure_forcing(self, forcing):
        """Every following update_state emits the closure by-products from the stress kernel; None detaches."""
        self._closure_keep = forcing
        _abi.check(self.lib.coflux_attach_closure_forcing(self._ctx, C.byref(forcing) if forcing is not None else None), self.lib)

    # --- NormalizeSalinity (omip_simulation.jl:187-220) ---
    def salinity_flux_sums(self, norm, device_sums_ptr, stream=None):
        _abi.check(self.lib.coflux_salinity_flux_sums(self._ctx, C.byref(norm), device_sums_ptr, _stream_handle(stream)), self.lib)

    def subtract_mean_flux(self, norm, device_sums_ptr, stream=None):
        _abi.check(self.lib.coflux_subtract_mean_flux(self._ctx, C.byref(norm), device_sums_ptr, _stream_handle(stream)), self.lib)

    def normalize_salinity_flux(self, norm, stream=None):
        _abi.check(self.lib.coflux_normalize_salinity_flux(self._ctx, C.byref(norm), _stream_handle(stream)), self.lib)

    def update_state_host(self, series, step, time):
        h2d, d2h = C.c_int64(), C.c_int64()
        _abi.check(self.lib.coflux_update_state_host(self._ctx, C.byref(series), C.byref(step), float(time),
                                                     C.byref(h2d), C.byref(d2h)), self.lib)
        return h2d.value, d2h.value

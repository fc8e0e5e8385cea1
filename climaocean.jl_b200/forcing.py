"""Forcing-side window management of the prescribed atmosphere: the host half of what feeds
`interpolate_atmosphere_state!` every coupling step.

The reference builds its JRA55 atmosphere with `time_indices_in_memory = backend_size` (`/root/reference/src/
OMIPConfigurations/atmosphere.jl:22-27`, `experiments/OMIPSimulations/scripts/launch.sh:86-87`): of the year-long series
only `backend_size` consecutive time levels live on the device; when the clock leaves them the window is re-based on the
current level and reloaded (Oceananigans' `InMemory(start, length)` backend), cyclically for repeat-year forcing.

`InMemoryWindow` is that mechanism for coflux: it owns the `length`-level FieldTimeSeries handed to the kernels (host
numpy for the oracle, device torch tensors for the CUDA library) and the matching `times`, and re-bases them from a
host-resident source (pinned memory in production; NetCDF reading and file staging are out of scope).
Invariant: interpolating through the window gives bit-identical results to interpolating with the whole series in
memory — the window never contains a re-timed copy of the bracketing levels except the single wrapped level
`times[0] + period` that the cyclical indexing of coflux_time_indices itself uses.  No flux arithmetic happens here.
"""
import math

import numpy as np

from . import _abi
from .fields import FieldTimeSeries


class InMemoryWindow:
    def __init__(self, source, times, length, halo, device=None, time_indexing=_abi.TIME_LINEAR, cycle_period=0.0):
        """source: name -> numpy array (Nt, nk, nj, ni) on the host; times: (Nt,) strictly increasing."""
        self.source = source
        self.times = np.ascontiguousarray(times, dtype=np.float64)
        self.Nt = int(self.times.size)
        assert self.Nt >= 2 and np.all(np.diff(self.times) > 0), "series times must be strictly increasing"
        self.length = int(max(2, min(length, self.Nt + (1 if time_indexing == _abi.TIME_CYCLICAL else 0))))
        self.halo, self.device = tuple(halo), device
        self.mode = time_indexing
        self.period = float(cycle_period) if cycle_period and cycle_period > 0 else \
            float(self.times[-1] - self.times[0] + (self.times[-1] - self.times[-2]))
        self.start = None
        self.series = {}            # name -> FieldTimeSeries holding the window
        self.window_times = None
        self.reloads = 0
        self.bytes_loaded = 0

    # -- the global bracket of `time`, exactly as coflux_time_indices computes it -------------------------------------
    def _global_bracket(self, time):
        t, T, ts = float(time), self.period, self.times
        if self.mode == _abi.TIME_CYCLICAL:
            rel = math.fmod(t - ts[0], T)
            if rel < 0.0:
                rel += T
            t = float(ts[0] + rel)
            if t >= ts[-1]:
                return self.Nt - 1, t
        lo, hi = 0, self.Nt - 2
        while lo < hi:
            mid = (lo + hi + 1) // 2
            if ts[mid] <= t:
                lo = mid
            else:
                hi = mid - 1
        return lo, t

    def _load(self, start):
        Nt, L = self.Nt, self.length
        if self.mode == _abi.TIME_CYCLICAL:
            g = start + np.arange(L)
        else:
            start = min(start, Nt - 2)
            g = start + np.arange(min(L, Nt - start))
        idx, wraps = g % Nt, g // Nt
        self.window_times = self.times[idx] + wraps * self.period
        for name, arr in self.source.items():
            win = np.ascontiguousarray(arr[idx])
            self.series[name] = FieldTimeSeries.from_numpy(win, self.halo, self.window_times, self.device, name)
            self.bytes_loaded += win.nbytes
        self.start = int(start)
        self.reloads += 1

    def update(self, time):
        """Make sure the levels bracketing `time` are in memory.  Returns (effective_time, reloaded): pass effective_time
        and TIME_LINEAR (TIME_CLAMP if that was the mode) with `window_times` to the flux path."""
        n1, t = self._global_bracket(time)
        have = 0 if self.start is None else len(self.window_times)
        reloaded = False
        if self.start is None or n1 < self.start or n1 + 1 > self.start + have - 1:
            self._load(n1)
            reloaded = True
        return t, reloaded

    def apply(self, data, time):
        """Point a SurfaceFluxData at the window and return the time to hand to update_state / interpolate."""
        t, _ = self.update(time)
        for name, fts in self.series.items():
            data.atmos[name] = fts
        data.times = self.window_times
        data.time_indexing = _abi.TIME_CLAMP if self.mode == _abi.TIME_CLAMP else _abi.TIME_LINEAR
        return t


class DeviceForcingWindow:
    """`time_indices_in_memory` levels of every series in a DEVICE ring buffer, the next levels prefetched asynchronously
    (`prefetch = true`, /root/reference/src/OMIPConfigurations/atmosphere.jl:22-27; launch.sh:86-87) — SURVEY §8f row 2.

    Where `InMemoryWindow` re-bases and re-uploads the whole window synchronously when the clock leaves it, this class
    keeps a ring of `capacity` time slots per field on the device (coflux_forcing_window_*, include/coflux.h): a level is
    uploaded ONCE, from pinned host memory, on the window's own copy stream, `prefetch` levels ahead of the step that
    needs it; the compute stream is ordered behind the two levels a step reads by event waits, never by a host
    synchronisation.  The kernels read the ring in place through (ring_start, ring_capacity) of coflux_atmos_series.
    Invariant (tests/test_forcing_ring.py): bit-identical results to the whole series resident on the device."""

    def __init__(self, engine, source, times, capacity, halo, time_indexing=_abi.TIME_LINEAR, cycle_period=0.0, prefetch=1,
                 pin=True):
        import ctypes as C
        import torch
        self.engine, self.lib = engine, engine.lib
        self.names = list(source)
        self.times = np.ascontiguousarray(times, dtype=np.float64)
        self.Nt = int(self.times.size)
        assert self.Nt >= 2 and np.all(np.diff(self.times) > 0), "series times must be strictly increasing"
        self.mode = time_indexing
        self.period = float(cycle_period) if cycle_period and cycle_period > 0 else \
            float(self.times[-1] - self.times[0] + (self.times[-1] - self.times[-2]))
        self.capacity = int(capacity)
        self.prefetch = int(prefetch)
        assert self.capacity >= 2 + self.prefetch, "capacity must hold the two bracketing levels plus the prefetched ones"
        self.halo = tuple(halo)
        first = source[self.names[0]]
        self.shape = tuple(first.shape[1:])                       # (nk, nj, ni) of one time level
        self.plane = int(np.prod(self.shape))
        self.host = {}
        for n in self.names:                                       # pinned host copies: the source of every upload
            t = torch.from_numpy(np.ascontiguousarray(source[n]))
            self.host[n] = t.pin_memory() if pin else t
        self._w = C.c_void_p()
        _abi.check(self.lib.coflux_forcing_window_create(C.byref(self._w), engine._ctx, len(self.names), self.plane, self.capacity), self.lib)
        self.ptr = {}
        for k, n in enumerate(self.names):
            p = C.c_void_p()
            _abi.check(self.lib.coflux_forcing_window_field(self._w, k, C.byref(p)), self.lib)
            self.ptr[n] = p.value
        self.first = None          # lowest / highest logical (unwrapped) level uploaded so far
        self.last = None
        self._busy = None          # levels read by the step enqueued last

    def close(self):
        if self._w:
            self.lib.coflux_forcing_window_destroy(self._w)
            self._w = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # the bracket of `time`, as coflux_time_indices computes it, plus the cycle number (repeat-year forcing)
    def _bracket(self, time):
        t, T, ts = float(time), self.period, self.times
        cycle = 0
        if self.mode == _abi.TIME_CYCLICAL:
            cycle = int(math.floor((t - ts[0]) / T))
            rel = math.fmod(t - ts[0], T)
            if rel < 0.0:
                rel += T
            t = float(ts[0] + rel)
            if t >= ts[-1]:
                return self.Nt - 1, t, cycle
        lo, hi = 0, self.Nt - 2
        while lo < hi:
            mid = (lo + hi + 1) // 2
            if ts[mid] <= t:
                lo = mid
            else:
                hi = mid - 1
        return lo, t, cycle

    def _upload(self, G):
        import ctypes as C
        n = G % self.Nt
        planes = (C.c_void_p * len(self.names))(*[self.host[name][n].data_ptr() for name in self.names])
        _abi.check(self.lib.coflux_forcing_window_upload(self._w, G, planes), self.lib)

    def apply(self, data, time, stream=None):
        """Make the two levels bracketing `time` (and `prefetch` more) resident, order `stream` behind their uploads, point
        `data` at the ring.  Returns the time to hand to update_state.  Call `release(stream)` after the step is enqueued."""
        from .engine import _stream_handle
        n1, t, cycle = self._bracket(time)
        cyc = self.mode == _abi.TIME_CYCLICAL
        G1 = cycle * self.Nt + n1
        top = G1 + 1 + self.prefetch
        if not cyc:
            top = min(top, self.Nt - 1)
        if self.last is None or G1 > self.last or G1 < self.first:      # (re)start the ring at G1
            self.first, self.last = G1, G1 - 1
        for G in range(self.last + 1, top + 1):
            self._upload(G)
            self.last = G
        self.first = max(self.first, self.last - self.capacity + 1)
        G2 = G1 + 1 if (cyc or n1 + 1 < self.Nt) else G1
        _abi.check(self.lib.coflux_forcing_window_wait(self._w, G1, G2, _stream_handle(stream)), self.lib)
        self._busy = (G1, G2)
        # logical window [first, last] → descriptors into the ring
        lo, hi = self.first, self.last
        Gs = np.arange(lo, hi + 1)
        wt = self.times[Gs % self.Nt] + (Gs // self.Nt - cycle) * self.period
        nk, nj, ni = self.shape
        Hx, Hy, Hz = self.halo
        for name in self.names:
            data.atmos[name] = _RingSeries(_abi.Array(self.ptr[name], 1, ni, ni * nj, ni * nj * nk, Hx, Hy, Hz, 0))
        data.times = wt
        data.time_indexing = _abi.TIME_CLAMP if self.mode == _abi.TIME_CLAMP else _abi.TIME_LINEAR
        data.ring_start, data.ring_capacity = int(lo % self.capacity), self.capacity
        return t

    def release(self, stream=None):
        """Record on `stream` that the step enqueued last is the final reader of the levels it used."""
        from .engine import _stream_handle
        if self._busy is not None:
            _abi.check(self.lib.coflux_forcing_window_release(self._w, self._busy[0], self._busy[1], _stream_handle(stream)), self.lib)

    def stats(self):
        import ctypes as C
        b, n = C.c_int64(), C.c_int64()
        _abi.check(self.lib.coflux_forcing_window_stats(self._w, C.byref(b), C.byref(n)), self.lib)
        return b.value, n.value


class _RingSeries:
    """A series living in a DeviceForcingWindow ring: only its descriptor is known on the Python side."""

    def __init__(self, array):
        self._array = array

    def array(self):
        return self._array

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

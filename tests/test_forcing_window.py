"""`time_indices_in_memory` window management of the prescribed atmosphere (/root/reference/src/OMIPConfigurations/
atmosphere.jl:22-27): interpolating through an `InMemoryWindow` of 2–5 levels is bit-identical to
interpolating with the whole series in memory, for linear and for cyclical (repeat-year) time indexing, while the window
is re-based as the clock moves.  Arithmetic: the CPU oracle (the checker); under test: the host-side window logic."""
import numpy as np
import pytest

import climaocean.jl_b200 as cj
from climaocean.jl_b200 import _abi
from climaocean.jl_b200.forcing import InMemoryWindow
from oracle import pyoracle


def _interp(host, cfg, time):
    pyoracle.interpolate_atmosphere(cfg, host.atmos_series(), time, host.exchange_state())
    return {k: v.numpy().copy() for k, v in host.exchange.items()}


@pytest.mark.parametrize("mode", [_abi.TIME_LINEAR, _abi.TIME_CYCLICAL])
@pytest.mark.parametrize("length", [2, 3, 5])
def test_window_is_bit_identical_to_the_full_series(mode, length):
    grid = cj.LatitudeLongitudeGrid((24, 12, 1), latitude=(-60.0, 60.0), halo=(3, 3, 0))
    full = cj.SurfaceFluxData.synthetic(grid, Nt=8, atmos_size=(64, 32))
    full.time_indexing = mode
    cfg = cj.default_config(24, 12, 1, 64)
    win_data = cj.SurfaceFluxData.synthetic(grid, Nt=8, atmos_size=(64, 32))
    source = {n: f.data.copy() for n, f in full.atmos.items()}
    w = InMemoryWindow(source, full.times, length, full.atmos["u"].halo, None, mode)
    dt = float(full.times[1] - full.times[0])
    span = (full.times[-1] - full.times[0]) + (2.6 * dt if mode == _abi.TIME_CYCLICAL else 0.0)
    times = full.times[0] + np.concatenate([np.linspace(0.0, span, 23), [1.37 * dt, 6.999 * dt, 7.0 * dt, 0.0]])
    if mode == _abi.TIME_CYCLICAL:
        times = np.concatenate([times, [full.times[0] + 3 * w.period + 0.4 * dt, full.times[0] - 0.25 * dt]])
    for t in times:
        ref = _interp(full, cfg, float(t))
        t_eff = w.apply(win_data, float(t))
        assert len(w.window_times) <= max(length, 2) + (1 if mode == _abi.TIME_CYCLICAL else 0)
        got = _interp(win_data, cfg, t_eff)
        for k in ref:
            assert np.array_equal(got[k], ref[k]), (k, float(t), w.start)
    assert w.reloads >= (8 // max(length - 1, 1)) - 1          # the window really moved
    assert w.reloads < len(times)                                # … and was not reloaded on every call


def test_window_traffic_accounting_and_monotone_clock():
    grid = cj.LatitudeLongitudeGrid((16, 8, 1), latitude=(-60.0, 60.0), halo=(3, 3, 0))
    full = cj.SurfaceFluxData.synthetic(grid, Nt=8, atmos_size=(32, 16))
    source = {n: f.data.copy() for n, f in full.atmos.items()}
    w = InMemoryWindow(source, full.times, 4, full.atmos["u"].halo)
    level_bytes = sum(a[0].nbytes for a in source.values())
    dt = float(full.times[1] - full.times[0])
    for k in range(71):                                           # a clock advancing by a tenth of the forcing interval
        w.update(full.times[0] + 0.1 * k * dt)
    # levels 0..7 are visited once: windows start at levels 0, 3 and 6 — 4 + 4 + 2 levels (the series ends at level 7)
    assert w.reloads == 3 and w.bytes_loaded == 10 * level_bytes

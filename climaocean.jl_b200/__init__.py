"""climaocean.jl_b200 — Blackwell-native surface-flux engine behind ClimaOcean's coupling API.

Holds only what the hot path needs: `csrc/` (CUDA kernels + C ABI → lib/libcoflux.so), the ctypes
mirror of include/coflux.h, and the host-side mirror of the reference's interface for this path
(`OceanSeaIceModel`, `ComponentInterfaces`, `SimilarityTheoryFluxes`, `update_state`, `time_step`).
Importing the package requires the built CUDA library; there is no CPU fallback.
"""
from . import _abi
from ._abi import CofluxError, load_library, default_config
from .fields import Field, FieldTimeSeries, LatitudeLongitudeGrid, TripolarGrid, fractional_indices
from .state import SurfaceFluxData
from .engine import Engine
from .forcing import InMemoryWindow, DeviceForcingWindow
from .models import *  # noqa: F401,F403  (reference-facing names)

import os as _os

# fail loudly at import when lib/libcoflux.so has not been built.  COFLUX_NO_LIBRARY=1 is for bench.py's CPU arm
# (`--impl reference`), which only needs the grid / synthetic-input / struct-layout helpers and must not load the product.
if not _os.environ.get("COFLUX_NO_LIBRARY"):
    load_library()

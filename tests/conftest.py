import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """`gpu`-marked tests need a CUDA device: skip them (instead of failing with 'no NVIDIA driver') on a CPU-only box."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (B200); coflux has no CPU fallback")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the CUDA library and the oracle if they are missing (nvcc cross-compiles without a GPU)."""
    sys.path.insert(0, os.path.join(ROOT, "climaocean.jl_b200"))
    import importlib.util
    spec = importlib.util.spec_from_file_location("coflux_build", os.path.join(ROOT, "climaocean.jl_b200", "build.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    b.build_cuda()
    b.build_tools()
    b.build_oracle()
    sys.path.pop(0)
    yield

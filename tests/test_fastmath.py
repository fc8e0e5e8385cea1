"""The lean Float64 elementary functions of the CUDA hot loop (csrc/coflux_fastmath.cuh).

CPU: the same header compiled as plain C++ (SFU seeds emulated at 20-bit accuracy — what rcp.approx.ftz.f64 /
rsqrt.approx.ftz.f64 deliver on B200, tools/fm_check.cu) against 40-digit references: pins the polynomials, the
tables and the argument reductions.  GPU: lib/fm_check evaluates the real device code against the CUDA math library.
"""
import ctypes as C
import os
import re
import subprocess

import mpmath as mp
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
mp.mp.dps = 40


@pytest.fixture(scope="module")
def fm(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("fm") / "libfm.so")
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", out, os.path.join(ROOT, "tests", "fastmath_host.cpp")], check=True)
    return C.CDLL(out)


def _call(lib, name, *arrs):
    n = len(arrs[0])
    y = np.empty(n)
    getattr(lib, name)(*[a.ctypes.data_as(C.c_void_p) for a in arrs], y.ctypes.data_as(C.c_void_p), C.c_long(n))
    return y


def _ulps(y, ref):
    return float(max(abs((mp.mpf(float(a)) - b) / mp.mpf(2) ** (mp.floor(mp.log(abs(b), 2)) - 52)) for a, b in zip(y, ref)))


def test_rcp_div_sqrt_cbrt_within_two_ulp(fm):
    rng = np.random.default_rng(7)
    n = 4000
    x = np.exp(rng.uniform(-40, 40, n))
    a = np.exp(rng.uniform(-40, 40, n))
    assert _ulps(_call(fm, "fm_rcp", x), [1 / mp.mpf(float(v)) for v in x]) <= 1.0
    assert _ulps(_call(fm, "fm_div", a, x), [mp.mpf(float(u)) / mp.mpf(float(v)) for u, v in zip(a, x)]) <= 1.0
    assert _ulps(_call(fm, "fm_sqrt", x), [mp.sqrt(mp.mpf(float(v))) for v in x]) <= 1.0
    assert _ulps(_call(fm, "fm_cbrt", x), [mp.cbrt(mp.mpf(float(v))) for v in x]) <= 2.0


def test_log_absolute_error_at_rounding_level(fm):
    rng = np.random.default_rng(8)
    x = np.concatenate([np.exp(rng.uniform(-40, 40, 4000)), rng.uniform(0.5, 2.0, 2000), 1 + rng.uniform(-1e-3, 1e-3, 1000)])
    y = _call(fm, "fm_log", x)
    ref = [mp.log(mp.mpf(float(v))) for v in x]
    err = max(abs(mp.mpf(float(a)) - b) / max(1, abs(b)) for a, b in zip(y, ref))
    assert float(err) <= 2 * 2.0 ** -52       # ≤ 2 ulp of max(1, |ln x|): ln enters sums of O(10)


def test_exp_within_two_ulp(fm):
    rng = np.random.default_rng(9)
    e = rng.uniform(-80, 30, 6000)
    assert _ulps(_call(fm, "fm_exp", e), [mp.exp(mp.mpf(float(v))) for v in e]) <= 2.0


@pytest.mark.gpu
def test_device_functions_against_cuda_libm():
    exe = os.path.join(ROOT, "climaocean.jl_b200", "lib", "fm_check")
    assert os.path.exists(exe), "lib/fm_check missing: run __graft_entry__.build()"
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    vals = {k: float(v) for k, v in re.findall(r"(rcp|div|sqrt|cbrt|log|exp) ([0-9.]+)", out.split("\n")[0])}
    assert set(vals) == {"rcp", "div", "sqrt", "cbrt", "log", "exp"}, out
    assert max(vals.values()) <= 2.0, out          # ≤ 2 ulp from the CUDA math library on 4.2 M arguments each
    seeds = [float(v) for v in re.findall(r"rel err ([0-9.e+-]+)", out)]
    assert len(seeds) == 2 and max(seeds) < 2.0 ** -18, out   # the Newton step counts assume ≥ 18-bit seeds

"""Committed fixtures (tests/golden/*.npz — ORACLE-generated, see make_golden.py): SELF-CONSISTENCY only.  They guard the
oracle and the CUDA path against drift between commits; they say nothing about agreement with the Julia reference
(tests/test_julia_golden.py is the slot for reference-generated vectors).
CPU: the oracle still reproduces them (to 1e-13 in Float64 — glibc's libm dispatches FMA / non-FMA
variants per host CPU, so bitwise equality across machines is not guaranteed).  GPU: the CUDA path matches them within the
north_star tolerances without executing the oracle."""
import os

import numpy as np
import pytest

from tests.common import RTOL, gpu_update, make_case, rel_err
from tests.golden.make_golden import CASES, run_case

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", sorted(CASES))
def test_self_consistency_oracle_reproduces_its_own_fixture(name):
    gold = np.load(os.path.join(HERE, name + ".npz"))
    res = run_case(CASES[name])
    bits = CASES[name]["bits"]
    for k in gold.files:
        if k == "iterations":
            if bits == 64:
                assert np.array_equal(res[k], gold[k]), (name, k)
            continue
        assert rel_err(res[k], gold[k], bits) <= (1e-13 if bits == 64 else RTOL[32]), (name, k)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_self_consistency_cuda_matches_oracle_generated_fixture(name):
    spec = CASES[name]
    gold = np.load(os.path.join(HERE, name + ".npz"))
    grid, host, cfg = make_case(spec["Nx"], spec["Ny"], spec["Nz"], spec["bits"], flux_configuration=spec["flux_configuration"])
    gpu, _ = gpu_update(host, cfg)
    for k in gold.files:
        if k == "iterations":
            if spec["bits"] == 64:
                assert np.array_equal(gpu["_iterations"][0, 7:-7, 7:-7], gold[k]), (name, k)
            continue
        assert rel_err(gpu[k], gold[k], spec["bits"]) <= RTOL[spec["bits"]], (name, k)

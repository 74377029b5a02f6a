"""CPU-side checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads, exports every symbol
include/ibmfoam_b200.h declares, and refuses to run without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from ibmfoam_b200 import build, capi, cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return capi.load()


def test_header_symbols_all_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "ibmfoam_b200.h")).read()
    declared = set(re.findall(r"\b(ibm_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"ibm_ctx"}
    assert declared, "no declarations parsed"
    assert declared == set(capi.SYMBOLS), (declared ^ set(capi.SYMBOLS))
    for name in declared:
        assert hasattr(lib, name), f"{name} not exported"


def test_abi_version(lib):
    assert lib.ibm_abi_version() == 1


def test_substep_schedule_literal_loop(lib):
    # 100 or 101 sub-steps in fp64, integrating exactly dt
    import numpy as np
    from ibmfoam_b200.cloud import substep_schedule
    for dt, n in ((1e-3, 100), (1e-4, 101), (2.5e-5, 101)):
        dts = substep_schedule(dt)
        assert len(dts) == n


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from ibmfoam_b200.cloud import ParticleCloud
    with pytest.raises(capi.IbmError) as ei:
        ParticleCloud(cases.cylinder_re30())
    assert "IBM_E_CUDA" in str(ei.value)

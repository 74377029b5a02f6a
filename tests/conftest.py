import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import pyoracle
    pyoracle.build()
    return pyoracle


# TEST INFRASTRUCTURE ONLY: `IBM_TEST_EMU_LIB=tools/cuda_emu/_build/libibmfoam_b200_emu.so pytest -m gpu ...` points the
# ctypes loader at the CPU emulation build of the library (tools/cuda_emu/build_emu.py) so that the GPU parity tests can
# be exercised on a machine without a GPU.  The product package has no such switch: ibmfoam_b200.capi always loads
# ibmfoam_b200/lib/libibmfoam_b200.so, and that library refuses to run without a CUDA device.
_EMU = os.environ.get("IBM_TEST_EMU_LIB")
if _EMU:
    from ibmfoam_b200 import capi as _capi
    _capi.LIB_PATH = os.path.abspath(_EMU)


def pytest_collection_modifyitems(config, items):
    """On hardware no GPU test may hang the run: a kernel that never returns (a grid-wide barrier that cannot complete, say)
    is turned into a failure of that test after 10 minutes - `thread` method, because a Python-level signal cannot interrupt
    a blocked CUDA call.  Not under the emulation, where a test is legitimately slow (every CUDA thread is a fiber)."""
    if _EMU:
        return
    try:
        import pytest_timeout  # noqa: F401
    except Exception:
        return
    for item in items:
        if item.get_closest_marker("gpu") and not item.get_closest_marker("timeout"):
            item.add_marker(pytest.mark.timeout(600, method="thread"))

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _cuda_device_usable():
    """True when a CUDA device can actually be opened (driver present and new enough)"""
    try:
        import ctypes
        cu = ctypes.CDLL("libcuda.so.1")
        if cu.cuInit(0) != 0:
            return False
        n = ctypes.c_int(0)
        return cu.cuDeviceGetCount(ctypes.byref(n)) == 0 and n.value > 0
    except OSError:
        return False


def pytest_collection_modifyitems(config, items):
    # `pytest tests` on a box without a usable GPU skips the gpu-marked tests instead of failing them
    if _cuda_device_usable():
        return
    skip = pytest.mark.skip(reason="no usable CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    import helpers
    return helpers.load_oracle()


@pytest.fixture(scope="session")
def ref():
    import helpers
    lib = helpers.load_ref()
    if lib is None:
        pytest.skip("oracle/_ref/libref.so not built (reference sources absent)")
    return lib


@pytest.fixture(scope="session")
def scans():
    import helpers
    return helpers.fixture_scans()

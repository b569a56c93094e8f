import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


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

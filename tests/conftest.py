import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def c_oracle():
    from oracle import c_oracle as c
    c.build()
    return c


@pytest.fixture(scope="session")
def hx():
    """The built CUDA library; GPU tests fail loudly (not skip) when it is missing."""
    from gretel_b200 import _lib
    return _lib.load()

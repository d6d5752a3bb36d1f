import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def lib():
    """libpgstrom_cuda.so, built in-tree if it is not there yet."""
    import __graft_entry__ as ge
    ge.build()
    from pg_strom_b200 import _capi
    return _capi.load()


@pytest.fixture(scope="session")
def cuda(lib):
    """Initialises the device layer; GPU tests fail loudly without a device."""
    from pg_strom_b200 import gpupreagg as gp
    n = gp.cuda_init()
    assert n >= 1
    return n

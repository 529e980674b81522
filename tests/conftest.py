import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built():
    """Compile the CUDA library and the oracle once per session (nvcc cross-compiles on CPU)."""
    import __graft_entry__ as entry
    entry.build()
    return True


@pytest.fixture(scope="session")
def planner(built):
    from pdmpc_b200 import capi
    p = capi.Planner(0)   # no CPU fallback: raises without a device
    yield p
    p.close()

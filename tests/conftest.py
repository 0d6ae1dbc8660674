import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _build_native():
    """Make sure the oracle and the CUDA library exist (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as ge

    ge.build()
    yield

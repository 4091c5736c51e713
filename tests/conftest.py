import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run on the GPU box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the CUDA library and the oracle are built (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as g

    g.build(quiet=True)

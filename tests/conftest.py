import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "3d-particle-simulation-_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run on the GPU box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure libp3d.so and the oracle exist (nvcc cross-compiles without a GPU)."""
    import __graft_entry__ as g

    g.build()


@pytest.fixture(scope="session")
def default_params():
    import particle_3d as p3

    return p3.default_params_dict()

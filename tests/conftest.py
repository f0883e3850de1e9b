import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run on the GPU box with -m gpu)")


@pytest.fixture(scope="session")
def ctx():
    """One library context on cuda:0 for the whole GPU session (fails loudly without a GPU)."""
    from rrtqx_3d_b200.device import Context
    c = Context(0)
    yield c
    # not closed explicitly: trees / results created by the tests may outlive this fixture until
    # they are garbage collected, and each of them keeps the context alive through a reference


@pytest.fixture(scope="session")
def building2():
    from rrtqx_3d_b200 import workloads as W
    return W.building2_spheres()

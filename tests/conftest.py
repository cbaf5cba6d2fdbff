import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: minutes of CPU time (CUDA sources on CPU threads); runs only with CG3D_SLOW_TESTS=1")


def pytest_collection_modifyitems(config, items):
    if os.environ.get("CG3D_SLOW_TESTS", "0") == "1":
        return
    skip = pytest.mark.skip(reason="slow: set CG3D_SLOW_TESTS=1")
    for it in items:
        if "slow" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def lib():
    """Build (if needed) and load the C-ABI library; GPU tests call the product path through it."""
    from cagroup3d_b200 import build, _lib
    build.build()
    return _lib.load()

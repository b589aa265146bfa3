import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "ref: needs the unmodified reference built into oracle/_ref")


@pytest.fixture(scope="session")
def built_lib():
    """libsph_b200.so, built in-tree (nvcc cross-compiles without a GPU)."""
    from sph_b200.build import build
    return build()

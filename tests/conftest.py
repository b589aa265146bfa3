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


def pytest_collection_finish(session):
    """The emulator tests load ten builds of the kernel source (tests/emu/build_emu.py VARIANTS), ~10 s each when
    compiled one after the other: compile the stale ones side by side before the first test needs one."""
    if not any("test_emu_" in item.nodeid or "test_slab_gloo" in item.nodeid for item in session.items):
        return
    try:
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        from emu.build_emu import prebuild
        prebuild()
    except Exception as e:      # a test that needs a build will report the real error
        print(f"[conftest] emulator prebuild skipped: {e!r}", file=sys.stderr)

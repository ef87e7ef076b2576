import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    # The library must be built before anything imports gwen_b200 (in-tree .so, travels to the box).
    from gwen_b200 import _lib
    if _lib.needs_build() and os.path.exists("/usr/local/cuda/bin/nvcc"):
        _lib.build()


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # the C oracle is test infrastructure: build it once per session if the .so is absent
    from oracle import oracle as O
    O.build()


@pytest.fixture(scope="session")
def cuda_lib():
    """The product library.  Loading it must not need a GPU; computing with it does."""
    from forces_resilient_planner_b200 import _lib, build
    build.build()
    return _lib.load()

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure): compiled on demand; oracle/_ref only when /root/reference exists."""
    from oracle import pyoracle as po

    po.build(ref=os.path.exists("/root/reference/src/sim/kernels.cu"))
    return po


@pytest.fixture(scope="session")
def engine_lib():
    """The product library; built on demand (nvcc cross-compiles without a GPU)."""
    from spinwalk_b200 import _lib, build

    build.build()
    return _lib.load()

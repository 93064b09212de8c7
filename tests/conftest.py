import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_lib():
    import _oracle
    return _oracle.lib()


@pytest.fixture(scope="session")
def psim_lib():
    """The CUDA library. GPU tests must run the native code: fail (not skip) if it is missing."""
    import pedigreesim_b200
    return pedigreesim_b200.load()

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def built():
    """Make sure libapc.so and the oracle are compiled (no GPU needed to build)."""
    import __graft_entry__ as g
    g.build()
    return True


@pytest.fixture(scope="session")
def counter(built):
    from approx_counter_b200 import ApproxCounter
    c = ApproxCounter(0)
    yield c
    c.close()

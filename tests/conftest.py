import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from tests import oracle_util
    return oracle_util.load_oracle()


@pytest.fixture(scope="session")
def quda():
    """The CUDA library, initialised once per session.  GPU tests only."""
    import quda_b200 as q
    L = q.lib()
    L.initQuda(0)
    yield q
    L.endQuda()

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def built():
    """Make sure the CUDA module and the oracle libraries exist (no GPU needed to build)."""
    import __graft_entry__ as g
    from dtrenderer_b200 import api
    from oracle import dtro
    if not (os.path.exists(api.LIB_PATH) and dtro.available("port")):
        g.build()
    return True

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built():
    """Build the CUDA library + the C oracle once per session (no GPU needed)."""
    import __graft_entry__ as g

    g.build()
    return True


@pytest.fixture(scope="session")
def ctx(built):
    import datafusion_distributed_b200 as dfd

    c = dfd.WorkerContext(0)  # raises DfdError(DFD_ERR_CUDA) without a GPU: no CPU fallback
    yield c
    c.close()

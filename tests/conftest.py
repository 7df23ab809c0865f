import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def oracle_api():
    """The CPU oracle (test infrastructure); built on demand."""
    from oracle import oracle_cloud
    oracle_cloud.build()
    return oracle_cloud.api()


@pytest.fixture(scope="session")
def OracleCloud(oracle_api):
    from oracle.oracle_cloud import OracleCloud
    return OracleCloud


@pytest.fixture(scope="session")
def GpuCloud():
    """The product path.  No fallback: if libugf.so is missing or no GPU is visible this errors out."""
    import __graft_entry__ as g
    if not os.path.exists(g.LIB):
        g.build()
    from unigasfoam_b200.cloud import UniGasCloud
    return UniGasCloud

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the in-tree artefacts exist (no-op when they are current)."""
    import __graft_entry__ as g
    g.build_library()
    if not os.path.exists(os.path.join(ROOT, "oracle", "libtw_oracle.so")):
        g.build_oracle()


@pytest.fixture(scope="session")
def phold_tables_01():
    from scalesim_b200 import phold
    return phold.build_tables(0.1, 1.0)

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def ctx():
    """One libqsvc_b200 context on cuda:0 for the whole GPU session."""
    from qsvc_b200.mctf import Context
    c = Context(0)
    yield c
    c.close()

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (test infrastructure)."""
    import oracle

    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def ctx():
    """One libola_gpu context on cuda:0.  Fails loudly (no fallback) if the extension or the GPU is missing."""
    import olavm_b200

    c = olavm_b200.Context(0)
    yield c
    c.close()

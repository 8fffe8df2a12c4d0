import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.oracle_py import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def reference():
    """The reference's own code (oracle/_ref).  Only exists where it was built (this container)."""
    from oracle.oracle_py import Reference, ref_available
    if not ref_available():
        pytest.skip("oracle/_ref not built")
    return Reference()


@pytest.fixture(scope="session")
def ctx():
    from contextsv_b200.api import Context
    return Context(0)

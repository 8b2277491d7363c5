import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session")
def engine():
    """One engine on cuda:0 for the whole GPU test session. Fails loudly without a device."""
    from mamcts_b200.engine import Engine

    e = Engine(0)
    yield e
    e.close()

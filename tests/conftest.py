import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def native():
    import npp_b200
    return npp_b200._native


def pytest_collection_modifyitems(config, items):
    """Every GPU test gets a wall-clock limit (pytest-timeout): a dead-locked kernel must fail the run, not stall it.
    method="thread": the main thread would be blocked inside a CUDA call, where a signal handler never gets to run."""
    import pytest
    for item in items:
        if "gpu" in item.keywords and item.get_closest_marker("timeout") is None:
            item.add_marker(pytest.mark.timeout(600, method="thread"))

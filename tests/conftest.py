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
    """One lorb context on cuda:0 for the whole GPU session (fails loudly if the
    extension is missing or no device is usable — there is no fallback)."""
    from lorb_slam_b200 import capi
    c = capi.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: full BASELINE sizes against the oracle (tens of seconds of CPU)")


@pytest.fixture(scope="session", autouse=True)
def _library_present():
    """A fresh checkout has no built artefacts (they are git-ignored): compile the C-ABI library
    once (nvcc cross-compiles without a GPU) so that test order does not matter.  An existing
    library is left alone -- staleness is test_capi_cpu's business."""
    from lorb_slam_b200 import build
    if not os.path.exists(build.LIB):
        build.build_library()


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

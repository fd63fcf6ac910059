import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


# A kernel that does not end aborts itself after this many seconds (device-side watchdog, read at engine load) instead of
# the library's default of an hour: a hang in a test must cost a failed test, not the GPU box.
import os
os.environ.setdefault("PROCELL_WATCHDOG_S", "120")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib
    oracle_lib.lib()
    return oracle_lib


@pytest.fixture(scope="session")
def gpu_api():
    """The product API; GPU tests fail (not skip) when the CUDA extension is missing."""
    from cuda_pro_cell_b200 import _lib, api
    _lib.load()
    return api

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `pytest -m gpu`)")


def _gpu_count():
    try:
        from pvtrace_b200.engine import _cuda

        return _cuda.device_count()
    except Exception:
        return 0


@pytest.fixture(scope="session")
def gpu():
    """Skip-guard for gpu tests launched on a machine without a device (they are selected with -m gpu)."""
    n = _gpu_count()
    if n == 0:
        pytest.skip("no CUDA device")
    return n

import os, sys, subprocess
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the oracle (CPU) and the product library (nvcc cross-compiles without a GPU) once."""
    import __graft_entry__ as ge
    ge.build()
    yield


def has_gpu():
    try:
        from av1_base_b200 import abi
        return abi.lib().av1b_device_count() > 0
    except Exception:
        return False

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100) GPU; run with -m gpu on the GPU box")


def _gpu_available() -> bool:
    try:
        from cloud_merger_b200 import _lib
        return _lib.load().cm_device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def oracle():
    from oracle import cm_oracle_py
    cm_oracle_py.lib()
    return cm_oracle_py


@pytest.fixture(scope="session")
def gpu_ok():
    if not _gpu_available():
        pytest.fail("GPU test selected but libcloud_merger_gpu.so found no CUDA device (no CPU fallback exists)")
    return True

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_cuda():
    try:
        import rfs_slam_b200  # noqa: F401
        from rfs_slam_b200 import capi
        lib = capi.load_library()
        return lib.rfsb200_device_count() > 0
    except Exception:
        return False


@pytest.fixture(scope="session")
def cuda_required():
    """-m gpu tests must FAIL (not skip) when the extension is missing or no device is visible."""
    import rfs_slam_b200  # noqa: F401
    from rfs_slam_b200 import capi
    lib = capi.load_library()
    n = lib.rfsb200_device_count()
    assert n > 0, "no CUDA device visible: gpu tests cannot run (there is no CPU fallback)"
    return lib

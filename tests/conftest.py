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


def pytest_sessionfinish(session, exitstatus):
    """GPU sessions leave a parity report: per test the particles compared, those inside an epsilon band of a threshold
    (by rule) and those that differ from the reference (unexplained must be 0).  profiles/ holds the committed summary."""
    import json
    import os
    try:
        import helpers
    except Exception:
        return
    if not helpers.PARITY_RECORDS or not _has_cuda():   # (the CPU suite runs the same test bodies on the interpreter: no report)
        return
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = os.path.join(root, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    recs = helpers.PARITY_RECORDS
    tot = dict(tests=len(recs), particles=sum(r["particles"] for r in recs), in_epsilon_band=sum(r["in_epsilon_band"] for r in recs),
               differing=sum(r["differing"] for r in recs), differing_unexplained=sum(r["differing_unexplained"] for r in recs))
    with open(os.path.join(out, "parity_report.json"), "w") as f:
        json.dump(dict(total=tot, records=recs), f, indent=1)

"""TEST INFRASTRUCTURE ONLY — points the ctypes binding at the host-interpreter build of the ABI library
(tests/simt/_build/librfsb200_simt.so, see simt.h / simt_build.py) for the duration of a test, and back."""
from __future__ import annotations

import contextlib
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)


def load():
    """Build (if stale) and bind the interpreter library; the package's own loader state is untouched."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("simt_build", os.path.join(_HERE, "simt_build.py"))   # no sys.path games
    simt_build = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(simt_build)
    import rfs_slam_b200  # noqa: F401
    from rfs_slam_b200 import capi
    return capi.load_library(simt_build.build())


@contextlib.contextmanager
def interpreted(sm_count: int = 2, warps_per_cta: int | None = None):
    """Within the block every PHDUpdater talks to the interpreter build.  sm_count: how many "SMs" the fake device
    reports (few: several particles per warp and several CTAs per launch); warps_per_cta: RFSB200_WARPS_PER_CTA."""
    import rfs_slam_b200  # noqa: F401
    from rfs_slam_b200 import capi
    lib = load()
    saved_lib = capi._lib
    saved_env = {k: os.environ.get(k) for k in ("SIMT_SM_COUNT", "RFSB200_WARPS_PER_CTA")}
    os.environ["SIMT_SM_COUNT"] = str(sm_count)
    if warps_per_cta is None:
        os.environ.pop("RFSB200_WARPS_PER_CTA", None)
    else:
        os.environ["RFSB200_WARPS_PER_CTA"] = str(warps_per_cta)
    capi._lib = lib
    try:
        yield lib
    finally:
        capi._lib = saved_lib
        for k, v in saved_env.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v

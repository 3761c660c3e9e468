"""Per-phase times of one update (rfsb200_get_stage_times): python tools/stage_times.py [--config C3|C2|C3mf|N1k|N64k]"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rfs_slam_b200  # noqa
from rfs_slam_b200 import capi
from rfs_slam_b200.phd import PHDUpdater
import bench
ap = argparse.ArgumentParser()
ap.add_argument("--config", default="C3")
a = ap.parse_args()
wl, desc = bench.make_workload(a.config, 0)
up = PHDUpdater(wl.N, gm_capacity=256, z_capacity=32, lmk_dim=wl.dim)
up.load_workload(wl)
for k in range(3):
    so = up.update(wl.Z, flags=capi.UPDATE_NO_COMMIT)
print("product kernel: %.1f us" % so.elapsed_us)
for k in range(3):
    so = up.update(wl.Z, flags=capi.UPDATE_NO_COMMIT | capi.UPDATE_STAGE_TIMES)
    st = up.stage_times()
print("stage-timing kernel: %.1f us (events)" % so.elapsed_us)
for k, v in st.items():
    print("  %-22s %s" % (k, ("%.4f" % v) if isinstance(v, float) else v))

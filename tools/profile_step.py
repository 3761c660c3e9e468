"""Run a few PHD update steps of one configuration (for ncu / timing)."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rfs_slam_b200  # noqa
from rfs_slam_b200 import synth, capi
from rfs_slam_b200.phd import PHDUpdater
ap = argparse.ArgumentParser()
ap.add_argument("--N", type=int, default=8000)
ap.add_argument("--nM", type=int, default=200)
ap.add_argument("--nZ", type=int, default=30)
ap.add_argument("--sc", type=int, default=1)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--world", default="dense")
ap.add_argument("--cap", type=int, default=256, help="gm_capacity (Gaussians per particle)")
ap.add_argument("--vp", action="store_true", help="Victoria Park plugin set (C5 shape: --N 4000 --nM 150 --nZ 12 --sc 0)")
a = ap.parse_args()
if a.vp:
    wl = synth.make_vp_workload(N=a.N, nM=a.nM, nZ=a.nZ, use_cluster_process=a.sc, config_id=5)
else:
    wl = synth.make_workload(N=a.N, nM=a.nM, nZ=a.nZ, use_cluster_process=a.sc, world=a.world, config_id=3)
up = PHDUpdater(a.N, gm_capacity=a.cap, z_capacity=32, lmk_dim=wl.dim)
up.load_workload(wl)
for k in range(a.steps):
    so = up.update(wl.Z, flags=capi.UPDATE_NO_COMMIT)
    print("step %d: %.1f us  out %d" % (k, so.elapsed_us, so.gm_total_out))

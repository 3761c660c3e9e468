"""Print the merge diagnostics of one update (fallback reasons, pairs, clusters)."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rfs_slam_b200  # noqa
from rfs_slam_b200 import synth, capi
from rfs_slam_b200.phd import PHDUpdater
ap = argparse.ArgumentParser()
ap.add_argument("--N", type=int, default=8000); ap.add_argument("--nM", type=int, default=200)
ap.add_argument("--nZ", type=int, default=30); ap.add_argument("--sc", type=int, default=1)
ap.add_argument("--world", default="dense"); ap.add_argument("--id", type=int, default=3)
a = ap.parse_args()
wl = synth.make_workload(N=a.N, nM=a.nM, nZ=a.nZ, use_cluster_process=a.sc, world=a.world, config_id=a.id)
up = PHDUpdater(a.N, gm_capacity=256, z_capacity=32)
up.load_workload(wl)
for k in range(3):
    so = up.update(wl.Z, flags=capi.UPDATE_NO_COMMIT)
r = list(so.reserved)
print(f"N={a.N} nM={a.nM} nZ={a.nZ} sc={a.sc} {a.world}: {so.elapsed_us:.1f} us  out/particle {so.gm_total_out/a.N:.1f}  redo {so.n_merge_redo}  "
      f"fallback[nonPD {r[0]} pairs {r[1]} clusters {r[2]} members {r[3]}] conflicts {r[4]}  pairs/particle {r[5]/a.N:.1f}")

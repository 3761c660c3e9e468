"""Diagnostic: run the CUDA path and the oracle on a small synthetic workload and print the
differences (development aid; the assertions live in tests/)."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import rfs_slam_b200  # noqa
from rfs_slam_b200 import synth, capi
from rfs_slam_b200.phd import PHDUpdater
from oracle import binding as ob
import helpers

ap = argparse.ArgumentParser()
ap.add_argument("--N", type=int, default=256)
ap.add_argument("--nM", type=int, default=200)
ap.add_argument("--nZ", type=int, default=30)
ap.add_argument("--sc", type=int, default=1)
ap.add_argument("--prec", type=int, default=32)
ap.add_argument("--world", default="dense")
ap.add_argument("--extras", type=int, default=0)
ap.add_argument("--brute", type=int, default=0)
ap.add_argument("--sum_method", type=int, default=0)
a = ap.parse_args()

wl = synth.make_workload(N=a.N, nM=a.nM, nZ=a.nZ, use_cluster_process=a.sc, world=a.world,
                         parity_extras=bool(a.extras), config_id=7, cfg=dict(assignment_sum_method=a.sum_method))
t = time.time()
ref = ob.run(wl, sort_mode=ob.SORT_STABLE)
print(f"oracle: {time.time()-t:.3f}s  mean nM_out {ref.count.mean():.2f}")
up = PHDUpdater(a.N, gm_capacity=max(256, a.nM + 56), precision=a.prec, z_capacity=max(32, a.nZ))
up.set_model(wl.model)
up.set_filter_cfg(wl.cfg, brute_force_merge=bool(a.brute))
up.upload_maps(wl.count, wl.mean, wl.cov, wl.w)
up.set_poses(wl.pose, wl.pose_cov, wl.weight)
so = up.update(wl.Z, flags=capi.UPDATE_NO_NORMALIZE | capi.UPDATE_NO_COMMIT)
print("step_out: sum_w %.6e sum_w2 %.6e n_eff %.3f in %d out %d max %d overflow %d murty %d us %.1f" % (
    so.sum_w, so.sum_w2, so.n_eff, so.gm_total_in, so.gm_total_out, so.gm_max_out, so.n_overflow, so.n_murty, so.elapsed_us))
cnt, mean, cov, w = up.download_maps(which=1)
pw = up.get_weights(which=1)
tol = helpers.TOL32 if a.prec == 32 else helpers.TOL64
r = helpers.compare_maps(cnt, mean, cov, w, ref.count, ref.mean, ref.cov, ref.w, tol)
print("maps: structural/tolerance mismatches %d / %d ; max |dw| %.3e max |dmean| %.3e max cov rel %.3e" % (
    len(r["bad"]), a.N, r["max_w"], r["max_mean"], r["max_cov"]))
print(" bad particles:", r["bad"][:20], " counts dev/ref:", [(int(cnt[i]), int(ref.count[i])) for i in r["bad"][:10]])
rw = helpers.compare_weights(pw, ref.weight, tol)
print("weights: max |dlog w| %.3e n_bad %d" % (rw["max_dlog"], rw["n_bad"]), " sum_w ref %.6e" % ref.weight.sum())
mask, nfov = up.get_unused()
print("unused mask mismatches", int((mask != ref.unused_mask).sum()), " nfov mismatches", int((nfov != ref.n_in_fov).sum()))
fl = up.get_flags()
print("flags dev", np.bincount(fl, minlength=8)[:8], " ref murty", int((ref.flags & 2 > 0).sum()))
for k in range(3):
    so = up.update(wl.Z, flags=capi.UPDATE_NO_NORMALIZE | capi.UPDATE_NO_COMMIT)
print("timing (no-commit repeat): %.1f us -> %.3e updates/s" % (so.elapsed_us, a.N * a.nM * a.nZ / (so.elapsed_us * 1e-6)))

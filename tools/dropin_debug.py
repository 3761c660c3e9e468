import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import rfs_slam_b200  # noqa
from oracle import binding as ob
import helpers
import test_gpu_dropin as t
for sc in (1, 0):
    md, fc, poses, Z, nZ = t._scenario(sc=sc)
    kw = dict(pose_cov=[3e-5, 0, 0, 3e-5, 0, 3e-5], Q_lmk=[1e-5, 0, 0, 1e-5], neff_threshold=float(poses.shape[1]), seed48=7)
    for K in (5, 6, 8, 10):
        ref, nres_ref, tr_ref = ob.run_sequence("ref", poses[:K], Z[:K], nZ[:K], md, fc, **kw)
        got, nres, tr = ob.run_sequence("b200", poses[:K], Z[:K], nZ[:K], md, fc, precision=64, **kw)
        r = helpers.compare_maps(got.count, got.mean, got.cov, got.w, ref.count, ref.mean, ref.cov, ref.w, helpers.TOL64, ordered=False)
        print(f"sc={sc} K={K} nres {nres}/{nres_ref} trace {tr.tolist()} vs {tr_ref.tolist()} counts equal {np.array_equal(got.count, ref.count)} bad {r['bad'][:10]} "
              f"w equal {np.allclose(got.weight, ref.weight, rtol=1e-8)}")
        if np.array_equal(got.count, ref.count):
            print("   max |dmean| %.3e max |dcov| %.3e max |dw| %.3e" % (np.abs(np.sort(got.mean, 0) - np.sort(ref.mean, 0)).max(),
                  np.abs(np.sort(got.cov, 0) - np.sort(ref.cov, 0)).max(), np.abs(np.sort(got.w) - np.sort(ref.w)).max()))
        if r["bad"]:
            i = r["bad"][0]
            print("  particle", i, "counts", got.count[i], ref.count[i])
            o1, o2 = helpers.offsets(got.count), helpers.offsets(ref.count)
            print("  got w", np.sort(got.w[o1[i]:o1[i+1]])[::-1][:12])
            print("  ref w", np.sort(ref.w[o2[i]:o2[i+1]])[::-1][:12])

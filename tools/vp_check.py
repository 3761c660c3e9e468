"""GPU check of the Victoria Park path against the oracle (fp64 and fp32 device builds)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import rfs_slam_b200  # noqa
from rfs_slam_b200 import synth, capi
from oracle import binding as ob
import helpers
for sc in (1, 0):
    wl = synth.make_vp_workload(N=256, nM=120, nZ=12, use_cluster_process=sc, parity_extras=True, config_id=50 + sc)
    ref = ob.run(wl, sort_mode=ob.SORT_STABLE)
    for prec, tol in ((64, helpers.TOL64), (32, helpers.TOL32)):
        so, cnt, mean, cov, w, pw, up = helpers.run_device(wl, precision=prec, gm_capacity=192, work_capacity=256)
        r = helpers.compare_maps(cnt, mean, cov, w, ref.count, ref.mean, ref.cov, ref.w, tol)
        rw = helpers.compare_weights(pw, ref.weight, tol)
        mask, nfov = up.get_unused()
        print(f"sc={sc} prec={prec}: {so.elapsed_us:.0f} us, map mismatches {len(r['bad'])}/{wl.N} max|dw|={r['max_w']:.2e} "
              f"max|dmean|={r['max_mean']:.2e} maxcov={r['max_cov']:.2e}; weights bad {rw['n_bad']} max dlog {rw['max_dlog']:.2e}; "
              f"unused eq {np.array_equal(mask, ref.unused_mask)} nfov eq {np.array_equal(nfov, ref.n_in_fov)} "
              f"overflow {so.n_overflow} out {so.gm_total_out} ref {int(ref.count.sum())}")
        if r['bad'][:5]:
            for i in r['bad'][:3]:
                print("   particle", i, "dev", cnt[i], "ref", ref.count[i], "nfov", nfov[i], ref.n_in_fov[i])
        up.close()
# C5 shape, timed
import time
for sc in (0, 1):
    wl = synth.make_config("C5", use_cluster_process=sc)
    from rfs_slam_b200.phd import PHDUpdater
    up = PHDUpdater(wl.N, gm_capacity=192, work_capacity=256, z_capacity=16, lmk_dim=3)
    up.load_workload(wl)
    ts = []
    for k in range(6):
        so = up.update(wl.Z, flags=capi.UPDATE_NO_COMMIT)
        ts.append(so.elapsed_us)
    print(f"C5 sc={sc}: N={wl.N} nM=150 nZ={wl.nZ}: step {min(ts):.0f} us (runs {[int(t) for t in ts]}), out {so.gm_total_out}, overflow {so.n_overflow}, murty {so.n_murty}")
    t = time.time(); ref = ob.run(wl, which="ref", stage=5, sort_mode=ob.SORT_STD); print(f"   reference update() on host: {ref.elapsed_s*1e3:.0f} ms")
    up.close()

"""Replay an update() input dumped by the drop-in header (env RFSB200_DUMP_UPDATE=<k> while running an unchanged
driver) through the oracle / the compiled reference and, with --device, through the CUDA path.

File layout (little endian): int32 magic 'RFSB', N, LD, nZ, sizeof(model_desc), sizeof(filter_cfg); int32 count[N];
int64 total; f64 mean[total*LD], cov[total*NC], w[total], pose[N*3], pose_cov[N*6], weight[N], Z[nZ*LD];
rfsb200_model_desc; int32 scan_n; f64 scan[scan_n]; rfsb200_filter_cfg."""
import argparse, os, sys
import ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import rfs_slam_b200  # noqa
from rfs_slam_b200 import capi, synth


def load(path):
    b = open(path, "rb").read()
    hdr = np.frombuffer(b, np.int32, 6, 0)
    assert hdr[0] == 0x52465342
    N, LD, nZ, smd, sfc = (int(v) for v in hdr[1:6])
    NC = LD * (LD + 1) // 2
    o = 24
    def take(dt, n):
        nonlocal o
        a = np.frombuffer(b, dt, n, o).copy(); o += a.nbytes; return a
    cnt = take(np.int32, N); total = int(take(np.int64, 1)[0])
    mean = take(np.float64, total * LD).reshape(total, LD); cov = take(np.float64, total * NC).reshape(total, NC)
    w = take(np.float64, total); pose = take(np.float64, N * 3).reshape(N, 3); pcov = take(np.float64, N * 6).reshape(N, 6)
    weight = take(np.float64, N); Z = take(np.float64, nZ * LD).reshape(nZ, LD)
    assert smd == C.sizeof(capi.ModelDesc) and sfc == C.sizeof(capi.FilterCfg)
    md = capi.ModelDesc.from_buffer_copy(b[o:o + smd]); o += smd
    sn = int(take(np.int32, 1)[0]); scan = take(np.float64, sn)
    fc = capi.FilterCfg.from_buffer_copy(b[o:o + sfc])
    model = dict(model_id=md.model_id, R=list(md.R)[:LD * LD], Pd=md.Pd, clutter_intensity=md.clutter_intensity,
                 clutter_integral=md.clutter_integral, range_min=md.range_min, range_max=md.range_max,
                 range_buffer=md.range_buffer, innov_thr_range=md.innov_thr_range, innov_thr_bearing=md.innov_thr_bearing)
    if md.model_id == 2:
        model.update(bearing_min=md.bearing_min, bearing_max=md.bearing_max, Slb=md.Slb, buffer_zone_pd=md.buffer_zone_pd,
                     pd_table=list(md.pd_table)[:md.pd_table_n], scan=[float(v) for v in scan])
    cfg = {k: getattr(fc, k) for k, _ in capi.FilterCfg._fields_ if not k.startswith("reserved")}
    return synth.Workload(count=cnt, mean=mean, cov=cov, w=w, pose=pose, pose_cov=(pcov if pcov.any() else None),
                          weight=weight, Z=Z, model=model, cfg=cfg)


if __name__ == "__main__":
    ap = argparse.ArgumentParser(); ap.add_argument("dump"); ap.add_argument("--device", action="store_true")
    a = ap.parse_args()
    from oracle import binding as ob
    import helpers
    wl = load(a.dump)
    print("N", wl.N, "dim", wl.dim, "nZ", wl.nZ, "gaussians", int(wl.count.sum()))
    for st in (1, 2, 3, 4):
        o = ob.run(wl, stage=st, sort_mode=ob.SORT_STD)
        line = f"stage {st}: oracle out {int(o.count.sum())}"
        if ob.have_ref():
            r = ob.run(wl, which="ref", stage=st)
            cm = helpers.compare_maps(o.count, o.mean, o.cov, o.w, r.count, r.mean, r.cov, r.w, helpers.TOL64, ordered=False)
            cw = helpers.compare_weights(o.weight, r.weight, helpers.TOL64)
            line += f" | vs compiled reference: maps differ {cm['bad'][:8]} weights differ {list(cw['idx_bad'][:8])} max dlog {cw['max_dlog']:.2e}"
        print(line)
    if a.device:
        o = ob.run(wl, sort_mode=ob.SORT_STABLE)
        for prec, tol in ((64, helpers.TOL64), (32, helpers.TOL32)):
            so, cnt, mean, cov, w, pw, up = helpers.run_device(wl, precision=prec, gm_capacity=192, work_capacity=256)
            cm = helpers.compare_maps(cnt, mean, cov, w, o.count, o.mean, o.cov, o.w, tol, ordered=False)
            cw = helpers.compare_weights(pw, o.weight, tol)
            print(f"device fp{prec} vs oracle: maps differ {cm['bad'][:8]} weights differ {list(cw['idx_bad'][:8])} max dlog {cw['max_dlog']:.2e} murty {so.n_murty}")
            up.close()

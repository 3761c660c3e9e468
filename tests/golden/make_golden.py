"""Generates the golden fixtures under tests/golden/ by running the REFERENCE ITSELF
(oracle/_ref/libphd_ref.so = /root/reference compiled unmodified against oracle/compat).

Run in the build container only (needs /root/reference to have been compiled: `make -C oracle ref`):
    python tests/golden/make_golden.py
The .npz files are committed; tests read them on any box (no /root/reference needed).
"""
import ctypes as C
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
import rfs_slam_b200  # noqa: E402,F401
from rfs_slam_b200 import synth  # noqa: E402
from oracle import binding as ob  # noqa: E402

CASES = {
    # name: workload kwargs (small enough that the fixtures stay a few hundred KB)
    "sc_dense": dict(N=12, nM=60, nZ=12, use_cluster_process=1, config_id=101),
    "sc_extras": dict(N=16, nM=60, nZ=12, use_cluster_process=1, config_id=102, parity_extras=True),
    "mf_dense": dict(N=12, nM=60, nZ=12, use_cluster_process=0, config_id=103),
    "mf_extras": dict(N=16, nM=60, nZ=12, use_cluster_process=0, config_id=104, parity_extras=True),
    "mf_sparse_ragged": dict(N=12, nM=80, nZ=14, use_cluster_process=0, config_id=105, world="sparse", ragged=0.3),
    "mf_lowpd": dict(N=12, nM=60, nZ=12, use_cluster_process=0, config_id=106, model=dict(Pd=0.5, clutter_intensity=1e-2)),
    # partitions with nR + nC > 8: the reference's Murty-200 branch (include/RBPHDFilter.hpp:904-959, quirk Q7)
    "mf_murty": dict(N=16, nM=48, nZ=24, use_cluster_process=0, config_id=23, world="clumped",
                     model=dict(Pd=0.7, clutter_intensity=5e-3), cfg=dict(eval_point_gaussian_weight=0.1)),
}


# Victoria Park plugin set (3-D landmarks; rfs::MeasurementModel_VictoriaPark), synth.make_vp_workload
VP_CASES = {
    "vp_sc": dict(N=12, nM=70, nZ=10, use_cluster_process=1, config_id=151, parity_extras=True),
    "vp_mf": dict(N=12, nM=70, nZ=10, use_cluster_process=0, config_id=152, parity_extras=True),
    "vp_mf_ragged": dict(N=12, nM=90, nZ=14, use_cluster_process=0, config_id=153, ragged=0.3),
}


def main():
    assert ob.have_ref(), "oracle/_ref/libphd_ref.so missing: run `make -C oracle ref` in the build container"
    lib = C.CDLL(ob.REF_LIB)
    only = set(sys.argv[1:])   # optional: regenerate just these cases
    for name, kw in list(CASES.items()) + list(VP_CASES.items()):
        if only and name not in only:
            continue
        wl = synth.make_vp_workload(**kw) if name in VP_CASES else synth.make_workload(**kw)
        out = dict(kw_repr=repr(kw), count_in=wl.count, mean_in=wl.mean, cov_in=wl.cov, w_in=wl.w, pose=wl.pose,
                   pose_cov=(np.zeros(0) if wl.pose_cov is None else wl.pose_cov), weight_in=wl.weight, Z=wl.Z,
                   model_json=json.dumps(wl.model), cfg_json=json.dumps(wl.cfg))
        for st in (1, 2, 3, 4):
            r = ob.run(wl, which="ref", stage=st, n_threads=1)
            out[f"s{st}_count"] = r.count
            out[f"s{st}_mean"] = r.mean
            out[f"s{st}_cov"] = r.cov
            out[f"s{st}_w"] = r.w
            out[f"s{st}_wprev"] = r.wprev
            out[f"s{st}_weight"] = r.weight
            out[f"s{st}_unused"] = r.unused_mask
            out[f"s{st}_nfov"] = r.n_in_fov
        # the public entry point update() (stage 5): ends in normalizeWeights()
        r = ob.run(wl, which="ref", stage=5, n_threads=1)
        out["s5_weight"] = r.weight
        out["s5_count"] = r.count
        np.savez_compressed(os.path.join(HERE, f"phd_{name}.npz"), **out)
        print(name, "nM_out mean", r.count.mean())

    # ---- combinatorial KATs from the reference's own TUs -----------------------------------------
    rng = np.random.Generator(np.random.PCG64(777))
    lib.phd_ref_permanent.restype = C.c_double
    lib.phd_ref_permanent.argtypes = [C.c_void_p, C.c_int]
    lib.phd_ref_lexi_count.restype = C.c_int64
    lib.phd_ref_lexi_count.argtypes = [C.c_int, C.c_int]
    lib.phd_ref_partition.restype = C.c_int
    lib.phd_ref_partition.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.phd_ref_murty_sum.restype = C.c_double
    lib.phd_ref_murty_sum.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    if only:
        return
    kat = {}
    # MeasurementModel_VictoriaPark::probabilityOfDetection on individual landmarks (the scan geometry)
    wl = synth.make_vp_workload(N=4, nM=160, nZ=10, config_id=154, parity_extras=True)
    pdv, pdc = [], []
    for k in range(160):
        pd, close = ob.vp_pd(wl.model, wl.pose[0], wl.mean[k], wl.cov[k], which="ref")
        pdv.append(pd); pdc.append(int(close))
    kat["vp_pd_model_json"] = json.dumps(wl.model)
    kat["vp_pd_pose"] = wl.pose[0]
    kat["vp_pd_mean"] = wl.mean[:160]
    kat["vp_pd_cov"] = wl.cov[:160]
    kat["vp_pd_vals"] = np.array(pdv)
    kat["vp_pd_close"] = np.array(pdc)
    print("vp_pd values:", sorted(set(pdv)), "close:", sum(pdc))
    # permanents of random matrices
    perm_mats, perm_vals = [], []
    for n in range(1, 11):
        A = np.ascontiguousarray(rng.random((n, n)))
        perm_mats.append(A.ravel())
        perm_vals.append(lib.phd_ref_permanent(A.ctypes.data, n))
    kat["perm_n"] = np.arange(1, 11)
    kat["perm_mats"] = np.concatenate(perm_mats)
    kat["perm_vals"] = np.array(perm_vals)
    # enumeration counts
    pairs = [(3, 5), (2, 2), (3, 3), (4, 4), (1, 7), (1, 0), (0, 1), (0, 3), (2, 6), (4, 3)]
    kat["lexi_pairs"] = np.array(pairs)
    kat["lexi_counts"] = np.array([lib.phd_ref_lexi_count(a, b) for a, b in pairs])
    # partition labelling incl. Q6: random sparse tables + the shipped example + the Q6 trigger
    tabs = []
    ex = np.zeros((7, 7))
    for (i, j) in [(1, 4), (2, 1), (3, 3), (3, 5), (5, 2), (5, 3), (5, 5)]:
        ex[i, j] = 1.0
    tabs.append(ex)
    tabs.append(np.array([[0.0], [0.0], [0.7]]))
    for _ in range(30):
        nR, nC = int(rng.integers(0, 7)), int(rng.integers(1, 9))
        t = rng.random((nR, nC)) * (rng.random((nR, nC)) < 0.25)
        tabs.append(t)
    part_shapes, part_flat, part_out = [], [], []
    for t in tabs:
        t = np.ascontiguousarray(t, dtype=np.float64)
        nR, nC = t.shape
        a = np.zeros(64, np.int32); b = np.zeros(64, np.int32); z = np.zeros(64, np.int32)
        nP = lib.phd_ref_partition(t.ctypes.data, nR, nC, a.ctypes.data, b.ctypes.data, z.ctypes.data)
        part_shapes.append((nR, nC, nP))
        part_flat.append(t.ravel())
        part_out.append(np.concatenate([a[:nP], b[:nP], z[:nP]]))
    kat["part_shapes"] = np.array(part_shapes)
    kat["part_flat"] = np.concatenate(part_flat)
    kat["part_out"] = np.concatenate(part_out)
    # Murty-200 sums on partitions with nR + nC > 8
    ms_shapes, ms_flat, ms_pd, ms_cl, ms_val = [], [], [], [], []
    for (nR, nC) in [(4, 5), (5, 5), (3, 7), (6, 4), (5, 6), (2, 8), (7, 7)]:
        Lp = np.ascontiguousarray(rng.random((nR, nC)) * 20 * (rng.random((nR, nC)) < 0.7))
        pd = np.full(nR, 0.9) + 0.05 * rng.random(nR)
        cl = np.full(nC, 1e-2) * (0.5 + rng.random(nC))
        v = lib.phd_ref_murty_sum(Lp.ctypes.data, nR, nC, pd.ctypes.data, cl.ctypes.data)
        ms_shapes.append((nR, nC)); ms_flat.append(Lp.ravel()); ms_pd.append(pd); ms_cl.append(cl); ms_val.append(v)
    kat["murty_shapes"] = np.array(ms_shapes)
    kat["murty_flat"] = np.concatenate(ms_flat)
    kat["murty_pd"] = np.concatenate(ms_pd)
    kat["murty_cl"] = np.concatenate(ms_cl)
    kat["murty_vals"] = np.array(ms_val)
    np.savez_compressed(os.path.join(HERE, "kat_combinatorics.npz"), **kat)
    print("KATs:", {k: getattr(v, "shape", None) for k, v in kat.items()})


if __name__ == "__main__":
    main()

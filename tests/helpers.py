"""Comparison helpers shared by the parity tests (fp64 oracle vs device results)."""
from __future__ import annotations

import numpy as np

# SURVEY.md §8(d) tolerances, fp64 reference -> fp32 device
TOL32 = dict(w_abs=1e-5, w_rel=1e-4, mean_abs=1e-4, cov_rel=1e-3, logw_abs=2e-3, wnorm_rel=1e-3)
# fp64 device build vs fp64 oracle
TOL64 = dict(w_abs=1e-12, w_rel=1e-10, mean_abs=1e-10, cov_rel=1e-10, logw_abs=1e-9, wnorm_rel=1e-9)


def offsets(count):
    o = np.zeros(len(count) + 1, dtype=np.int64)
    np.cumsum(count, out=o[1:])
    return o


def compare_maps(count_a, mean_a, cov_a, w_a, count_b, mean_b, cov_b, w_b, tol, ordered=True):
    """Per-particle comparison. Returns dict with the particles whose STRUCTURE differs
    (different number of Gaussians or no one-to-one match) and the max errors over the rest."""
    N = len(count_a)
    oa, ob = offsets(count_a), offsets(count_b)
    bad = []
    max_w = max_mean = max_cov = 0.0
    for i in range(N):
        na, nb = int(count_a[i]), int(count_b[i])
        if na != nb:
            bad.append(i)
            continue
        if na == 0:
            continue
        ma, mb = mean_a[oa[i]:oa[i + 1]], mean_b[ob[i]:ob[i + 1]]
        ca, cb = cov_a[oa[i]:oa[i + 1]], cov_b[ob[i]:ob[i + 1]]
        wa, wb = w_a[oa[i]:oa[i + 1]], w_b[ob[i]:ob[i + 1]]
        perm = np.arange(na)
        if not ordered or np.abs(ma - mb).max() > tol["mean_abs"]:
            # match as sets: nearest (mean, weight) neighbour, must be one-to-one
            d = ((ma[:, None, :] - mb[None, :, :]) ** 2).sum(-1) + (wa[:, None] - wb[None, :]) ** 2
            perm = d.argmin(axis=1)
            if len(set(perm.tolist())) != na:
                bad.append(i)
                continue
        mb, cb, wb = mb[perm], cb[perm], wb[perm]
        ew = np.abs(wa - wb) - (tol["w_abs"] + tol["w_rel"] * np.abs(wb))
        em = np.abs(ma - mb).max()
        # Frobenius norm of the full symmetric matrix from its upper triangle (2-D: 3 entries, 3-D: 6)
        offd = np.array([0, 1, 0], bool) if ca.shape[1] == 3 else np.array([0, 1, 1, 0, 1, 0], bool)
        full = lambda c: np.concatenate([c, c[:, offd]], 1)
        ec = np.linalg.norm(full(ca) - full(cb), axis=1) / np.maximum(np.linalg.norm(full(cb), axis=1), 1e-300)
        if ew.max() > 0 or em > tol["mean_abs"] or ec.max() > tol["cov_rel"]:
            bad.append(i)
            continue
        max_w = max(max_w, float(np.abs(wa - wb).max()))
        max_mean = max(max_mean, float(em))
        max_cov = max(max_cov, float(ec.max()))
    return dict(bad=bad, max_w=max_w, max_mean=max_mean, max_cov=max_cov)


def compare_weights(w_dev, w_ref, tol, mask=None):
    """Unnormalised particle weights: compare log w (abs) and normalised weights (rel)."""
    w_dev = np.asarray(w_dev, dtype=np.float64)
    w_ref = np.asarray(w_ref, dtype=np.float64)
    if mask is None:
        mask = np.ones(len(w_ref), dtype=bool)
    ok = mask & (w_ref > 0) & (w_dev > 0)
    dl = np.abs(np.log(w_dev[ok]) - np.log(w_ref[ok]))
    return dict(max_dlog=float(dl.max()) if dl.size else 0.0,
                n_bad=int((dl > tol["logw_abs"]).sum()) + int((mask & ((w_ref > 0) != (w_dev > 0))).sum()),
                idx_bad=np.nonzero(ok)[0][dl > tol["logw_abs"]] if dl.size else np.array([], dtype=int))


# ---- golden fixtures (tests/golden/*.npz, generated from the compiled reference) -----------------
import json
import os

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_CASES = ["sc_dense", "sc_extras", "mf_dense", "mf_extras", "mf_sparse_ragged", "mf_lowpd", "mf_murty"]
GOLDEN_CASES_VP = ["vp_sc", "vp_mf", "vp_mf_ragged"]   # Victoria Park plugin set (3-D landmarks)


def load_golden(name):
    """-> (Workload rebuilt from the fixture's inputs, dict of the reference's outputs)."""
    from rfs_slam_b200 import synth
    g = np.load(os.path.join(GOLDEN_DIR, f"phd_{name}.npz"), allow_pickle=False)
    wl = synth.Workload(count=g["count_in"].astype(np.int32), mean=g["mean_in"], cov=g["cov_in"], w=g["w_in"],
                        pose=g["pose"], pose_cov=(g["pose_cov"] if g["pose_cov"].size else None),
                        weight=g["weight_in"], Z=g["Z"],
                        model=json.loads(str(g["model_json"])), cfg=json.loads(str(g["cfg_json"])))
    if name == "mf_murty":   # the reference's Murty-200 sums (quirk Q7) are reproduced behind this switch
        wl.cfg["murty_compat"] = 1
        wl.device_caps = dict(gm_capacity=128, work_capacity=256)   # the clumped world creates up to 168 Gaussians per particle
    return wl, g


def golden_stage(g, st):
    return dict(count=g[f"s{st}_count"], mean=g[f"s{st}_mean"], cov=g[f"s{st}_cov"], w=g[f"s{st}_w"],
                wprev=g[f"s{st}_wprev"], weight=g[f"s{st}_weight"], unused=g[f"s{st}_unused"], nfov=g[f"s{st}_nfov"])


# ---- epsilon-band exclusion (SURVEY.md §8d "structural decisions must agree except ...") ---------
def robust_mask(wl, eps=1e-3, rules=None):
    """Particles whose discrete decisions (gate, sensing limit, merge, prune, eval-point cut) do
    not change when every threshold is moved by +-eps (relative; absolute 1e-5 on range limits):
    for those an fp32 device result must agree STRUCTURALLY with the fp64 oracle.  The others sit
    inside the epsilon band of some threshold and are listed / excluded by the caller.
    rules (a dict, optional) receives how many particles each rule was the FIRST to exclude."""
    import copy
    from oracle import binding as ob
    rules = {} if rules is None else rules

    def note(name, before):
        rules[name] = rules.get(name, 0) + int(before.sum() - ok.sum())

    def run(scale, geometry=True, stages=(1, 3, 4)):
        w2 = copy.copy(wl)
        w2.cfg = dict(wl.cfg)
        w2.model = dict(wl.model)
        h = 0.5 * eps * scale  # thresholds are squared inside -> eps on the squared value
        for k in ("new_gaussian_create_innov_md_threshold", "meas_likelihood_md_threshold", "merging_threshold"):
            w2.cfg[k] = wl.cfg[k] * (1 + h)
        for k in ("pruning_threshold", "eval_point_gaussian_weight", "birth_gaussian_weight"):
            w2.cfg[k] = wl.cfg[k] * (1 + 1e-4 * scale)
        w2.model["range_max"] = wl.model["range_max"] + 1e-5 * scale
        w2.model["range_min"] = wl.model["range_min"] - 1e-5 * scale
        w2.model["range_buffer"] = wl.model["range_buffer"] + 2e-5 * scale
        for k in ("innov_thr_range", "innov_thr_bearing"):
            w2.model[k] = wl.model[k] * (1 + 1e-5 * scale)
        if wl.dim == 3 and geometry:
            # Victoria Park: the detection probability is a chain of floor / ceil / table look-ups on the
            # scan geometry; move its limits and the pose by far more than an fp32 rounding
            w2.model["bearing_max"] = wl.model["bearing_max"] + 1e-5 * scale
            w2.model["bearing_min"] = wl.model["bearing_min"] - 1e-5 * scale
            w2.model["scan"] = [v + 2e-5 * scale if v > 0 else v for v in wl.model["scan"]]
            w2.pose = wl.pose + np.array([3e-6, -3e-6, 2e-6]) * scale
        return [ob.run(w2, stage=s, sort_mode=ob.SORT_STABLE) for s in stages]

    base, up, dn = run(0), run(+1), run(-1)
    ok = np.ones(wl.N, dtype=bool)
    if not wl.cfg.get("use_cluster_process", 0):
        # multi-feature weighting sorts the mixture by weight (include/GaussianMixture.hpp:523-534) and both the
        # eval-point choice and the merge order follow that order: two weights closer than an fp32 rounding (but not
        # identical) are one more discrete decision an fp32 build may take differently
        s2 = ob.run(wl, stage=2, sort_mode=ob.SORT_STABLE)
        off = offsets(s2.count)
        for i in range(wl.N):
            w = s2.w[off[i]:off[i + 1]]
            if len(w) > 1:
                gap = w[:-1] - w[1:]
                # (2e-5 relative: the fp32 weights carry the rounding of exp(-md2 / 2); the interpreted sweep saw a
                #  swap at 5.9e-6)
                if np.any((gap > 0) & (gap < 2e-5 * np.maximum(w[:-1], 1e-30))):
                    ok[i] = False
    note("weight_sort_near_tie", np.ones(wl.N, dtype=bool))
    prev = ok.copy()
    for a, b, c in zip(base, up, dn):
        ok &= (a.count == b.count) & (a.count == c.count)
        ok &= (a.unused_mask == b.unused_mask) & (a.unused_mask == c.unused_mask)
        ok &= (a.n_in_fov == b.n_in_fov) & (a.n_in_fov == c.n_in_fov)
    note("count_or_mask_changes_under_threshold_move", prev)
    prev = ok.copy()
    # a merge decision can flip without changing the count (a small component joins another cluster): the merged
    # weights of the final mixture move then, while moving a threshold alone leaves them untouched.  (Thresholds only:
    # the Victoria Park runs above also move the geometry, which changes every weight a little.)
    if wl.dim == 3:
        final = [run(sc, geometry=False, stages=(4,))[0] for sc in (0, +1, -1)]
    else:
        final = [base[2], up[2], dn[2]]
        # Q3: the likelihood uses the UNWRAPPED bearing difference, so which side of +-pi a predicted bearing is wrapped to
        # decides whether a measurement near -+pi can pass the gate: a component whose predicted bearing lies within an
        # fp32 rounding of +-pi is one more discrete decision
        off_in = offsets(wl.count)
        for i in range(wl.N):
            m = wl.mean[off_in[i]:off_in[i + 1]]
            if len(m):
                b = np.arctan2(m[:, 1] - wl.pose[i, 1], m[:, 0] - wl.pose[i, 0]) - wl.pose[i, 2]
                b = np.abs((b + np.pi) % (2 * np.pi) - np.pi)
                if np.any(np.pi - b < 2e-6):
                    ok[i] = False
        note("predicted_bearing_at_pi", prev)
        prev = ok.copy()
    a, off = final[0], offsets(final[0].count)
    for o in final[1:]:
        off_o = offsets(o.count)
        for i in np.nonzero(ok)[0]:
            wa, wo = a.w[off[i]:off[i + 1]], o.w[off_o[i]:off_o[i + 1]]
            if len(wa) != len(wo) or not np.allclose(wa, wo, rtol=1e-9, atol=1e-12):
                ok[i] = False
    note("final_weights_move_under_threshold_move", prev)
    return ok


# ---- parity report (tests/conftest.py writes it to gpurun_out/parity_report.json at the end of a GPU session) ----------
PARITY_RECORDS = []


def parity_record(test, wl, robust, bad, rules=None, tol="fp32 vs fp64 reference (TOL32)"):
    """One line of the parity report: particles compared, particles inside an epsilon band (by the rule that put them
    there), particles that differ from the reference, and how many of those no rule explains (must be 0)."""
    bad = set(int(i) for i in bad)
    n = int(wl.N)
    if test is None:
        test = os.environ.get("PYTEST_CURRENT_TEST", "?").split(" ")[0]
    in_band = int(n - int(np.asarray(robust).sum())) if robust is not None else 0
    rec = dict(test=test, particles=n, in_epsilon_band=in_band, in_epsilon_band_by_rule=dict(rules or {}),
               differing=len(bad), differing_unexplained=len([i for i in bad if robust is None or robust[i]]),
               differing_frac=len(bad) / max(n, 1), tolerance=tol)
    PARITY_RECORDS.append(rec)
    return rec


def run_device(wl, precision=32, flags=None, gm_capacity=None, work_capacity=0, brute=False, z_capacity=None):
    """One update through the C ABI; returns (step_out, count, mean, cov, w, particle_weights, updater)."""
    from rfs_slam_b200 import capi
    from rfs_slam_b200.phd import PHDUpdater
    caps = getattr(wl, "device_caps", {})
    gm_capacity = gm_capacity or caps.get("gm_capacity")
    work_capacity = work_capacity or caps.get("work_capacity", 0)
    cap = gm_capacity or int(max(64, (int(wl.count.max()) + 63) // 8 * 8))
    up = PHDUpdater(wl.N, gm_capacity=cap, work_capacity=work_capacity, precision=precision,
                    z_capacity=z_capacity or max(8, wl.nZ), lmk_dim=wl.dim)
    up.set_model(wl.model)
    up.set_filter_cfg(wl.cfg, brute_force_merge=brute)
    up.upload_maps(wl.count, wl.mean, wl.cov, wl.w)
    up.set_poses(wl.pose, wl.pose_cov, wl.weight)
    f = capi.UPDATE_NO_NORMALIZE if flags is None else flags
    so = up.update(wl.Z, flags=f)
    which = 1 if (f & capi.UPDATE_NO_COMMIT) else 0
    cnt, mean, cov, w = up.download_maps(which)
    pw = up.get_weights(which)
    return so, cnt, mean, cov, w, pw, up


# ---- candidate-list births (include/RBPHDFilter.hpp:1000-1080): synthetic sequences without a device ---------------
BIRTH_CFG = dict(count_thr=3, check_thr=4, cur_count_thr=2, support_dist=2.0)


def birth_model(dim):
    """plugin descriptor of the range-bearing (dim 2) / Victoria Park (dim 3) set, as the update workloads use it"""
    from rfs_slam_b200 import synth
    return (synth.make_vp_workload(2, 8, 4, seed=1) if dim == 3 else synth.make_workload(2, 8, 4, seed=1)).model


def birth_scenario(dim, N, steps, seed, pose_cov=False):
    """Per step: poses, a measurement batch of noisy sightings of a few fixed objects plus clutter, random
    unused-measurement masks and nLandmarksInFOV_, and on two of the steps the parent slots of a resampling.
    Objects are re-observed over the steps, so candidates gather support, get promoted, age out or are erased."""
    rng = np.random.default_rng(seed)
    model = birth_model(dim)
    rlo, rhi = model["range_min"] + 1.0, min(model["range_max"] - 1.0, 30.0)
    off = np.pi / 2 if dim == 3 else 0.0                 # the Victoria Park sensor looks along theta - pi / 2
    n_obj = 8
    obj_r = rng.uniform(rlo, rhi, n_obj)
    obj_b = rng.uniform(0.4, 2.6, n_obj) if dim == 3 else rng.uniform(-2.5, 2.5, n_obj)
    th0 = 0.3 + off
    obj = np.stack([obj_r * np.cos(obj_b + th0 - off), obj_r * np.sin(obj_b + th0 - off), rng.uniform(0.3, 1.5, n_obj)], axis=1)
    pose = np.zeros((N, 3))
    pose[:, 2] = th0
    pose += rng.normal(0.0, [0.05, 0.05, 0.01], (N, 3))
    pcov = None
    if pose_cov and dim == 2:
        pcov = np.tile(np.array([0.02, 0.001, 0.0, 0.03, 0.0, 0.004]), (N, 1)) * rng.uniform(0.5, 1.5, (N, 1))
    out = []
    for t in range(steps):
        pose = pose + rng.normal(0.0, [0.02, 0.02, 0.002], (N, 3))
        seen = np.nonzero(rng.random(n_obj) < 0.75)[0]
        d = obj[seen, :2]
        zr = np.hypot(d[:, 0], d[:, 1]) + rng.normal(0.0, 0.05, len(seen))
        zb = np.arctan2(d[:, 1], d[:, 0]) - (th0 - off) + rng.normal(0.0, 0.01, len(seen))
        Z = np.stack([zr, zb, obj[seen, 2] + rng.normal(0.0, 0.05, len(seen))], axis=1)
        n_clutter = int(rng.integers(0, 4))
        clutter = np.stack([rng.uniform(rlo, rhi, n_clutter), rng.uniform(0.4, 2.6, n_clutter), rng.uniform(0.3, 1.5, n_clutter)], axis=1)
        Z = np.concatenate([Z, clutter])[:, :dim]
        Z = Z[rng.permutation(len(Z))]
        nZ = len(Z)
        mask = np.zeros(N, np.uint64)
        for z in range(nZ):
            mask |= (rng.random(N) < 0.6).astype(np.uint64) << np.uint64(z)
        nfov = rng.integers(0, 7, N).astype(np.int32)
        parent = None
        if t in (3, 5):
            # parent ids are not slot numbers after the first resampling (the reference indexes its per-slot lists with
            # getParentId() all the same): any slot can be named, including one that is itself a copy (chains)
            keep = rng.random(N) < 0.5
            keep[rng.integers(N)] = True
            parent = np.arange(N, dtype=np.int32)
            parent[~keep] = rng.integers(0, N, size=int((~keep).sum()))
            pose = pose[parent]
            if pcov is not None:
                pcov = pcov[parent]
        out.append(dict(pose=pose.copy(), Z=Z, mask=mask, nfov=nfov, parent=parent, pose_cov=None if pcov is None else pcov.copy()))
    return model, out

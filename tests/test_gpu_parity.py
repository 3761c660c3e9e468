"""Parity tests proper: the CUDA path (through the C ABI) against the golden vectors produced by
the reference itself and against the pinned fp64 oracle.  Tolerances are SURVEY.md §8(d):
fp32 device: |dw| <= 1e-5 + 1e-4 w, |dmean| <= 1e-4 m, cov rel Frobenius <= 1e-3, |dlog w_particle| <= 2e-3;
fp64 device build: 1e-10 relative with identical structure."""
import numpy as np
import pytest

import helpers
from helpers import TOL32, TOL64

pytestmark = pytest.mark.gpu


def _check(wl, prec, ref, so, cnt, mean, cov, w, pw, robust=None, max_excluded=0.02, rules=None):
    tol = TOL32 if prec == 32 else TOL64
    r = helpers.compare_maps(cnt, mean, cov, w, ref["count"], ref["mean"], ref["cov"], ref["w"], tol)
    bad = set(r["bad"])
    rw = helpers.compare_weights(pw, ref["weight"], tol)
    bad |= set(int(i) for i in rw["idx_bad"])
    if prec == 64 or robust is None:
        helpers.parity_record(None, wl, None, bad, tol="fp64 build vs fp64 reference (TOL64, identical structure)" if prec == 64 else "TOL32")
        assert not bad, f"particles differ from the reference: {sorted(bad)[:10]}"
    else:
        # fp32: a particle may only differ if it sits in the epsilon band of a threshold
        helpers.parity_record(None, wl, robust, bad, rules)
        unexplained = [i for i in bad if robust[i]]
        assert not unexplained, f"robust particles differ: {unexplained[:10]}"
        assert len(bad) <= max(1, int(max_excluded * wl.N)), f"{len(bad)} particles in an epsilon band"
    return r, rw


@pytest.mark.parametrize("prec", [32, 64])
@pytest.mark.parametrize("case", helpers.GOLDEN_CASES)
def test_against_reference_golden(cuda_required, case, prec):
    wl, g = helpers.load_golden(case)
    ref = helpers.golden_stage(g, 4)
    so, cnt, mean, cov, w, pw, up = helpers.run_device(wl, precision=prec)
    rules = {}
    robust = helpers.robust_mask(wl, rules=rules) if prec == 32 else None
    _check(wl, prec, ref, so, cnt, mean, cov, w, pw, robust, rules=rules)
    mask, nfov = up.get_unused()
    ok = robust if robust is not None else np.ones(wl.N, bool)
    assert np.array_equal(mask[ok], ref["unused"][ok])
    assert np.array_equal(nfov[ok], ref["nfov"][ok])
    assert so.n_launches >= 1 and so.n_overflow == 0
    # normalised weights of the public update() (ParticleFilter::normalizeWeights)
    up.normalize()
    wn = up.get_weights()
    assert wn.sum() == pytest.approx(1.0, abs=1e-12)
    assert np.allclose(wn, g["s5_weight"], rtol=1e-3 if prec == 32 else 1e-9, atol=1e-300)
    up.close()


@pytest.mark.parametrize("prec", [32, 64])
@pytest.mark.parametrize("kw", [
    dict(N=384, nM=200, nZ=30, use_cluster_process=1, config_id=11),
    dict(N=384, nM=100, nZ=20, use_cluster_process=0, config_id=12),
    dict(N=384, nM=200, nZ=30, use_cluster_process=1, config_id=13, parity_extras=True),
    dict(N=384, nM=100, nZ=20, use_cluster_process=0, config_id=14, parity_extras=True, ragged=0.2),
    dict(N=256, nM=180, nZ=30, use_cluster_process=1, config_id=15, world="sparse"),
    dict(N=256, nM=120, nZ=24, use_cluster_process=0, config_id=16, world="sparse", ragged=0.3),
    dict(N=256, nM=100, nZ=20, use_cluster_process=0, config_id=17, model=dict(Pd=0.6, clutter_intensity=5e-3)),
    dict(N=128, nM=60, nZ=64, use_cluster_process=1, config_id=18),
    dict(N=64, nM=1, nZ=1, use_cluster_process=0, config_id=19),
], ids=lambda k: "sc%d_nM%d_nZ%d_id%d" % (k["use_cluster_process"], k["nM"], k["nZ"], k["config_id"]))
def test_against_oracle(cuda_required, kw, prec):
    from oracle import binding as ob
    from rfs_slam_b200 import synth
    wl = synth.make_workload(**kw)
    o = ob.run(wl, sort_mode=ob.SORT_STABLE)
    ref = dict(count=o.count, mean=o.mean, cov=o.cov, w=o.w, weight=o.weight)
    so, cnt, mean, cov, w, pw, up = helpers.run_device(wl, precision=prec)
    rules = {}
    robust = helpers.robust_mask(wl, rules=rules) if prec == 32 else None
    _check(wl, prec, ref, so, cnt, mean, cov, w, pw, robust, rules=rules)
    if prec == 64:
        # identical structure AND identical order (weight desc, position asc) in the fp64 build
        assert np.array_equal(cnt, o.count)
        assert np.allclose(mean, o.mean, rtol=0, atol=1e-9)
        mask, nfov = up.get_unused()
        assert np.array_equal(mask, o.unused_mask) and np.array_equal(nfov, o.n_in_fov)
    assert so.gm_total_in == int(wl.count.sum())
    assert so.gm_total_out == int(cnt.sum())
    assert so.sum_w == pytest.approx(float(pw.sum()), rel=1e-12)
    assert so.sum_w2 == pytest.approx(float((pw * pw).sum()), rel=1e-12)
    up.close()


@pytest.mark.parametrize("sc", [0, 1])
def test_culled_merge_equals_exhaustive_merge(cuda_required, sc):
    from rfs_slam_b200 import synth
    wl = synth.make_workload(N=256, nM=150, nZ=30, use_cluster_process=sc, config_id=21 + sc, parity_extras=True)
    a = helpers.run_device(wl, precision=32, brute=False)
    b = helpers.run_device(wl, precision=32, brute=True)
    for k in range(1, 6):
        assert np.array_equal(a[k], b[k])  # bit-identical: same tests, same order
    a[6].close(); b[6].close()


def test_matrix_permanent_path_equals_enumeration(cuda_required):
    from rfs_slam_b200 import synth
    wl = synth.make_workload(N=256, nM=100, nZ=20, use_cluster_process=0, config_id=23,
                             model=dict(Pd=0.7, clutter_intensity=5e-3))
    a = helpers.run_device(wl, precision=64)
    wl.cfg["assignment_sum_method"] = 1
    b = helpers.run_device(wl, precision=64)
    assert np.allclose(a[5], b[5], rtol=1e-9)
    a[6].close(); b[6].close()


def test_multi_step_sequence_matches_chained_oracle(cuda_required):
    """three committed updates in a row; the oracle is chained on its own outputs"""
    import copy
    from oracle import binding as ob
    from rfs_slam_b200 import capi, synth
    from rfs_slam_b200.phd import PHDUpdater
    wl = synth.make_workload(N=128, nM=120, nZ=20, use_cluster_process=1, config_id=31, world="sparse")
    up = PHDUpdater(wl.N, gm_capacity=192, precision=64, z_capacity=32)
    up.load_workload(wl)
    cur = copy.copy(wl)
    rng = np.random.default_rng(5)
    for step in range(3):
        Z = wl.Z + rng.normal(0, [0.02, 0.005], wl.Z.shape)
        up.update(Z)  # default flags: commit + normalise
        cur.Z = Z
        o = ob.run(cur, sort_mode=ob.SORT_STABLE)
        wn = o.weight / o.weight.sum()
        cnt, mean, cov, w = up.download_maps()
        assert np.array_equal(cnt, o.count)
        assert np.allclose(mean, o.mean, atol=1e-9) and np.allclose(w, o.w, atol=1e-10)
        assert np.allclose(up.get_weights(), wn, rtol=1e-9)
        cur = synth.Workload(count=o.count, mean=o.mean, cov=o.cov, w=o.w, pose=wl.pose, pose_cov=wl.pose_cov,
                             weight=wn, Z=Z, model=wl.model, cfg=wl.cfg)
    up.close()


def test_no_commit_keeps_state_and_empty_Z_is_a_noop(cuda_required):
    from rfs_slam_b200 import capi, synth
    from rfs_slam_b200.phd import PHDUpdater
    wl = synth.make_workload(N=64, nM=40, nZ=8, use_cluster_process=1, config_id=41)
    up = PHDUpdater(wl.N, gm_capacity=64, precision=32, z_capacity=8)
    up.load_workload(wl)
    before = up.download_maps(0)
    so = up.update(wl.Z, flags=capi.UPDATE_NO_COMMIT)
    after0 = up.download_maps(0)
    after1 = up.download_maps(1)
    for a, b in zip(before, after0):
        assert np.array_equal(a, b)
    assert after1[0].sum() == so.gm_total_out and not np.array_equal(after1[0], before[0])
    # Q11: nZ == 0 -> nothing happens (include/RBPHDFilter.hpp:451-452)
    so0 = up.update(np.zeros((0, 2)))
    assert so0.n_launches == 0
    for a, b in zip(before, up.download_maps(0)):
        assert np.array_equal(a, b)
    assert np.array_equal(up.get_weights(0), wl.weight)
    up.close()


def test_pose_covariance_modes(cuda_required):
    from oracle import binding as ob
    from rfs_slam_b200 import synth
    wl = synth.make_workload(N=96, nM=80, nZ=16, use_cluster_process=0, config_id=51)
    rng = np.random.default_rng(1)
    for mode in (0, 1, 2):
        if mode == 0:
            wl.pose_cov = None
        elif mode == 1:
            wl.pose_cov = np.array([4e-5, 1e-5, -5e-6, 6e-5, 2e-6, 3e-5])
        else:
            s = rng.uniform(1e-5, 8e-5, (wl.N, 3))
            wl.pose_cov = np.stack([s[:, 0], 0.1 * s[:, 0], 0 * s[:, 0], s[:, 1], 0.05 * s[:, 1], s[:, 2]], 1)
        o = ob.run(wl, sort_mode=ob.SORT_STABLE)
        ref = dict(count=o.count, mean=o.mean, cov=o.cov, w=o.w, weight=o.weight)
        so, cnt, mean, cov, w, pw, up = helpers.run_device(wl, precision=64)
        _check(wl, 64, ref, so, cnt, mean, cov, w, pw)
        up.close()


def test_capacity_overflow_is_reported_not_fatal(cuda_required):
    from rfs_slam_b200 import synth
    wl = synth.make_workload(N=32, nM=60, nZ=30, use_cluster_process=1, config_id=61)
    so, cnt, *_rest, up = helpers.run_device(wl, precision=32, gm_capacity=64, work_capacity=64)
    assert so.n_overflow > 0            # 60 inputs + ~25 new Gaussians do not fit 64
    assert (cnt <= 64).all()
    assert (up.get_flags() & 1).sum() == so.n_overflow
    up.close()


def test_reference_api_surface(cuda_required):
    """getGMSize / getLandmark return conventions (include/RBPHDFilter.hpp:1153-1178) and error codes"""
    from rfs_slam_b200 import capi, synth
    from rfs_slam_b200.phd import PHDUpdater, RFSB200Error
    wl = synth.make_workload(N=16, nM=20, nZ=6, use_cluster_process=1, config_id=71)
    up = PHDUpdater(wl.N, gm_capacity=32, precision=32, z_capacity=8)
    with pytest.raises(RFSB200Error, match="ESTATE"):
        up.update(wl.Z)
    up.load_workload(wl)
    assert up.getGMSize(-1) == -1 and up.getGMSize(16) == -1 and up.getGMSize(3) == 20
    ok, mean, S, w = up.getLandmark(3, 5)
    k = 3 * 20 + 5
    assert ok and np.allclose(mean, wl.mean[k], atol=1e-6) and abs(w - wl.w[k]) < 1e-6
    assert np.allclose(S, [[wl.cov[k, 0], wl.cov[k, 1]], [wl.cov[k, 1], wl.cov[k, 2]]], atol=1e-8)
    assert up.getLandmark(3, 20)[0] is False and up.getLandmark(99, 0)[0] is False
    with pytest.raises(RFSB200Error, match="ECAPACITY"):
        up.update(np.zeros((9, 2)))
    with pytest.raises(RFSB200Error, match="EUNSUPPORTED"):
        up.set_model(dict(wl.model, model_id=3))
    with pytest.raises(RFSB200Error, match="EUNSUPPORTED"):
        up.set_filter_cfg(dict(wl.cfg, eval_point_count=33))
    bad = wl.count.copy(); bad[0] = 33
    with pytest.raises(RFSB200Error, match="ECAPACITY"):
        up.upload_maps(bad, np.zeros((bad.sum(), 2)), np.zeros((bad.sum(), 3)), np.zeros(bad.sum()))
    up.close()


def test_device_permanent_known_answers(cuda_required):
    from oracle import binding as ob
    from rfs_slam_b200.phd import PHDUpdater
    up = PHDUpdater(4, gm_capacity=32)
    der = {2: 1, 3: 2, 4: 9, 5: 44, 6: 265, 7: 1854, 8: 14833, 9: 133496, 10: 1334961, 11: 14684570, 12: 176214841}
    for n, v in der.items():   # test/MatrixPermanentTest.hpp:66-85
        got = up.permanent(np.ones((n, n)) - np.eye(n))[0]
        assert got == pytest.approx(v, rel=1e-12)
    rng = np.random.default_rng(3)
    A = rng.random((50, 7, 7))
    got = up.permanent(A)
    exp = np.array([ob.permanent(a) for a in A])
    assert np.allclose(got, exp, rtol=1e-11)
    up.close()


def test_large_partitions_are_flagged_and_summed_exactly(cuda_required):
    """Partitions with nR + nC > 8: the reference sums Murty's 200 best assignments (Q7, a truncated
    sum).  Default: the device computes the exact sum and flags the particle (bit 2) — the flagged set must be
    the oracle's Murty set, the exact weight can only be >= the truncated one (same maps), and
    unflagged particles agree to tolerance.  With rfsb200_filter_cfg::murty_compat the truncated sum is reproduced:
    every particle weight equals the reference's to 1e-9 (fp64 build)."""
    from oracle import binding as ob
    from rfs_slam_b200 import synth
    wl = synth.make_workload(N=64, nM=48, nZ=24, use_cluster_process=0, config_id=23, world="clumped",
                             model=dict(Pd=0.7, clutter_intensity=5e-3), cfg=dict(eval_point_gaussian_weight=0.1))
    o = ob.run(wl, sort_mode=ob.SORT_STABLE)
    murty = (o.flags & 2) > 0
    assert murty.sum() >= 4, "workload no longer exercises the Murty branch"
    so, cnt, mean, cov, w, pw, up = helpers.run_device(wl, precision=64, gm_capacity=128, work_capacity=256, z_capacity=24)
    f = up.get_flags()
    assert np.array_equal((f & 2) > 0, murty)
    assert so.n_murty == int(murty.sum()) and so.n_overflow == 0
    r = helpers.compare_maps(cnt, mean, cov, w, o.count, o.mean, o.cov, o.w, TOL64)
    assert not r["bad"]                                   # maps never depend on the weighting
    assert np.allclose(pw[~murty], o.weight[~murty], rtol=1e-9)
    assert (pw[murty] >= o.weight[murty] * (1 - 1e-9)).all()
    ratio = pw[murty] / o.weight[murty]                      # how much mass the 200 best miss
    assert np.isfinite(ratio).all() and ratio.max() < 1e3
    up.close()
    # murty_compat: the host replaces the exact sums of those partitions by the sums of Murty's 200 best assignments
    # (own k-best enumeration in the library, csrc/murty_compat.hpp): now EVERY particle weight is the reference's
    import copy
    wq = copy.copy(wl)
    wq.cfg = dict(wl.cfg, murty_compat=1)
    so, cnt, mean, cov, w, pw, up = helpers.run_device(wq, precision=64, gm_capacity=128, work_capacity=256, z_capacity=24)
    assert np.array_equal((up.get_flags() & 2) > 0, murty) and so.n_murty == int(murty.sum())
    assert np.allclose(pw, o.weight, rtol=1e-9, atol=0)
    assert so.sum_w == pytest.approx(float(pw.sum()), rel=1e-12)
    assert not helpers.compare_maps(cnt, mean, cov, w, o.count, o.mean, o.cov, o.w, TOL64)["bad"]
    up.normalize()
    assert up.get_weights().sum() == pytest.approx(1.0, abs=1e-12)
    up.close()
    so, cnt, mean, cov, w, pw32, up = helpers.run_device(wq, precision=32, gm_capacity=128, work_capacity=256, z_capacity=24)
    rw = helpers.compare_weights(pw32, o.weight, TOL32)
    assert rw["n_bad"] <= 1, rw
    up.close()


def test_partitions_beyond_the_on_chip_tables_are_summed_in_the_global_workspace(cuda_required):
    """A partition whose smaller side has 8..15 members does not fit the shared-memory tables of the assignment-sum
    DP: it is summed in the warp's block of the global workspace instead of being skipped (found by the randomised
    sweep: 24 eval points, P_D 0.5, clumped world).  No overflow is reported, the flagged set is the oracle's Murty set,
    the exact sum is >= the reference's truncated one — and forcing EVERY partition through the workspace
    (RFSB200_DP_ONCHIP_MAXB=0 / 2, a test aid) gives bit-identical particle weights, for both plugin sets."""
    import os
    from oracle import binding as ob
    from rfs_slam_b200 import synth
    cfg = dict(merging_threshold=0.5, merging_cov_inflation_factor=1.5, pruning_threshold=0.01, eval_point_count=24,
               eval_point_gaussian_weight=0.2, new_gaussian_create_innov_md_threshold=5.0, meas_likelihood_md_threshold=3.0)
    wl = synth.make_workload(N=11, nM=46, nZ=30, use_cluster_process=0, cfg=cfg, seed=522333278, ragged=0.3, parity_extras=True,
                             world="clumped", model=dict(Pd=0.5, clutter_intensity=1e-4, innov_thr_bearing=-1.0))
    wv = synth.make_vp_workload(N=24, nM=60, nZ=14, use_cluster_process=0, config_id=83, cfg=dict(eval_point_count=15))
    saved = os.environ.get("RFSB200_DP_ONCHIP_MAXB")
    try:
        for w_, caps in ((wl, dict(gm_capacity=512, work_capacity=1024)), (wv, dict(gm_capacity=128))):
            o = ob.run(w_, sort_mode=ob.SORT_STABLE)
            murty = (o.flags & 2) > 0
            got = {}
            for onchip in (None, 2, 0):
                if onchip is None:
                    os.environ.pop("RFSB200_DP_ONCHIP_MAXB", None)
                else:
                    os.environ["RFSB200_DP_ONCHIP_MAXB"] = str(onchip)   # read by rfsb200_create
                so, cnt, mean, cov, w, pw, up = helpers.run_device(w_, precision=64, **caps)
                f = up.get_flags()
                up.close()
                assert so.n_overflow == 0 and not (f & 4).any()
                assert np.array_equal((f & 2) > 0, murty)
                assert np.allclose(pw[~murty], o.weight[~murty], rtol=1e-9)
                assert np.isfinite(pw).all() and (pw[murty] >= o.weight[murty] * (1 - 1e-9)).all()
                got[onchip] = pw
            assert np.array_equal(got[None], got[2]) and np.array_equal(got[None], got[0])
    finally:
        if saved is None:
            os.environ.pop("RFSB200_DP_ONCHIP_MAXB", None)
        else:
            os.environ["RFSB200_DP_ONCHIP_MAXB"] = saved


def test_fused_normalisation_equals_two_launch_path(cuda_required):
    """RFSB200_UPDATE_FUSED_ALLREDUCE on one rank: sums + normalisation inside the update kernel."""
    from rfs_slam_b200 import capi, synth
    from rfs_slam_b200.phd import PHDUpdater
    wl = synth.make_workload(N=300, nM=80, nZ=16, use_cluster_process=1, config_id=81)
    res = []
    for flags in (capi.UPDATE_NO_COMMIT, capi.UPDATE_NO_COMMIT | capi.UPDATE_FUSED_ALLREDUCE):
        up = PHDUpdater(wl.N, gm_capacity=128, precision=32, z_capacity=16)
        up.load_workload(wl)
        so = up.update(wl.Z, flags=flags)
        res.append((so.n_launches, so.sum_w, up.get_weights(1)))
        assert not up.comm_error()
        up.close()
    assert res[0][0] == 2 and res[1][0] == 1
    assert res[0][1] == res[1][1]
    assert np.array_equal(res[0][2], res[1][2])
    assert res[1][2].sum() == pytest.approx(1.0, abs=1e-12)


def test_update_host_equals_the_separate_calls(cuda_required):
    """rfsb200_update_host = set_poses + update + get_weights + get_unused in one call with one synchronisation."""
    from rfs_slam_b200 import capi, synth
    from rfs_slam_b200.phd import PHDUpdater, pinned_array
    wl = synth.make_workload(N=512, nM=120, nZ=24, use_cluster_process=1, config_id=41, parity_extras=True)
    a = PHDUpdater(wl.N, gm_capacity=192, z_capacity=32)
    a.load_workload(wl)
    a.update(wl.Z)
    wa = a.get_weights()
    ma, fa = a.get_unused()
    b = PHDUpdater(wl.N, gm_capacity=192, z_capacity=32)
    b.set_model(wl.model); b.set_filter_cfg(wl.cfg); b.upload_maps(wl.count, wl.mean, wl.cov, wl.w)
    pose = pinned_array((wl.N, 3)); pose[:] = wl.pose
    w_in = pinned_array((wl.N,)); w_in[:] = wl.weight
    w_out = pinned_array((wl.N,)); mask = pinned_array((wl.N,), np.uint64); nfov = pinned_array((wl.N,), np.int32)
    so = b.update_host(pose, wl.pose_cov, w_in, np.ascontiguousarray(wl.Z), w_out=w_out, unused_out=mask, nfov_out=nfov,
                       want_stats=True)
    assert np.array_equal(w_out, wa) and np.array_equal(mask, ma) and np.array_equal(nfov, fa)
    assert so.n_launches == 2 and so.gm_total_in == int(wl.count.sum())   # update (reads the pinned inputs itself) + normalisation
    ca, cb = a.download_maps(), b.download_maps()
    for x, y in zip(ca, cb):
        assert np.array_equal(x, y)
    a.close(); b.close()


def test_update_host_zero_copy_equals_the_staged_path(cuda_required):
    """rfsb200_update_host with pinned caller buffers (inputs read and results written by the kernels directly, no
    copies) against the same call staged through copies (RFSB200_ZERO_COPY=0; also what pageable buffers get): same
    bits in every output, for every normalisation mode, pose-covariance mode, precision and both plugin sets."""
    from rfs_slam_b200 import capi, synth
    import os
    from rfs_slam_b200.phd import PHDUpdater, pinned_array
    saved = os.environ.get("RFSB200_ZERO_COPY")
    try:
        _zero_copy_cases(capi, synth, PHDUpdater, pinned_array)
    finally:
        if saved is None:
            os.environ.pop("RFSB200_ZERO_COPY", None)
        else:
            os.environ["RFSB200_ZERO_COPY"] = saved


def _zero_copy_cases(capi, synth, PHDUpdater, pinned_array):
    import os
    for dim, sc, prec, mode, flags in [(2, 1, 32, 1, capi.UPDATE_DEFAULT), (2, 0, 32, 2, capi.UPDATE_FUSED_ALLREDUCE),
                                       (2, 1, 64, 0, capi.UPDATE_NO_NORMALIZE), (3, 0, 32, 0, capi.UPDATE_FUSED_ALLREDUCE),
                                       (3, 1, 64, 0, capi.UPDATE_DEFAULT), (2, 1, 32, 1, capi.UPDATE_NO_COMMIT)]:
        mk = synth.make_vp_workload if dim == 3 else synth.make_workload
        wl = mk(N=96, nM=60, nZ=14, use_cluster_process=sc, config_id=70 + dim + 2 * sc)
        pcov = None
        if dim == 2:
            pcov = [None, wl.pose_cov, None][mode] if mode < 2 else None
        res = []
        for zero_copy in (1, 0):
            os.environ["RFSB200_ZERO_COPY"] = str(zero_copy)   # read by rfsb200_create
            up = PHDUpdater(wl.N, gm_capacity=128, z_capacity=16, precision=prec, lmk_dim=dim)
            up.set_model(wl.model); up.set_filter_cfg(wl.cfg); up.upload_maps(wl.count, wl.mean, wl.cov, wl.w)
            pose = pinned_array((wl.N, 3)); pose[:] = wl.pose
            w_in = pinned_array((wl.N,)); w_in[:] = wl.weight * np.linspace(0.5, 2.0, wl.N)
            cov = pcov
            if mode == 2:   # one covariance per particle
                cov = pinned_array((wl.N, 6)); cov[:] = np.asarray(wl.pose_cov).reshape(1, 6) * np.linspace(0.5, 1.5, wl.N)[:, None]
            w_out = pinned_array((wl.N,)); mask = pinned_array((wl.N,), np.uint64); nfov = pinned_array((wl.N,), np.int32)
            w_out[:] = -1; mask[:] = 12345; nfov[:] = -1
            so = up.update_host(pose, cov, w_in, np.ascontiguousarray(wl.Z), flags=flags, w_out=w_out, unused_out=mask,
                                nfov_out=nfov, want_stats=True)
            which = 1 if (flags & capi.UPDATE_NO_COMMIT) else 0
            res.append((w_out.copy(), mask.copy(), nfov.copy(), up.get_weights(which), up.download_maps(which), up.get_poses(),
                        (so.sum_w, so.sum_w2, so.gm_total_in, so.gm_total_out, so.gm_max_out, so.n_overflow, so.n_murty, so.n_launches)))
            # a second step on the same context, without statistics and without the optional outputs
            up.update_host(pose, cov, None, np.ascontiguousarray(wl.Z), flags=flags, w_out=w_out)
            res[-1] += (w_out.copy(), up.get_weights(which))
            up.close()
        a, b = res
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
        assert np.array_equal(a[0], a[3]) and np.array_equal(a[3], b[3])          # host copy == device state
        for x, y in zip(a[4], b[4]):
            assert np.array_equal(x, y)
        assert np.array_equal(a[5], b[5]) and np.array_equal(a[5], wl.pose)
        assert a[6][:7] == b[6][:7] and a[6][2] == int(wl.count.sum())
        assert a[6][7] == b[6][7] - (1 if dim == 2 else 0)   # launches: the 2-D kernels read the pinned inputs themselves
        assert np.array_equal(a[7], b[7]) and np.array_equal(a[8], b[8]) and np.array_equal(a[7], a[8])
        if flags & capi.UPDATE_NO_NORMALIZE:
            assert a[6][0] == pytest.approx(float(a[0].sum()), rel=1e-12)
        elif not (flags & capi.UPDATE_NO_COMMIT):
            assert a[0].sum() == pytest.approx(1.0, abs=1e-12)


def test_randomised_sweep_fp64(cuda_required):
    """tools/fuzz_parity.py: random sizes / thresholds / detection and clutter levels / world types for both plugin sets,
    fp64 device build against the oracle, exact structure (this sweep found the 1-warp-CTA table bug and the
    cancellation of the matrix-permanent path)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "fuzz_parity.py"), "60", "4242"], capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


@pytest.mark.parametrize("sc", [1, 0], ids=["sc", "mf"])
def test_stage_timing_build_equals_the_product_kernel(cuda_required, sc):
    """RFSB200_UPDATE_STAGE_TIMES runs the stage-timing instantiation of the update kernel (TimingInfo per phase,
    include/RBPHDFilter.hpp:152-167): same results (to the last bits of fp32), shares that add up, a wall-time split that adds up."""
    from rfs_slam_b200 import capi, synth
    from rfs_slam_b200.phd import PHDUpdater
    wl = synth.make_workload(N=600, nM=90, nZ=18, use_cluster_process=sc, config_id=33 + sc)
    res = []
    for flag in (0, capi.UPDATE_STAGE_TIMES):
        up = PHDUpdater(wl.N, gm_capacity=160, z_capacity=32)
        up.load_workload(wl)
        up.update(wl.Z, flags=capi.UPDATE_NO_NORMALIZE | flag)
        res.append((up.get_weights(), up.download_maps(), up.get_unused()))
        if flag:
            st = up.stage_times()
            shares = [st[k] for k in ("share_load", "share_map_update_kf", "share_weighting", "share_merge", "share_prune")]
            assert all(s >= 0 for s in shares) and sum(shares) == pytest.approx(1.0, abs=1e-9)
            assert st["share_map_update_kf"] > 0.02 and st["share_merge"] > 0.02
            assert st["kernel_us"] > 0 and st["kernel_us"] == pytest.approx(st["setup_us"] + st["particles_us"] + st["epilogue_us"], rel=1e-6)
            assert sum(st["merge_parts"].values()) <= st["share_merge"] + 1e-9
            if sc == 0:
                assert st["share_weighting"] > 0.1
        up.close()
    a, b = res
    # a different instantiation of the same source: the compiler may contract / schedule the fp32 arithmetic differently,
    # so the last bits can move (they did not in the single-cluster kernel, they do in the multi-feature weighting)
    assert np.allclose(a[0], b[0], rtol=1e-5, atol=0)
    assert np.array_equal(a[1][0], b[1][0])
    for x, y in zip(a[1][1:], b[1][1:]):
        assert np.allclose(x, y, rtol=1e-5, atol=1e-7)
    assert np.array_equal(a[2][0], b[2][0]) and np.array_equal(a[2][1], b[2][1])

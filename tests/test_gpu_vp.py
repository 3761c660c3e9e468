"""Parity of the Victoria Park CUDA path (phd_update_vp_kernel, through the C ABI) with the golden
vectors produced by the reference itself and with the pinned fp64 oracle.  Same tolerances as the 2-D
model (C5 allows 1e-3 m on the means, the tighter 1e-4 is used here)."""
import numpy as np
import pytest

import helpers
from helpers import TOL32, TOL64
from test_gpu_parity import _check

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("prec", [32, 64])
@pytest.mark.parametrize("case", helpers.GOLDEN_CASES_VP)
def test_vp_against_reference_golden(cuda_required, case, prec):
    wl, g = helpers.load_golden(case)
    ref = helpers.golden_stage(g, 4)
    so, cnt, mean, cov, w, pw, up = helpers.run_device(wl, precision=prec, gm_capacity=128)
    rules = {}
    robust = helpers.robust_mask(wl, rules=rules) if prec == 32 else None
    _check(wl, prec, ref, so, cnt, mean, cov, w, pw, robust, rules=rules)
    mask, nfov = up.get_unused()
    ok = robust if robust is not None else np.ones(wl.N, bool)
    assert np.array_equal(mask[ok], ref["unused"][ok])
    assert np.array_equal(nfov[ok], ref["nfov"][ok])
    assert so.n_launches >= 1 and so.n_overflow == 0
    up.normalize()
    wn = up.get_weights()
    assert wn.sum() == pytest.approx(1.0, abs=1e-12)
    assert np.allclose(wn, g["s5_weight"], rtol=1e-3 if prec == 32 else 1e-9, atol=1e-300)
    up.close()


@pytest.mark.parametrize("prec", [32, 64])
@pytest.mark.parametrize("kw", [
    dict(N=384, nM=150, nZ=12, use_cluster_process=0, config_id=61),
    dict(N=384, nM=150, nZ=12, use_cluster_process=1, config_id=62),
    dict(N=256, nM=120, nZ=16, use_cluster_process=0, config_id=63, parity_extras=True, ragged=0.2),
    dict(N=256, nM=200, nZ=8, use_cluster_process=1, config_id=64, parity_extras=True),
    dict(N=128, nM=60, nZ=40, use_cluster_process=0, config_id=65),
    dict(N=64, nM=1, nZ=1, use_cluster_process=0, config_id=66),
    dict(N=128, nM=100, nZ=12, use_cluster_process=0, config_id=67,
         model=dict(pd_table=(0.0, 0.3, 0.5, 0.6), buffer_zone_pd=0.55, innov_thr_range=-1.0, innov_thr_bearing=-1.0)),
], ids=lambda k: "sc%d_nM%d_nZ%d_id%d" % (k["use_cluster_process"], k["nM"], k["nZ"], k["config_id"]))
def test_vp_against_oracle(cuda_required, kw, prec):
    from oracle import binding as ob
    from rfs_slam_b200 import synth
    wl = synth.make_vp_workload(**kw)
    o = ob.run(wl, sort_mode=ob.SORT_STABLE)
    ref = dict(count=o.count, mean=o.mean, cov=o.cov, w=o.w, weight=o.weight)
    so, cnt, mean, cov, w, pw, up = helpers.run_device(wl, precision=prec, gm_capacity=256)
    rules = {}
    robust = helpers.robust_mask(wl, rules=rules) if prec == 32 else None
    _check(wl, prec, ref, so, cnt, mean, cov, w, pw, robust, rules=rules)
    if prec == 64:
        assert np.array_equal(cnt, o.count)
        assert np.allclose(mean, o.mean, rtol=0, atol=1e-9)
        mask, nfov = up.get_unused()
        assert np.array_equal(mask, o.unused_mask) and np.array_equal(nfov, o.n_in_fov)
    assert so.gm_total_in == int(wl.count.sum()) and so.gm_total_out == int(cnt.sum())
    assert so.sum_w == pytest.approx(float(pw.sum()), rel=1e-12)
    up.close()


def test_vp_full_size_c5_properties(cuda_required):
    """BASELINE config 5 shape (4 000 particles): run-to-run bit determinism, sharding invariance, and every particle
    against the oracle (at most 0.5 % may differ, each of them inside an epsilon band)."""
    from oracle import binding as ob
    from rfs_slam_b200 import synth
    wl = synth.make_config("C5")
    a = helpers.run_device(wl, precision=32, gm_capacity=192, work_capacity=256)
    b = helpers.run_device(wl, precision=32, gm_capacity=192, work_capacity=256)
    for k in range(1, 6):
        assert np.array_equal(a[k], b[k])
    a[6].close(); b[6].close()
    sh = wl.shard(1, 4)
    c = helpers.run_device(sh, precision=32, gm_capacity=192, work_capacity=256)
    lo, hi = wl.N // 4, wl.N // 2
    assert np.array_equal(c[1], a[1][lo:hi]) and np.array_equal(c[5], a[5][lo:hi])
    c[6].close()
    # ALL 4 000 particles against the oracle (0.4 s; the epsilon-band classification re-runs it ten times)
    o = ob.run(wl, sort_mode=ob.SORT_STABLE)
    rules = {}
    robust = helpers.robust_mask(wl, rules=rules)
    _check(wl, 32, dict(count=o.count, mean=o.mean, cov=o.cov, w=o.w, weight=o.weight), a[0], a[1], a[2], a[3], a[4], a[5],
           robust, max_excluded=0.005, rules=rules)


def test_vp_predict_maps_births_and_process_noise(cuda_required):
    """rfsb200_predict_maps for the 3-D model: births at MeasurementModel_VictoriaPark::inverseMeasure
    (src/MeasurementModel_VictoriaPark.cpp:75-102) from the unused measurements (descending index), P += Q."""
    from rfs_slam_b200 import synth
    wl = synth.make_vp_workload(N=64, nM=40, nZ=10, use_cluster_process=0, config_id=71)
    so, cnt0, mean0, cov0, w0, pw, up = helpers.run_device(wl, precision=64, gm_capacity=128, flags=0)
    mask, _ = up.get_unused()
    Q = np.array([1e-3, 1e-4, 0.0, 2e-3, 0.0, 5e-4])
    up.predict_maps(Q, add_births=True, birth_weight=0.02)
    cnt, mean, cov, w = up.download_maps()
    R = np.array(wl.model["R"]).reshape(3, 3)
    o0, o1 = helpers.offsets(cnt0), helpers.offsets(cnt)
    for i in range(wl.N):
        zs = [z for z in range(wl.nZ - 1, -1, -1) if (int(mask[i]) >> z) & 1]
        assert cnt[i] == cnt0[i] + len(zs)
        m_old, c_old = mean[o1[i]:o1[i] + cnt0[i]], cov[o1[i]:o1[i] + cnt0[i]]
        assert np.array_equal(m_old, mean0[o0[i]:o0[i + 1]])
        assert np.allclose(c_old, cov0[o0[i]:o0[i + 1]] + Q, rtol=1e-14, atol=0)
        px, py, th = wl.pose[i]
        th = th - np.pi / 2
        for k, z in enumerate(zs):
            r, b, d = wl.Z[z]
            j = o1[i] + cnt0[i] + k
            Hinv = np.array([[np.cos(th + b), -r * np.sin(th + b)], [np.sin(th + b), r * np.cos(th + b)]])
            P2 = Hinv @ R[:2, :2] @ Hinv.T
            assert np.allclose(mean[j], [px + r * np.cos(th + b), py + r * np.sin(th + b), d], rtol=0, atol=1e-9)
            want = np.array([P2[0, 0], P2[0, 1], 0.0, P2[1, 1], 0.0, R[2, 2]]) + Q
            assert np.allclose(cov[j], want, rtol=1e-10, atol=1e-15)
            assert w[j] == 0.02
    m2, _ = up.get_unused()
    assert not m2.any()
    # resampling moves all ten planes
    src = np.arange(wl.N, dtype=np.int32)[::-1].copy()
    up.resample(src)
    cnt2, mean2, cov2, w2 = up.download_maps()
    assert np.array_equal(cnt2, cnt[src])
    o2 = helpers.offsets(cnt2)
    for i in (0, 7, wl.N - 1):
        s = src[i]
        assert np.array_equal(mean2[o2[i]:o2[i + 1]], mean[o1[s]:o1[s + 1]])
        assert np.array_equal(cov2[o2[i]:o2[i + 1]], cov[o1[s]:o1[s + 1]])
    up.close()


def test_vp_model_dimension_mismatch_is_rejected(cuda_required):
    from rfs_slam_b200 import synth
    from rfs_slam_b200.phd import PHDUpdater, RFSB200Error
    up = PHDUpdater(8, gm_capacity=64, lmk_dim=2)
    with pytest.raises(RFSB200Error):
        up.set_model(synth.make_vp_workload(N=2, nM=20, nZ=4).model)
    up.close()
    up = PHDUpdater(8, gm_capacity=64, lmk_dim=3)
    with pytest.raises(RFSB200Error):
        up.set_model(dict(synth.DEFAULT_MODEL, clutter_integral=1.0))
    up.close()

"""Victoria Park plugin set (rfs::MeasurementModel_VictoriaPark / KalmanFilter_VictoriaPark, BASELINE
config 5): the 3-D oracle restatement (oracle/phd_oracle_vp.cpp) against golden vectors produced by the
reference itself (tests/golden/make_golden.py ran the reference's own MeasurementModel_VictoriaPark.cpp,
compiled unmodified, through RBPHDFilter<MotionModel_Ackerman2d, StaticProcessModel<Landmark3d>, ...>)."""
import json
import os

import numpy as np
import pytest

import helpers
from oracle import binding as ob

TIGHT = dict(w_abs=1e-13, w_rel=1e-11, mean_abs=1e-12, cov_rel=1e-10, logw_abs=1e-10, wnorm_rel=1e-10)


@pytest.mark.parametrize("case", helpers.GOLDEN_CASES_VP)
@pytest.mark.parametrize("stage", [1, 2, 3, 4])
def test_vp_oracle_matches_reference_golden(case, stage):
    wl, g = helpers.load_golden(case)
    assert wl.dim == 3 and wl.model["model_id"] == 2
    ref = helpers.golden_stage(g, stage)
    got = ob.run(wl, stage=stage, sort_mode=ob.SORT_STD, n_threads=1)
    assert np.array_equal(got.count, ref["count"])
    r = helpers.compare_maps(got.count, got.mean, got.cov, got.w, ref["count"], ref["mean"], ref["cov"], ref["w"],
                             TIGHT, ordered=True)
    assert r["bad"] == []
    assert np.allclose(got.wprev, ref["wprev"], rtol=1e-12, atol=0)
    assert helpers.compare_weights(got.weight, ref["weight"], TIGHT)["n_bad"] == 0
    assert np.array_equal(got.unused_mask, ref["unused"])
    assert np.array_equal(got.n_in_fov, ref["nfov"])


@pytest.mark.parametrize("case", helpers.GOLDEN_CASES_VP)
def test_vp_public_update_normalises(case):
    wl, g = helpers.load_golden(case)
    got = ob.run(wl, stage=4, sort_mode=ob.SORT_STD, n_threads=1)
    assert np.allclose(got.weight / got.weight.sum(), g["s5_weight"], rtol=1e-10, atol=0)


def test_vp_detection_probability_kat():
    """MeasurementModel_VictoriaPark::probabilityOfDetection (src/MeasurementModel_VictoriaPark.cpp:153-266)
    on 160 landmarks: every entry of the P_D table, in / out of the range and bearing limits, buffer zone."""
    k = np.load(os.path.join(helpers.GOLDEN_DIR, "kat_combinatorics.npz"), allow_pickle=False)
    model = json.loads(str(k["vp_pd_model_json"]))
    vals, close = k["vp_pd_vals"], k["vp_pd_close"]
    assert set(np.unique(vals)) == set(model["pd_table"])    # the fixture reaches every table entry
    assert 0 < close.sum() < len(close)
    for i in range(len(vals)):
        pd, cl = ob.vp_pd(model, k["vp_pd_pose"], k["vp_pd_mean"][i], k["vp_pd_cov"][i])
        assert pd == vals[i] and int(cl) == close[i], i


def test_vp_empty_measurement_set_and_empty_maps():
    wl, _ = helpers.load_golden("vp_sc")
    wl.Z = np.zeros((0, 3))
    got = ob.run(wl, n_threads=1)
    assert np.array_equal(got.count, wl.count) and np.array_equal(got.mean, wl.mean)   # Q11
    wl, _ = helpers.load_golden("vp_sc")
    empty = np.nonzero(wl.count == 0)[0]
    assert len(empty) >= 1                                                                 # Q10
    got = ob.run(wl, n_threads=1)
    assert np.all(got.count[empty] == 0) and np.all(got.weight[empty] == wl.weight[empty])
    assert np.all(got.unused_mask[empty] == (1 << wl.nZ) - 1)


@pytest.mark.skipif(not ob.have_ref(), reason="oracle/_ref/libphd_ref.so not built (needs /root/reference)")
@pytest.mark.parametrize("sc", [0, 1])
def test_vp_oracle_vs_compiled_reference_live(sc):
    from rfs_slam_b200 import synth
    wl = synth.make_vp_workload(N=24, nM=110, nZ=14, use_cluster_process=sc, config_id=350 + sc, parity_extras=True,
                                ragged=0.1)
    for st in (1, 2, 3, 4):
        a = ob.run(wl, which="oracle", stage=st, sort_mode=ob.SORT_STD, n_threads=1)
        b = ob.run(wl, which="ref", stage=st, n_threads=1)
        r = helpers.compare_maps(a.count, a.mean, a.cov, a.w, b.count, b.mean, b.cov, b.w, TIGHT, ordered=True)
        assert r["bad"] == []
        assert helpers.compare_weights(a.weight, b.weight, TIGHT)["n_bad"] == 0
        assert np.array_equal(a.unused_mask, b.unused_mask) and np.array_equal(a.n_in_fov, b.n_in_fov)

"""Candidate-list births on the device (rfsb200_birth_candidates = RBPHDFilter::addBirthGaussians() with
birthGaussianMeasurementCountThreshold_ != 1, reference include/RBPHDFilter.hpp:1000-1080) against the oracle
restatement (oracle/phd_oracle_births.cpp, itself pinned to the reference's own addBirthGaussians() by
tests/test_oracle_births.py): multi-step sequences update -> [resample] -> births -> P += Q for both plugin sets and
both precisions.  The unused-measurement masks and nLandmarksInFOV_ the births consume are the device update's own; the
candidate arithmetic is fp64 on both sides, so lists and counters must agree to rounding, and the Gaussians that become
real must appear behind the existing ones of the map, in the reference's order."""
import numpy as np
import pytest

import helpers
from oracle import binding

pytestmark = pytest.mark.gpu

BCFG = dict(count_thr=3, check_thr=4, cur_count_thr=2, support_dist=2.0)
BIRTH_W = 0.01


def _workload(dim, N, seed):
    from rfs_slam_b200 import synth
    if dim == 3:
        wl = synth.make_vp_workload(N, 24, 10, use_cluster_process=1, ragged=0.7, seed=seed)
    else:
        wl = synth.make_workload(N, 24, 10, use_cluster_process=1, world="sparse", ragged=0.7, seed=seed)
    return wl


def _resample_plan(N, rng):
    """survivors keep their slot, the slots of the others take copies (ParticleFilter::resample placement)"""
    keep = rng.random(N) < 0.6
    keep[rng.integers(N)] = True
    survivors = np.nonzero(keep)[0]
    src = np.arange(N, dtype=np.int32)
    src[~keep] = rng.choice(survivors, size=int((~keep).sum()))
    parent = src.copy()
    # the lists follow getParentId(), which after the first resampling need not be the slot the map came from and may
    # name a slot that is itself a copy: chains towards lower slots have to be resolved in the reference's order
    chain = (~keep) & (rng.random(N) < 0.5)
    parent[chain] = rng.integers(0, N, size=int(chain.sum()))
    aux = np.where(parent == np.arange(N), np.arange(N), np.where(parent > np.arange(N), parent, -1)).astype(np.int32)
    return src, parent, aux


def _run_sequence(dim, prec, N=48, steps=7, seed=11, pose_cov=False, bcfg=None):
    from rfs_slam_b200 import capi
    from rfs_slam_b200.phd import PHDUpdater
    bcfg = bcfg or BCFG
    wl = _workload(dim, N, seed)
    rng = np.random.default_rng(seed)
    cap = 96
    up = PHDUpdater(N, gm_capacity=cap, precision=prec, z_capacity=16, lmk_dim=dim)
    up.set_model(wl.model)
    up.set_filter_cfg(wl.cfg)
    up.upload_maps(wl.count, wl.mean, wl.cov, wl.w)
    state = binding.BirthState(N, dim, cap=capi.BIRTH_CAND_CAP)
    pose = wl.pose.copy()
    pcov = None
    if pose_cov and dim == 2:
        pcov = np.tile(np.array([0.02, 0.001, 0.0, 0.03, 0.0, 0.004]), (N, 1)) * rng.uniform(0.5, 1.5, (N, 1))
    tol = 2e-6 if prec == 32 else 1e-11
    n_real = n_support = n_dropped = n_copies = n_chain = 0
    for t in range(steps):
        pose = pose + rng.normal(0.0, [0.02, 0.02, 0.002], pose.shape)
        Z = wl.Z.reshape(-1, dim) + rng.normal(0.0, 0.02 if t % 3 else 0.3, (wl.nZ, dim))
        if dim == 3:
            Z[:, 2] = np.abs(Z[:, 2]) + 0.05
        Z[:, 0] = np.abs(Z[:, 0]) + 0.5
        up.set_poses(pose, pcov, wl.weight)
        up.update(Z)
        mask, nfov = up.get_unused()
        parent = None
        if t in (2, 4):
            src, parent, aux = _resample_plan(N, rng)
            up.resample(src, aux)
            pose = pose[src]
            if pcov is not None:
                pcov = pcov[src]
            n_copies += int((parent != np.arange(N)).sum())
            lower = parent < np.arange(N)
            n_chain += int((lower & (parent[parent] < parent)).sum())   # the lower parent is itself a copy from below
        cnt0, mean0, cov0, w0 = up.download_maps(0)
        before = state.copy()
        add_n, add_mean, add_cov = binding.birth_candidates(wl.model, bcfg, state, pose, Z, mask, nfov, parent=parent,
                                                            pose_cov=pcov, add_cap=cap)
        up.birth_candidates(BIRTH_W, bcfg["support_dist"], bcfg["count_thr"], bcfg["check_thr"], bcfg["cur_count_thr"],
                            parent=parent)
        n, mean, cov, sup, chk = up.get_birth_candidates()
        assert np.array_equal(n, state.n), (t, np.nonzero(n != state.n)[0][:8])
        for i in range(N):
            k = int(n[i])
            assert np.array_equal(sup[i, :k], state.support[i, :k]) and np.array_equal(chk[i, :k], state.checks[i, :k]), (t, i)
            assert np.allclose(mean[i, :k], state.mean[i, :k], rtol=1e-11, atol=1e-11), (t, i)
            assert np.allclose(cov[i, :k], state.cov[i, :k], rtol=1e-10, atol=1e-13), (t, i)
        cnt1, mean1, cov1, w1 = up.download_maps(0)
        assert np.array_equal(cnt1, cnt0 + add_n), (t, np.nonzero(cnt1 != cnt0 + add_n)[0][:8])
        o0, o1 = helpers.offsets(cnt0), helpers.offsets(cnt1)
        for i in range(N):
            a, b = int(o1[i]), int(o1[i]) + int(cnt0[i])
            assert np.array_equal(mean1[a:b], mean0[o0[i]:o0[i] + cnt0[i]]) and np.array_equal(w1[a:b], w0[o0[i]:o0[i] + cnt0[i]])
            k = int(add_n[i])
            assert np.allclose(mean1[b:b + k], add_mean[i, :k], rtol=tol, atol=tol * 10), (t, i)
            assert np.allclose(cov1[b:b + k], add_cov[i, :k], rtol=tol * 10, atol=tol), (t, i)
            assert np.allclose(w1[b:b + k], BIRTH_W, rtol=1e-7)
        assert not up.get_unused()[0].any()      # the masks are consumed
        assert not (up.get_flags() & 32).any()
        n_real += int(add_n.sum())
        n_support += int((state.support[:, :].sum() - before.support.sum()) > 0)
        n_dropped += int(np.maximum(before.n - state.n, 0).sum())
        q = np.array([1e-4, 0.0, 1e-4]) if dim == 2 else np.array([1e-4, 0, 0, 1e-4, 0, 1e-5])
        up.predict_maps(Q_lmk=q, add_births=False)
    up.close()
    return dict(real=n_real, support=n_support, dropped=n_dropped, copies=n_copies, lists=int(state.n.sum()), chains=n_chain)


@pytest.mark.parametrize("prec", [32, 64])
@pytest.mark.parametrize("dim", [2, 3], ids=["rngbrg", "victoriapark"])
def test_candidate_list_births_follow_the_reference_through_a_sequence(cuda_required, dim, prec):
    r = _run_sequence(dim, prec)
    # the sequence exercises every branch: candidates opened, supported, promoted, aged out, copied after a resampling
    assert r["real"] > 0 and r["support"] > 0 and r["dropped"] > 0 and r["copies"] > 0 and r["chains"] > 0, r


def test_candidate_list_births_with_the_victoria_park_thresholds(cuda_required):
    """cfg/rbphdslam_VictoriaPark_artificialClutter.xml:71-77: support 5, checks 10, current-count 2, distance 2"""
    r = _run_sequence(3, 32, N=32, steps=14, seed=23, bcfg=dict(count_thr=5, check_thr=10, cur_count_thr=2, support_dist=2.0))
    assert r["real"] > 0 and r["lists"] > 0


def test_candidate_list_births_with_a_pose_covariance(cuda_required):
    """MeasurementModel_RngBrg::measure adds Hr Sx Hr^T of the particle's pose covariance to the expected measurement"""
    r = _run_sequence(2, 64, N=24, steps=5, seed=5, pose_cov=True)
    assert r["real"] > 0


def test_candidate_list_overflow_and_capacity_flags(cuda_required):
    """a candidate beyond RFSB200_BIRTH_CAND_CAP is dropped with flag bit 32; a real Gaussian beyond gm_capacity with bits 1 | 8"""
    from rfs_slam_b200 import capi
    from rfs_slam_b200.phd import PHDUpdater
    wl = _workload(3, 8, 3)
    cap = int(wl.count.max()) + 2
    up = PHDUpdater(wl.N, gm_capacity=cap, work_capacity=192, precision=64, z_capacity=16, lmk_dim=3)
    up.set_model(wl.model)
    up.set_filter_cfg(wl.cfg)
    up.upload_maps(wl.count, wl.mean, wl.cov, wl.w)
    up.set_poses(wl.pose, None, wl.weight)
    C = capi.BIRTH_CAND_CAP
    n = np.full(wl.N, C, np.int32)
    mean = np.zeros((wl.N, C, 3)); mean[..., 0] = 1e4 + 50.0 * np.arange(C)[None, :]; mean[..., 2] = 1.0   # far from everything
    cov = np.zeros((wl.N, C, 6)); cov[..., 0] = cov[..., 3] = cov[..., 5] = 0.01
    up.set_birth_candidates(n, mean, cov, np.ones((wl.N, C), np.int32), np.zeros((wl.N, C), np.int32))
    up.update(wl.Z)
    mask, nfov = up.get_unused()
    up.birth_candidates(BIRTH_W, 2.0, 50, 100, 0)           # nothing is promoted or aged out: new candidates do not fit
    fl = up.get_flags()
    full = (mask != 0) & (nfov > 0)
    assert full.any() and ((fl[full] & 32) != 0).all()
    assert (up.get_birth_candidates()[0] == C).all()
    so = up.update(wl.Z, want_stats=True)      # the drop shows in the next update's result
    assert so.n_overflow >= int(full.sum())
    # now every candidate has enough support: 64 real Gaussians per particle do not fit the maps
    up.update(wl.Z)
    up.set_birth_candidates(n, mean, cov, np.full((wl.N, C), 5, np.int32), np.zeros((wl.N, C), np.int32))
    up.birth_candidates(BIRTH_W, 2.0, 2, 100, 0)
    fl = up.get_flags()
    sizes = up.get_gm_sizes()
    assert ((fl & 9) == 9).all() and (sizes >= cap).all() and (sizes == sizes[0]).all()   # filled to the (rounded-up) capacity
    assert (up.get_birth_candidates()[0] == 0).all()
    so = up.update(wl.Z, want_stats=True)      # the next update reports the drop
    assert so.n_overflow == wl.N
    up.close()

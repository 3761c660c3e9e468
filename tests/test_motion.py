"""Particle propagation: the motion-model restatement against golden vectors from the reference's
own step() functions (CPU), and rfsb200_propagate against the restatement (GPU): exact for step(), statistical
(moments) for the noise — the reference draws from a host mt19937 seeded by an unseeded rand()."""
import os

import numpy as np
import pytest

import helpers
from oracle import motion


def _golden():
    return np.load(os.path.join(helpers.GOLDEN_DIR, "motion_steps.npz"))


def test_motion_restatement_matches_reference_steps():
    g = _golden()
    for k in range(len(g["poses"])):
        o = motion.odometry2d_step(g["poses"][k], g["u_odometry"][k])
        a = motion.ackerman2d_step(g["poses"][k], g["u_ackerman"][k], g["dt"][k], *g["ackerman_params"])
        assert np.allclose(o, g["out_odometry"][k], rtol=0, atol=1e-12)
        assert np.allclose(a, g["out_ackerman"][k], rtol=0, atol=1e-12)
    assert (np.abs(g["out_ackerman"][:, 2] - g["poses"][:, 2]) > 3).any()   # the heading wrap is exercised


def _updater(poses, precision=64):
    from rfs_slam_b200 import synth
    from rfs_slam_b200.phd import PHDUpdater
    N = len(poses)
    up = PHDUpdater(N, gm_capacity=64, precision=precision)
    up.upload_maps(np.zeros(N, np.int32), np.zeros((0, 2)), np.zeros((0, 3)), np.zeros(0))
    up.set_poses(poses, None, np.ones(N))
    return up


@pytest.mark.gpu
def test_device_step_is_exact(cuda_required):
    g = _golden()
    poses = np.repeat(g["poses"], 3, axis=0)
    for k in (0, 1, 7, 20):
        up = _updater(poses)
        up.propagate("odometry2d", g["u_odometry"][k], use_model_noise=False)
        assert np.allclose(up.get_poses(), motion.odometry2d_step(poses, g["u_odometry"][k]), rtol=0, atol=1e-12)
        up.close()
        up = _updater(poses)
        up.propagate("ackerman2d", g["u_ackerman"][k], dt=float(g["dt"][k]), use_model_noise=False,
                     ackerman=tuple(g["ackerman_params"]))
        want = motion.ackerman2d_step(poses, g["u_ackerman"][k], g["dt"][k], *g["ackerman_params"])
        assert np.allclose(up.get_poses(), want, rtol=0, atol=1e-12)
        up.close()


@pytest.mark.gpu
def test_device_noise_moments_and_streams(cuda_required):
    N = 200000
    pose0 = np.array([3.0, -2.0, 0.7])
    poses = np.tile(pose0, (N, 1))
    Q = np.array([[4e-4, 1e-4, 0.0], [1e-4, 9e-4, -5e-5], [0.0, -5e-5, 1e-4]])
    u = np.array([0.4, 0.05, 0.1])
    Su = np.diag([1e-3, 4e-4, 2.5e-4]); Su[0, 1] = Su[1, 0] = 2e-4
    up = _updater(poses)
    up.propagate("odometry2d", u, Q=Q, input_cov=Su, use_model_noise=True, use_input_noise=True, seed=11, step=3)
    x = up.get_poses()
    f0, J = motion.numeric_jacobian(lambda v: motion.odometry2d_step(pose0, v), u)
    want_cov = J @ Su @ J.T + Q        # first order in the input noise + additive model noise
    d = x - f0
    d[:, 2] = (d[:, 2] + np.pi) % (2 * np.pi) - np.pi
    assert np.abs(d.mean(0)).max() < 5 * np.sqrt(np.diag(want_cov).max() / N) + 2e-4   # second-order bias of the rotation
    got_cov = np.cov(d.T)
    assert np.allclose(got_cov, want_cov, rtol=0.03, atol=3e-6)
    # the stream is a function of (seed, step, particle): same call -> same poses, other step / seed -> other poses
    up2 = _updater(poses)
    up2.propagate("odometry2d", u, Q=Q, input_cov=Su, use_model_noise=True, use_input_noise=True, seed=11, step=3)
    assert np.array_equal(up2.get_poses(), x)
    up2.propagate("odometry2d", u, Q=Q, input_cov=Su, use_model_noise=True, use_input_noise=True, seed=11, step=4)
    y = up2.get_poses()
    assert np.abs(np.corrcoef((y - motion.odometry2d_step(x, u))[:, 0], d[:, 0])[0, 1]) < 0.02
    up.close(); up2.close()
    # Ackerman: input noise on (velocity, steering) only, no model noise -> poses carry no covariance afterwards
    up = _updater(poses, precision=32)
    prm = (0.76, 2.83, 3.78, 0.5)
    ua, Sa = np.array([6.0, 0.15]), np.diag([0.2, 0.025]) * 0.05
    up.propagate("ackerman2d", ua, input_cov=Sa, dt=0.025, use_model_noise=False, use_input_noise=True, ackerman=prm, seed=5)
    x = up.get_poses()
    f0, J = motion.numeric_jacobian(lambda v: motion.ackerman2d_step(pose0, v, 0.025, *prm), ua)
    assert np.allclose(np.cov((x - f0).T), J @ Sa @ J.T, rtol=0.05, atol=1e-7)
    up.close()


@pytest.mark.gpu
def test_propagated_poses_feed_the_update(cuda_required):
    """poses propagated on the device are the poses the next update uses (incl. pose covariance = Q, Q1)"""
    from oracle import binding as ob
    from rfs_slam_b200 import capi, synth
    wl = synth.make_workload(N=128, nM=60, nZ=12, use_cluster_process=1, config_id=77)
    so, cnt, mean, cov, w, pw, up = helpers.run_device(wl, precision=64, flags=capi.UPDATE_NO_COMMIT | capi.UPDATE_NO_NORMALIZE)
    Q = np.diag([2e-5, 3e-5, 1e-5])
    u = np.array([0.02, -0.01, 0.004])
    up.propagate("odometry2d", u, Q=Q, use_model_noise=True, seed=1, step=0)
    poses = up.get_poses()
    up.update(wl.Z, flags=capi.UPDATE_NO_NORMALIZE)
    got = up.download_maps()
    import copy
    w2 = copy.copy(wl)
    w2.pose = poses
    w2.pose_cov = np.array([Q[0, 0], Q[0, 1], Q[0, 2], Q[1, 1], Q[1, 2], Q[2, 2]])
    o = ob.run(w2, sort_mode=ob.SORT_STABLE)
    r = helpers.compare_maps(got[0], got[1], got[2], got[3], o.count, o.mean, o.cov, o.w, helpers.TOL64)
    assert r["bad"] == []
    assert helpers.compare_weights(up.get_weights(), o.weight, helpers.TOL64)["n_bad"] == 0
    up.close()

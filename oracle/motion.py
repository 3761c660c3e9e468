"""numpy restatement of the reference's motion models — TEST INFRASTRUCTURE, NOT PRODUCT.

  odometry2d_step : MotionModel_Odometry2d::step  (src/ProcessModel_Odometry2D.cpp:41-89)
  ackerman2d_step : MotionModel_Ackerman2d::step  (src/ProcessModel_Ackerman2D.cpp:49-77)
  sample_moments  : mean / covariance of ProcessModel::sample() (include/ProcessModel.hpp:125-150) for a fixed
                    previous pose, to first order in the input noise (what the statistical parity test checks)

Pinned against the reference's own step() functions (oracle/_ref, phd_ref_odometry2d_step /
phd_ref_ackerman2d_step) through tests/golden/motion_steps.npz.
"""
import numpy as np


def odometry2d_step(pose, u):
    pose = np.asarray(pose, dtype=np.float64)
    x, y, th = pose[..., 0], pose[..., 1], pose[..., 2]
    ct, st = np.cos(th), np.sin(th)
    cd, sd = np.cos(u[2]), np.sin(u[2])
    # p_k = p_km + C_km^T dp with C_km = [[ct, st], [-st, ct]]
    xn = x + ct * u[0] - st * u[1]
    yn = y + st * u[0] + ct * u[1]
    # C_k = C(dtheta) C(theta); theta_k = atan2(C_k(0,1), C_k(0,0))
    c00 = cd * ct - sd * st
    c01 = cd * st + sd * ct
    return np.stack([xn, yn, np.arctan2(c01, c00)], axis=-1)


def ackerman2d_step(pose, u, dt, h, l, dx, dy):
    pose = np.asarray(pose, dtype=np.float64)
    x, y, r = pose[..., 0], pose[..., 1], pose[..., 2]
    cr, sr = np.cos(r), np.sin(r)
    tu = np.tan(u[1])
    v = u[0] / (1 - tu * h / l)
    xn = x + dt * (v * cr - v / l * tu * (dx * sr + dy * cr))
    yn = y + dt * (v * sr + v / l * tu * (dx * cr - dy * sr))
    rn = r + dt * v / l * tu
    rn = np.where(rn > np.pi, rn - 2 * np.pi, np.where(rn < -np.pi, rn + 2 * np.pi, rn))
    return np.stack([xn, yn, rn], axis=-1)


def numeric_jacobian(f, u, eps=1e-7):
    u = np.asarray(u, dtype=np.float64)
    f0 = f(u)
    J = np.zeros((len(f0), len(u)))
    for k in range(len(u)):
        d = np.zeros(len(u)); d[k] = eps
        J[:, k] = (f(u + d) - f(u - d)) / (2 * eps)
    return f0, J

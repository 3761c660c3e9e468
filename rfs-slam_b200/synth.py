"""Deterministic synthetic 2-D range-bearing workloads for the PHD update path.

Follows SURVEY.md §8(d): RngBrg model with the thresholds of the reference's
cfg/rbphdslam2dSim.xml:46-80, a "dense" world (every component in range) and a "sparse"
one, particles scattered about the true pose, per-particle Gaussian mixtures scattered about
the landmarks, 80 % landmark measurements + 20 % clutter.  All arrays are fp64 and are fed
unchanged to every arm (CUDA path, oracle restatement, compiled reference).

numpy's PCG64 stream + explicit Box-Muller is used so that the data do not depend on
numpy's normal() implementation.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

# ---- the descriptors mirrored by include/rfsb200.h ---------------------------------------

DEFAULT_MODEL = dict(
    model_id=1,
    R=(5e-3, 0.0, 0.0, 5e-4),  # cfg varzr/varzb x 10 inflation (cfg/rbphdslam2dSim.xml:37-38,50)
    Pd=0.99,
    clutter_intensity=1e-4,
    range_min=0.5,
    range_max=10.0,
    range_buffer=0.25,
    innov_thr_range=1.0,
    innov_thr_bearing=0.2,
)

DEFAULT_CFG = dict(
    birth_gaussian_weight=0.01,
    new_gaussian_create_innov_md_threshold=3.0,
    eval_point_gaussian_weight=0.75,
    meas_likelihood_md_threshold=3.0,
    merging_threshold=0.5,
    merging_cov_inflation_factor=1.5,
    pruning_threshold=0.01,
    eval_point_count=15,
    use_cluster_process=0,
    assignment_sum_method=0,
)


def clutter_integral(model: dict) -> float:
    """MeasurementModel_RngBrg::clutterIntensityIntegral (src/MeasurementModel_RngBrg.cpp:175-178)."""
    return model["clutter_intensity"] * 2.0 * math.acos(-1.0) * (model["range_max"] - model["range_min"])


# BASELINE.json configs that are synthetic shapes (C1/C5 are driver runs, not shapes)
CONFIGS = {
    "C2": dict(N=1000, nM=100, nZ=20, use_cluster_process=0, config_id=2),
    "C3": dict(N=8000, nM=200, nZ=30, use_cluster_process=1, config_id=3),
    "C4": dict(N=64000, nM=200, nZ=30, use_cluster_process=1, config_id=4),
}


@dataclass
class Workload:
    count: np.ndarray      # [N] int32
    mean: np.ndarray       # [sum,2]
    cov: np.ndarray        # [sum,3] (xx,xy,yy)
    w: np.ndarray          # [sum]
    pose: np.ndarray       # [N,3]
    pose_cov: np.ndarray   # [6] shared upper triangle of the 3x3 pose covariance
    weight: np.ndarray     # [N]
    Z: np.ndarray          # [nZ,2]
    model: dict = field(default_factory=dict)
    cfg: dict = field(default_factory=dict)
    landmarks: np.ndarray | None = None

    @property
    def N(self) -> int:
        return int(self.count.shape[0])

    @property
    def nZ(self) -> int:
        return int(self.Z.shape[0])

    @property
    def offsets(self) -> np.ndarray:
        o = np.zeros(self.N + 1, dtype=np.int64)
        np.cumsum(self.count, out=o[1:])
        return o

    def shard(self, rank: int, world: int) -> "Workload":
        """Contiguous block partition of particle indices."""
        lo = rank * self.N // world
        hi = (rank + 1) * self.N // world
        off = self.offsets
        a, b = int(off[lo]), int(off[hi])
        return Workload(self.count[lo:hi].copy(), self.mean[a:b].copy(), self.cov[a:b].copy(),
                        self.w[a:b].copy(), self.pose[lo:hi].copy(), self.pose_cov.copy(),
                        self.weight[lo:hi].copy(), self.Z.copy(), dict(self.model), dict(self.cfg),
                        self.landmarks)


class _Rng:
    """uniform doubles from PCG64 + explicit Box-Muller."""

    def __init__(self, seed: int):
        self.g = np.random.Generator(np.random.PCG64(seed))

    def uniform(self, lo, hi, size):
        return lo + (hi - lo) * self.g.random(size)

    def normal(self, size):
        size = tuple(np.atleast_1d(size))
        u1 = 1.0 - self.g.random(size)  # (0,1]
        u2 = self.g.random(size)
        return np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * math.pi * u2)


def make_workload(N: int, nM: int, nZ: int, *, use_cluster_process: int = 1, world: str = "dense",
                  config_id: int = 0, parity_extras: bool = False, ragged: float = 0.0,
                  model: dict | None = None, cfg: dict | None = None, seed: int | None = None,
                  shard_id: int = 0) -> Workload:
    md = dict(DEFAULT_MODEL)
    if model:
        md.update(model)
    fc = dict(DEFAULT_CFG)
    fc["use_cluster_process"] = int(use_cluster_process)
    if cfg:
        fc.update(cfg)
    md["clutter_integral"] = clutter_integral(md)
    rng = _Rng(0xB2000000 + config_id if seed is None else seed)
    rmin, rmax, buf = md["range_min"], md["range_max"], md["range_buffer"]

    # ---- landmarks (true pose x* = (0,0,0)) ----
    if world == "dense":
        r_lo, r_hi = rmin + buf, rmax - buf
    elif world == "sparse":
        r_lo, r_hi = 0.0, 3.0 * rmax
    elif world == "clumped":
        r_lo, r_hi = rmin + 4 * buf, rmax - 4 * buf
    else:
        raise ValueError(world)
    u = rng.uniform(0.0, 1.0, nM)
    r = np.sqrt(r_lo ** 2 + u * (r_hi ** 2 - r_lo ** 2))  # uniform in area
    b = rng.uniform(-math.pi, math.pi, nM)
    if world == "clumped":
        # landmarks in tight groups of 6 (0.1 m): gates overlap, so the (eval point, measurement)
        # graph has large connected partitions (the reference's Murty branch, nR + nC > 8)
        g = np.arange(nM) // 6
        r = r[g * 6] + rng.uniform(-0.1, 0.1, nM)
        b = b[g * 6] + rng.uniform(-0.1, 0.1, nM) / np.maximum(r, 1.0)
    if parity_extras and nM >= 20:
        # 5 % of the landmarks in the four buffer bands (Q2), two at bearing pi +- 0.01 (Q3)
        k = max(4, nM // 20)
        bands = [(rmin - buf, rmin), (rmin, rmin + buf), (rmax - buf, rmax), (rmax, rmax + buf)]
        for j in range(k):
            lo, hi = bands[j % 4]
            r[j] = max(1e-3, rng.uniform(lo + 0.03, hi - 0.03, 1)[0])
        b[k] = math.pi - 0.01
        b[k + 1] = -math.pi + 0.01
        r[k] = 0.5 * (r_lo + r_hi)
        r[k + 1] = 0.4 * (r_lo + r_hi)
    lmk = np.stack([r * np.cos(b), r * np.sin(b)], axis=1)

    # ---- measurements (shared) ----
    n_real = int(math.floor(0.8 * nZ))
    rl = np.hypot(lmk[:, 0], lmk[:, 1])
    in_range = np.nonzero((rl >= rmin) & (rl <= rmax))[0]
    n_real = min(n_real, len(in_range))
    perm = rng.g.permutation(len(in_range))[:n_real]
    sel = in_range[perm]
    if parity_extras and nM >= 20 and n_real >= 2:
        k = max(4, nM // 20)
        sel[0], sel[1] = k, k + 1  # the two landmarks at bearing ~ +-pi are observed
    zr = rl[sel] + math.sqrt(5e-4) * rng.normal(n_real)
    zb = np.arctan2(lmk[sel, 1], lmk[sel, 0]) + math.sqrt(5e-5) * rng.normal(n_real)
    zb = (zb + math.pi) % (2 * math.pi) - math.pi
    n_cl = nZ - n_real
    cr = rng.uniform(rmin, rmax, n_cl)
    cb = rng.uniform(-math.pi, math.pi, n_cl)
    Z = np.stack([np.concatenate([zr, cr]), np.concatenate([zb, cb])], axis=1)

    # ---- particles (their own stream, so that shards of one job share the world and Z) ----
    rng = _Rng((0xB2000000 + config_id if seed is None else seed) + 7919 * (1 + shard_id))
    pose = rng.normal((N, 3)) * np.array([0.05, 0.05, 0.01])
    pose_cov = np.array([3e-5, 0.0, 0.0, 3e-5, 0.0, 3e-5])
    weight = np.ones(N)

    # ---- per-particle Gaussian mixtures ----
    mean = lmk[None, :, :] + 0.03 * rng.normal((N, nM, 2))
    a = rng.uniform(1e-3, 1e-2, (N, nM))
    bb = rng.uniform(1e-3, 1e-2, (N, nM))
    phi = rng.uniform(0.0, math.pi, (N, nM))
    c, s = np.cos(phi), np.sin(phi)
    cov = np.stack([c * c * a + s * s * bb, c * s * (a - bb), s * s * a + c * c * bb], axis=2)
    w = rng.uniform(0.3, 1.0, (N, nM))
    light = rng.uniform(0.0, 1.0, (N, nM)) < 0.10
    w[light] = 0.01
    keep = np.ones((N, nM), dtype=bool)
    if ragged > 0:
        keep = rng.uniform(0.0, 1.0, (N, nM)) >= ragged
    if parity_extras:
        empty = rng.uniform(0.0, 1.0, N) < 0.01  # Q10: empty-map particles
        if N >= 4:
            empty[N // 3] = True
        keep[empty, :] = False
    count = keep.sum(axis=1).astype(np.int32)
    mean = mean[keep]
    cov = cov[keep]
    w = w[keep]

    return Workload(count=np.ascontiguousarray(count), mean=np.ascontiguousarray(mean),
                    cov=np.ascontiguousarray(cov), w=np.ascontiguousarray(w),
                    pose=np.ascontiguousarray(pose), pose_cov=pose_cov, weight=weight,
                    Z=np.ascontiguousarray(Z), model=md, cfg=fc, landmarks=lmk)


def make_config(name: str, **kw) -> Workload:
    c = dict(CONFIGS[name])
    c.update(kw)
    return make_workload(**c)

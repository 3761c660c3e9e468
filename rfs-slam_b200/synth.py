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
    pose_cov: np.ndarray | None   # [6] shared upper triangle of the 3x3 pose covariance (None: zero)
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
    def dim(self) -> int:
        """landmark / measurement dimension: 2 (RngBrg) or 3 (VictoriaPark)"""
        return int(self.mean.shape[1]) if self.mean.ndim == 2 else 2

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
                        self.w[a:b].copy(), self.pose[lo:hi].copy(),
                        None if self.pose_cov is None else self.pose_cov.copy(),
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
    if c.pop("vp", False):
        return make_vp_workload(**c)
    return make_workload(**c)


# ---- Victoria Park plugin set (BASELINE config 5): 3-D landmarks (x, y, diameter) -----------------------

# cfg/rbphdslam_VictoriaPark_artificialClutter.xml:40-60,66-118 (R = diag(varzr, varzb, varzd) x the
# measurement-noise inflation 40, src/rbphdslam_VictoriaPark.cpp:366-371)
DEFAULT_MODEL_VP = dict(
    model_id=2,
    R=(1.0, 0.0, 0.0, 0.0, 1e-3, 0.0, 0.0, 0.0, 0.08),
    Slb=1e-5,                                   # varza
    pd_table=(0.00, 0.05, 0.35, 0.76, 0.89, 0.90),
    buffer_zone_pd=0.4,
    range_min=5.0,
    range_max=70.0,
    bearing_min=6.3025 * math.pi / 180.0,
    bearing_max=177.0 * math.pi / 180.0,
    expected_clutter=6.0,
    innov_thr_range=7.5,
    innov_thr_bearing=0.2,
    Pd=0.0, range_buffer=0.0,                   # unused by this model
)

DEFAULT_CFG_VP = dict(DEFAULT_CFG, merging_threshold=1.0)

CONFIGS["C5"] = dict(vp=True, N=4000, nM=150, nZ=12, config_id=5)


def vp_clutter_intensity(expected_clutter: float, scan: np.ndarray) -> float:
    """MeasurementModel_VictoriaPark::setLaserScan (src/MeasurementModel_VictoriaPark.cpp:268-281)."""
    area = 0.0
    for i in range(1, len(scan)):
        area += scan[i] * scan[i - 1]
    area += scan[0] * scan[len(scan) - 1]
    area *= math.sin(math.acos(-1.0) / 360) / 2
    return expected_clutter / area


def make_vp_workload(N: int, nM: int, nZ: int, *, use_cluster_process: int = 0, config_id: int = 5,
                     ragged: float = 0.0, parity_extras: bool = False, model: dict | None = None,
                     cfg: dict | None = None, seed: int | None = None, shard_id: int = 0) -> Workload:
    """Synthetic park: trees (x, y, diameter) in front of the vehicle, a 720-beam half-degree lidar scan
    with returns on the trees, a few occluders and 'no return' beams, 80 % tree measurements + 20 %
    clutter.  mean [.,3], cov [.,6] (xx,xy,xz,yy,yz,zz), Z [nZ,3]; every array holds values that are
    exactly representable in fp32, so the fp32 device state equals the oracle's input."""
    md = dict(DEFAULT_MODEL_VP)
    if model:
        md.update(model)
    fc = dict(DEFAULT_CFG_VP)
    fc["use_cluster_process"] = int(use_cluster_process)
    if cfg:
        fc.update(cfg)
    base = 0xB2000000 + config_id if seed is None else seed
    rng = _Rng(base)
    f32 = lambda a: np.asarray(a, dtype=np.float32).astype(np.float64)
    rmin, rmax = md["range_min"], md["range_max"]
    th_true = math.pi / 2 + 0.1                    # sensor frame = theta - pi/2
    th_s = th_true - math.pi / 2

    # ---- trees about the true pose (0, 0, th_true): some beyond the range / bearing limits ----
    u = rng.uniform(0.0, 1.0, nM)
    r = np.sqrt((rmin - 2.0) ** 2 + u * ((rmax + 8.0) ** 2 - (rmin - 2.0) ** 2))
    b = rng.uniform(-0.15, math.pi + 0.15, nM)   # bearing in the sensor frame
    D = rng.uniform(0.25, 1.6, nM)
    if parity_extras and nM >= 12:
        r[0], b[0] = rmax - 0.05, 1.0            # just inside / outside the range limits
        r[1], b[1] = rmax + 0.05, 1.3
        r[2], b[2] = rmin + 0.05, 2.0
        r[3], b[3] = 30.0, md["bearing_min"] + 0.004   # straddling the bearing limits
        r[4], b[4] = 30.0, md["bearing_max"] - 0.004
        r[5], b[5], D[5] = 6.0, 0.9, 1.6         # wide trunk close by (many beams)
        r[6], b[6], D[6] = 65.0, 1.7, 0.3        # thin trunk far away (no beam: table entry 0)
    lmk = np.stack([r * np.cos(b + th_s), r * np.sin(b + th_s), D], axis=1)

    # ---- lidar scan: 720 half-degree beams; index k <-> bearing k * 2 pi / 720 ----
    scan = np.full(720, 80.0)
    scan[rng.uniform(0.0, 1.0, 720) < 0.10] = 0.0                      # no return
    visible = np.zeros(nM, dtype=bool)
    order = np.argsort(-r)                                             # nearer trees overwrite farther ones
    for j in order:
        if not (rmin <= r[j] <= rmax):
            continue
        g = math.atan(D[j] / 2 / r[j])
        k0, k1 = math.ceil((b[j] - g) * 720 / (2 * math.pi)), math.floor((b[j] + g) * 720 / (2 * math.pi))
        for k in range(k0, k1 + 1):
            scan[k % 720] = r[j] - D[j] / 2 * 0.9
    occl = rng.uniform(0.0, 1.0, 720) < 0.08                           # occluders in front of some trees
    scan[occl] = rng.uniform(3.0, 25.0, int(occl.sum()))
    scan = f32(scan)
    for j in range(nM):
        if rmin <= r[j] <= rmax and md["bearing_min"] <= b[j] <= md["bearing_max"]:
            k = int(round(b[j] * 720 / (2 * math.pi))) % 720
            visible[j] = scan[k] == 0.0 or scan[k] > r[j] - D[j] / 2 - 0.18
    md["scan"] = [float(v) for v in scan]
    md["clutter_integral"] = md["expected_clutter"]
    md["clutter_intensity"] = vp_clutter_intensity(md["expected_clutter"], scan)

    # ---- measurements (shared): visible trees + clutter ----
    vis = np.nonzero(visible)[0]
    n_real = min(int(math.floor(0.8 * nZ)), len(vis))
    sel = vis[rng.g.permutation(len(vis))[:n_real]]
    if parity_extras and nM >= 12 and n_real >= 2 and visible[5]:
        sel[0] = 5
    zr = r[sel] + 0.15 * rng.normal(n_real)
    zb = b[sel] + 0.005 * rng.normal(n_real)
    zd = D[sel] + 0.05 * rng.normal(n_real)
    n_cl = nZ - n_real
    Z = np.stack([np.concatenate([zr, rng.uniform(rmin, rmax, n_cl)]),
                  np.concatenate([zb, rng.uniform(md["bearing_min"], md["bearing_max"], n_cl)]),
                  np.concatenate([zd, rng.uniform(0.2, 1.5, n_cl)])], axis=1)
    Z = f32(Z)

    # ---- particles ----
    rng = _Rng(base + 7919 * (1 + shard_id))
    pose = np.array([0.0, 0.0, th_true]) + rng.normal((N, 3)) * np.array([0.08, 0.08, 0.004])
    pose = f32(pose)
    weight = np.ones(N)

    # ---- per-particle mixtures: random SPD 3x3 covariances (xy block rotated, small xz / yz terms) ----
    mean = lmk[None, :, :] + rng.normal((N, nM, 3)) * np.array([0.10, 0.10, 0.03])
    mean[:, :, 2] = np.maximum(mean[:, :, 2], 0.15)
    a = rng.uniform(0.01, 0.12, (N, nM))
    bb = rng.uniform(0.01, 0.12, (N, nM))
    phi = rng.uniform(0.0, math.pi, (N, nM))
    c, s_ = np.cos(phi), np.sin(phi)
    pxx, pxy, pyy = c * c * a + s_ * s_ * bb, c * s_ * (a - bb), s_ * s_ * a + c * c * bb
    pdd = rng.uniform(1e-3, 1e-2, (N, nM))
    k1 = rng.uniform(-0.3, 0.3, (N, nM))
    k2 = rng.uniform(-0.3, 0.3, (N, nM))
    pxd, pyd = k1 * np.sqrt(pxx * pdd), k2 * np.sqrt(pyy * pdd)
    cov = np.stack([pxx, pxy, pxd, pyy, pyd, pdd], axis=2)
    w = rng.uniform(0.3, 1.0, (N, nM))
    light = rng.uniform(0.0, 1.0, (N, nM)) < 0.10
    w[light] = 0.015625   # 2^-6: exact in fp32 and clear of the prune threshold 0.01
    keep = np.ones((N, nM), dtype=bool)
    if ragged > 0:
        keep = rng.uniform(0.0, 1.0, (N, nM)) >= ragged
    if parity_extras:
        empty = rng.uniform(0.0, 1.0, N) < 0.01
        if N >= 4:
            empty[N // 3] = True
        keep[empty, :] = False
    count = keep.sum(axis=1).astype(np.int32)
    return Workload(count=np.ascontiguousarray(count), mean=np.ascontiguousarray(f32(mean[keep])),
                    cov=np.ascontiguousarray(f32(cov[keep])), w=np.ascontiguousarray(f32(w[keep])),
                    pose=np.ascontiguousarray(pose), pose_cov=None, weight=weight,
                    Z=np.ascontiguousarray(Z), model=md, cfg=fc, landmarks=lmk)

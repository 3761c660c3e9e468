"""ctypes binding of the CPU oracle — TEST INFRASTRUCTURE, NOT PRODUCT.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  Two libraries share one entry-point signature (oracle/phd_oracle.h):

  libphd_oracle.so        phd_oracle_update  — our fp64 restatement (oracle/phd_oracle.cpp)
  _ref/libphd_ref.so      phd_ref_update     — the reference's own sources compiled against
                                               oracle/compat shims (oracle/ref_harness.cpp)
"""
from __future__ import annotations

import ctypes as C
import importlib
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

ORACLE_LIB = os.path.join(_HERE, "libphd_oracle.so")
REF_LIB = os.path.join(_HERE, "_ref", "libphd_ref.so")

STAGE_UPDATE_MAP, STAGE_WEIGHTING, STAGE_MERGE, STAGE_FULL = 1, 2, 3, 4
SORT_STD, SORT_STABLE = 0, 1


def _capi():
    return importlib.import_module("rfs_slam_b200.capi")


class PhdIO(C.Structure):
    _fields_ = [
        ("N", C.c_int32), ("count_in", C.c_void_p), ("mean_in", C.c_void_p), ("cov_in", C.c_void_p),
        ("w_in", C.c_void_p), ("pose", C.c_void_p), ("pose_cov", C.c_void_p),
        ("pose_cov_mode", C.c_int32), ("weight_in", C.c_void_p), ("Z", C.c_void_p), ("nZ", C.c_int32),
        ("model", C.c_void_p), ("cfg", C.c_void_p), ("sort_mode", C.c_int32), ("stage", C.c_int32),
        ("n_threads", C.c_int32),
        ("cap_total", C.c_int64), ("count_out", C.c_void_p), ("mean_out", C.c_void_p),
        ("cov_out", C.c_void_p), ("w_out", C.c_void_p), ("wprev_out", C.c_void_p),
        ("weight_out", C.c_void_p), ("unused_mask", C.c_void_p), ("n_in_fov", C.c_void_p),
        ("flags", C.c_void_p), ("elapsed_s", C.c_double),
    ]


_libs = {}


def _load(which: str):
    if which in _libs:
        return _libs[which]
    path = ORACLE_LIB if which == "oracle" else REF_LIB
    if not os.path.exists(path):
        raise FileNotFoundError(path)
    lib = C.CDLL(path)
    fn = getattr(lib, "phd_oracle_update" if which == "oracle" else "phd_ref_update")
    fn.restype = C.c_int
    fn.argtypes = [C.POINTER(PhdIO)]
    for name in ("phd_oracle_update_vp", "phd_ref_update_vp"):
        if hasattr(lib, name):
            getattr(lib, name).restype = C.c_int
            getattr(lib, name).argtypes = [C.POINTER(PhdIO)]
    for name in ("phd_oracle_vp_pd", "phd_ref_vp_pd"):
        if hasattr(lib, name):
            getattr(lib, name).restype = C.c_double
            getattr(lib, name).argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    if which == "oracle":
        lib.phd_oracle_permanent.restype = C.c_double
        lib.phd_oracle_permanent.argtypes = [C.c_void_p, C.c_int]
        lib.phd_oracle_lexi_count.restype = C.c_int64
        lib.phd_oracle_lexi_count.argtypes = [C.c_int, C.c_int]
        lib.phd_oracle_partition_likelihood.restype = C.c_double
        lib.phd_oracle_partition_likelihood.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                                        C.c_void_p, C.c_void_p]
        lib.phd_oracle_partition.restype = C.c_int
        lib.phd_oracle_partition.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.phd_oracle_murty_sum.restype = C.c_double
        lib.phd_oracle_murty_sum.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    else:
        for name in ("phd_ref_permanent",):
            if hasattr(lib, name):
                getattr(lib, name).restype = C.c_double
                getattr(lib, name).argtypes = [C.c_void_p, C.c_int]
    _libs[which] = (lib, fn)
    return _libs[which]


def have_ref() -> bool:
    return os.path.exists(REF_LIB)


def have_oracle() -> bool:
    return os.path.exists(ORACLE_LIB)


class Result:
    """Packed posterior state returned by an oracle run."""

    def __init__(self, count, mean, cov, w, wprev, weight, unused_mask, n_in_fov, flags, elapsed_s):
        self.count, self.mean, self.cov, self.w, self.wprev = count, mean, cov, w, wprev
        self.weight, self.unused_mask, self.n_in_fov, self.flags = weight, unused_mask, n_in_fov, flags
        self.elapsed_s = elapsed_s

    @property
    def offsets(self):
        o = np.zeros(len(self.count) + 1, dtype=np.int64)
        np.cumsum(self.count, out=o[1:])
        return o

    def particle(self, i):
        o = self.offsets
        a, b = int(o[i]), int(o[i + 1])
        return self.mean[a:b], self.cov[a:b], self.w[a:b]


def run(wl, *, which: str = "oracle", stage: int = STAGE_FULL, sort_mode: int = SORT_STABLE,
        n_threads: int = 0, cap_factor: float = 2.0) -> Result:
    """Run one RBPHDFilter::update() on a synth.Workload through the oracle or the reference."""
    capi = _capi()
    lib, fn = _load(which)
    N = wl.N
    D = wl.dim                      # 2: RngBrg, 3: VictoriaPark
    NC = D * (D + 1) // 2
    if D == 3:
        fn = lib.phd_oracle_update_vp if which == "oracle" else lib.phd_ref_update_vp
    total_in = int(wl.count.sum())
    cap_total = int(total_in * cap_factor) + 64 * N + 1024
    md = capi.model_desc(wl.model)
    fc = capi.filter_cfg(wl.cfg)
    count_in = np.ascontiguousarray(wl.count, dtype=np.int32)
    arrs = dict(mean_in=np.ascontiguousarray(wl.mean, dtype=np.float64),
                cov_in=np.ascontiguousarray(wl.cov, dtype=np.float64),
                w_in=np.ascontiguousarray(wl.w, dtype=np.float64),
                pose=np.ascontiguousarray(wl.pose, dtype=np.float64),
                weight_in=np.ascontiguousarray(wl.weight, dtype=np.float64),
                Z=np.ascontiguousarray(wl.Z, dtype=np.float64).reshape(-1, D))
    pose_cov = None if wl.pose_cov is None else np.ascontiguousarray(wl.pose_cov, dtype=np.float64)
    if pose_cov is None:
        mode = 0
    elif pose_cov.ndim == 1:
        mode = 1
    else:
        mode = 2
    out = dict(count_out=np.zeros(N, np.int32), mean_out=np.zeros((cap_total, D)),
               cov_out=np.zeros((cap_total, NC)), w_out=np.zeros(cap_total), wprev_out=np.zeros(cap_total),
               weight_out=np.zeros(N), unused_mask=np.zeros(N, np.uint64), n_in_fov=np.zeros(N, np.int32),
               flags=np.zeros(N, np.int32))
    io = PhdIO()
    io.N = N
    io.count_in = count_in.ctypes.data
    for k, a in arrs.items():
        setattr(io, k, a.ctypes.data)
    io.pose_cov = None if pose_cov is None else pose_cov.ctypes.data
    io.pose_cov_mode = mode
    io.nZ = arrs["Z"].shape[0]
    io.model = C.addressof(md)
    io.cfg = C.addressof(fc)
    io.sort_mode = sort_mode
    io.stage = stage
    io.n_threads = n_threads
    io.cap_total = cap_total
    for k, a in out.items():
        setattr(io, k, a.ctypes.data)
    rc = fn(C.byref(io))
    if rc == -4 and cap_total < total_in * (1 + io.nZ) + N:
        # output arrays too small (few Gaussians, many measurements): every Gaussian can spawn one per measurement
        return run(wl, which=which, stage=stage, sort_mode=sort_mode, n_threads=n_threads,
                   cap_factor=float(2 + io.nZ))
    if rc != 0:
        raise RuntimeError(f"{which} update failed rc={rc}")
    tot = int(out["count_out"].sum())
    return Result(out["count_out"], out["mean_out"][:tot].copy(), out["cov_out"][:tot].copy(),
                  out["w_out"][:tot].copy(), out["wprev_out"][:tot].copy(), out["weight_out"],
                  out["unused_mask"], out["n_in_fov"], out["flags"], io.elapsed_s)


def vp_pd(model: dict, pose, lx, lcov6, which: str = "oracle"):
    """MeasurementModel_VictoriaPark::probabilityOfDetection for one landmark -> (Pd, close)."""
    capi = _capi()
    lib, _ = _load(which)
    md = capi.model_desc(model)
    pose = np.ascontiguousarray(pose, dtype=np.float64)
    lx = np.ascontiguousarray(lx, dtype=np.float64)
    lc = np.ascontiguousarray(lcov6, dtype=np.float64)
    close = C.c_int(0)
    f = lib.phd_oracle_vp_pd if which == "oracle" else lib.phd_ref_vp_pd
    pd = f(C.addressof(md), pose.ctypes.data, lx.ctypes.data, lc.ctypes.data, C.addressof(close))
    return float(pd), bool(close.value)


class PhdBirthIO(C.Structure):
    _fields_ = [
        ("model", C.c_void_p), ("N", C.c_int32), ("nZ", C.c_int32), ("cand_cap", C.c_int32), ("add_cap", C.c_int32),
        ("resample_occurred", C.c_int32), ("count_thr", C.c_uint32), ("check_thr", C.c_uint32),
        ("cur_count_thr", C.c_uint32), ("support_dist", C.c_double),
        ("pose", C.c_void_p), ("pose_cov", C.c_void_p), ("Z", C.c_void_p), ("parent", C.c_void_p),
        ("unused", C.c_void_p), ("nfov", C.c_void_p), ("cand_n", C.c_void_p), ("cand_mean", C.c_void_p),
        ("cand_cov", C.c_void_p), ("cand_support", C.c_void_p), ("cand_checks", C.c_void_p),
        ("add_n", C.c_void_p), ("add_mean", C.c_void_p), ("add_cov", C.c_void_p),
    ]


class BirthState:
    """Candidate lists of all particles (birthGaussians_, include/RBPHDFilter.hpp:255): n [N], mean [N][cap][D],
    cov [N][cap][NC] (upper triangle), support / checks [N][cap]."""

    def __init__(self, N: int, D: int, cap: int = 32):
        self.N, self.D, self.cap = N, D, cap
        NC = D * (D + 1) // 2
        self.n = np.zeros(N, np.int32)
        self.mean = np.zeros((N, cap, D))
        self.cov = np.zeros((N, cap, NC))
        self.support = np.zeros((N, cap), np.int32)
        self.checks = np.zeros((N, cap), np.int32)

    def copy(self):
        o = BirthState(self.N, self.D, self.cap)
        o.n, o.mean, o.cov, o.support, o.checks = (self.n.copy(), self.mean.copy(), self.cov.copy(),
                                                     self.support.copy(), self.checks.copy())
        return o


def birth_candidates(model: dict, bcfg: dict, state: BirthState, pose, Z, unused, nfov, *, parent=None, pose_cov=None,
                     which: str = "oracle", add_cap: int = 96):
    """addBirthGaussians() in its candidate-list form (include/RBPHDFilter.hpp:1000-1080) for all particles, in place
    on `state`.  bcfg: count_thr / check_thr / cur_count_thr / support_dist.  parent != None <=> resampleOccured_.
    Returns (add_n [N], add_mean [N][add_cap][D], add_cov [N][add_cap][NC]): the Gaussians that became real."""
    capi = _capi()
    lib, _ = _load(which)
    D = state.D
    NC = D * (D + 1) // 2
    N = state.N
    md = capi.model_desc(model)
    pose = np.ascontiguousarray(pose, dtype=np.float64)
    Z = np.ascontiguousarray(Z, dtype=np.float64).reshape(-1, D)
    un = np.ascontiguousarray(unused, dtype=np.uint64).copy()
    nf = np.ascontiguousarray(nfov, dtype=np.int32)
    par = None if parent is None else np.ascontiguousarray(parent, dtype=np.int32)
    pc = None if pose_cov is None else np.ascontiguousarray(pose_cov, dtype=np.float64)
    add_n = np.zeros(N, np.int32)
    add_mean = np.zeros((N, add_cap, D))
    add_cov = np.zeros((N, add_cap, NC))
    io = PhdBirthIO()
    io.model = C.addressof(md)
    io.N, io.nZ, io.cand_cap, io.add_cap = N, Z.shape[0], state.cap, add_cap
    io.resample_occurred = 0 if par is None else 1
    io.count_thr, io.check_thr, io.cur_count_thr = bcfg["count_thr"], bcfg["check_thr"], bcfg["cur_count_thr"]
    io.support_dist = bcfg["support_dist"]
    io.pose, io.Z, io.unused, io.nfov = pose.ctypes.data, Z.ctypes.data, un.ctypes.data, nf.ctypes.data
    io.pose_cov = None if pc is None else pc.ctypes.data
    io.parent = None if par is None else par.ctypes.data
    io.cand_n, io.cand_mean, io.cand_cov = state.n.ctypes.data, state.mean.ctypes.data, state.cov.ctypes.data
    io.cand_support, io.cand_checks = state.support.ctypes.data, state.checks.ctypes.data
    io.add_n, io.add_mean, io.add_cov = add_n.ctypes.data, add_mean.ctypes.data, add_cov.ctypes.data
    if which == "oracle":
        fn = lib.phd_oracle_birth_candidates
    else:
        fn = lib.phd_ref_birth_candidates_vp if D == 3 else lib.phd_ref_birth_candidates
    fn.restype = C.c_int
    fn.argtypes = [C.POINTER(PhdBirthIO)]
    rc = fn(C.byref(io))
    if rc != 0:
        raise RuntimeError(f"{which} birth_candidates failed rc={rc}")
    return add_n, add_mean, add_cov


def permanent(A: np.ndarray, which: str = "oracle") -> float:
    lib, _ = _load(which)
    A = np.ascontiguousarray(A, dtype=np.float64)
    f = lib.phd_oracle_permanent if which == "oracle" else lib.phd_ref_permanent
    return float(f(A.ctypes.data, A.shape[0]))


def lexi_count(nM: int, nZ: int) -> int:
    lib, _ = _load("oracle")
    return int(lib.phd_oracle_lexi_count(nM, nZ))


def partition_likelihood(L: np.ndarray, evalPd: np.ndarray, clutter: np.ndarray):
    lib, _ = _load("oracle")
    L = np.ascontiguousarray(L, dtype=np.float64)
    evalPd = np.ascontiguousarray(evalPd, dtype=np.float64)
    clutter = np.ascontiguousarray(clutter, dtype=np.float64)
    flags = np.zeros(1, np.int32)
    nE, nZ = L.shape
    v = lib.phd_oracle_partition_likelihood(L.ctypes.data, nE, nZ, evalPd.ctypes.data,
                                            clutter.ctypes.data, flags.ctypes.data)
    return float(v), int(flags[0])


def murty_sum(Lp: np.ndarray, rowPd: np.ndarray, colClutter: np.ndarray) -> float:
    lib, _ = _load("oracle")
    Lp = np.ascontiguousarray(Lp, dtype=np.float64)
    rowPd = np.ascontiguousarray(rowPd, dtype=np.float64)
    colClutter = np.ascontiguousarray(colClutter, dtype=np.float64)
    return float(lib.phd_oracle_murty_sum(Lp.ctypes.data, Lp.shape[0], Lp.shape[1],
                                          rowPd.ctypes.data, colClutter.ctypes.data))


def partition(L: np.ndarray):
    """(nP, nRows[nP], nCols[nP], isZero[nP]) as the reference's caller sees them (Q6)."""
    lib, _ = _load("oracle")
    L = np.ascontiguousarray(L, dtype=np.float64)
    nR, nC = L.shape
    a = np.zeros(128, np.int32); b = np.zeros(128, np.int32); z = np.zeros(128, np.int32)
    nP = lib.phd_oracle_partition(L.ctypes.data, nR, nC, a.ctypes.data, b.ctypes.data, z.ctypes.data)
    return nP, a[:nP].copy(), b[:nP].copy(), z[:nP].copy()


# ---- scripted sequences through the PUBLIC filter API (oracle/seq_harness.cpp) --------------------
SEQ_REF_LIB = os.path.join(_HERE, "_ref", "libseq_ref.so")
SEQ_B200_LIB = os.path.join(_HERE, "_ref", "libseq_b200.so")


class SeqIO(C.Structure):
    _fields_ = [
        ("N", C.c_int32), ("n_steps", C.c_int32), ("nZ_max", C.c_int32),
        ("poses", C.c_void_p), ("pose_cov", C.c_void_p), ("Z", C.c_void_p), ("nZ", C.c_void_p),
        ("R", C.c_void_p), ("Q_lmk", C.c_void_p),
        ("Pd", C.c_double), ("clutter", C.c_double), ("range_min", C.c_double), ("range_max", C.c_double),
        ("range_buffer", C.c_double), ("thr_r", C.c_double), ("thr_b", C.c_double),
        ("birth_w", C.c_double), ("gate", C.c_double), ("eval_w", C.c_double), ("wl_gate", C.c_double),
        ("merge_t", C.c_double), ("merge_f", C.c_double), ("prune_t", C.c_double),
        ("n_eval", C.c_int32), ("use_sc", C.c_int32), ("min_updates_before_resample", C.c_int32),
        ("neff_threshold", C.c_double), ("precision", C.c_int32), ("seed48", C.c_uint32),
        ("cap_total", C.c_int64), ("count_out", C.c_void_p), ("mean_out", C.c_void_p), ("cov_out", C.c_void_p),
        ("w_out", C.c_void_p), ("weight_out", C.c_void_p), ("n_resampled", C.c_void_p), ("gm_size_trace", C.c_void_p),
        ("birth_count_thr", C.c_uint32), ("birth_check_thr", C.c_uint32), ("birth_cur_thr", C.c_uint32), ("birth_reserved", C.c_uint32),
        ("birth_support_dist", C.c_double),
    ]


def have_seq() -> bool:
    return os.path.exists(SEQ_REF_LIB) and os.path.exists(SEQ_B200_LIB)


def run_sequence(which: str, poses, Z, nZ, model: dict, cfg: dict, *, pose_cov, Q_lmk=None, neff_threshold=0.0,
                 min_updates_before_resample=1, precision=64, seed48=1, births=None):
    """which = 'ref' (the reference's RBPHDFilter.hpp) or 'b200' (include/rfs_b200/RBPHDFilter.hpp).
    poses [K][N][3], Z [K][nZmax][2], nZ [K].  Returns (Result, n_resampled, gm_size_trace)."""
    lib = C.CDLL(SEQ_REF_LIB if which == "ref" else SEQ_B200_LIB)
    fn = lib.seq_run_ref if which == "ref" else lib.seq_run_b200
    fn.restype = C.c_int
    fn.argtypes = [C.POINTER(SeqIO)]
    poses = np.ascontiguousarray(poses, dtype=np.float64)
    Z = np.ascontiguousarray(Z, dtype=np.float64)
    nZ = np.ascontiguousarray(nZ, dtype=np.int32)
    K, N = poses.shape[0], poses.shape[1]
    R = np.ascontiguousarray(np.array(model["R"], dtype=np.float64).reshape(4))
    pc = np.ascontiguousarray(pose_cov, dtype=np.float64)
    Q = None if Q_lmk is None else np.ascontiguousarray(np.array(Q_lmk, dtype=np.float64).reshape(4))
    cap_total = N * 512
    out = dict(count_out=np.zeros(N, np.int32), mean_out=np.zeros((cap_total, 2)), cov_out=np.zeros((cap_total, 3)),
               w_out=np.zeros(cap_total), weight_out=np.zeros(N), n_resampled=np.zeros(1, np.int32),
               gm_size_trace=np.zeros(K, np.int32))
    io = SeqIO()
    io.N, io.n_steps, io.nZ_max = N, K, Z.shape[1]
    io.poses, io.pose_cov, io.Z, io.nZ, io.R = poses.ctypes.data, pc.ctypes.data, Z.ctypes.data, nZ.ctypes.data, R.ctypes.data
    io.Q_lmk = None if Q is None else Q.ctypes.data
    io.Pd, io.clutter = model["Pd"], model["clutter_intensity"]
    io.range_min, io.range_max, io.range_buffer = model["range_min"], model["range_max"], model["range_buffer"]
    io.thr_r, io.thr_b = model["innov_thr_range"], model["innov_thr_bearing"]
    io.birth_w, io.gate = cfg["birth_gaussian_weight"], cfg["new_gaussian_create_innov_md_threshold"]
    io.eval_w, io.wl_gate = cfg["eval_point_gaussian_weight"], cfg["meas_likelihood_md_threshold"]
    io.merge_t, io.merge_f, io.prune_t = cfg["merging_threshold"], cfg["merging_cov_inflation_factor"], cfg["pruning_threshold"]
    io.n_eval, io.use_sc = cfg["eval_point_count"], cfg["use_cluster_process"]
    io.min_updates_before_resample = min_updates_before_resample
    io.neff_threshold, io.precision, io.seed48 = neff_threshold, precision, seed48
    io.cap_total = cap_total
    if births:   # candidate-list births: dict(count_thr, check_thr, cur_count_thr, support_dist)
        io.birth_count_thr, io.birth_check_thr = births["count_thr"], births["check_thr"]
        io.birth_cur_thr, io.birth_support_dist = births["cur_count_thr"], births["support_dist"]
    for k, a in out.items():
        setattr(io, k, a.ctypes.data)
    rc = fn(C.byref(io))
    if rc != 0:
        raise RuntimeError(f"sequence harness ({which}) failed rc={rc}")
    tot = int(out["count_out"].sum())
    res = Result(out["count_out"], out["mean_out"][:tot].copy(), out["cov_out"][:tot].copy(), out["w_out"][:tot].copy(),
                 None, out["weight_out"], None, None, None, 0.0)
    return res, int(out["n_resampled"][0]), out["gm_size_trace"]

"""PHDUpdater — host-side mirror of rfs::RBPHDFilter<...>::update() over the C ABI.

Names follow the reference (include/RBPHDFilter.hpp:183-251): update(Z), getGMSize(i),
getLandmark(i, m), particle weights; configs are the reference's public config structs as
dicts (synth.DEFAULT_MODEL / DEFAULT_CFG).  All compute happens in librfsb200.so on the GPU;
this class only marshals numpy arrays.  No CPU fallback: construction raises without a GPU.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi


class RFSB200Error(RuntimeError):
    pass


def _check(lib, ctx, rc, what):
    if rc != 0:
        msg = lib.rfsb200_last_error(ctx)
        raise RFSB200Error(f"{what}: {capi.ERRORS.get(rc, rc)}: {msg.decode() if msg else ''}")


def pinned_array(shape, dtype=np.float64) -> np.ndarray:
    """numpy view of page-locked host memory from rfsb200_host_alloc (never freed: bench/test aid)."""
    lib = capi.load_library()
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = C.c_void_p()
    rc = lib.rfsb200_host_alloc(C.byref(p), max(n, 16))
    if rc != 0:
        raise RFSB200Error(f"host_alloc: {capi.ERRORS.get(rc, rc)}")
    buf = (C.c_char * max(n, 16)).from_address(p.value)
    return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)


class PHDUpdater:
    def __init__(self, n_particles: int, gm_capacity: int = 256, work_capacity: int = 0,
                 z_capacity: int = 64, device: int = 0, precision: int = 32, lmk_dim: int = 2):
        """lmk_dim = 2: MeasurementModel_RngBrg (x, y) / (range, bearing);
        lmk_dim = 3: MeasurementModel_VictoriaPark (x, y, diameter) / (range, bearing, diameter)."""
        self.lib = capi.load_library()
        self.N = int(n_particles)
        self.D = int(lmk_dim)
        self.NC = self.D * (self.D + 1) // 2
        self.gm_capacity = int(gm_capacity)
        d = capi.Dims()
        d.n_particles = self.N
        d.gm_capacity = self.gm_capacity
        d.work_capacity = int(work_capacity) if work_capacity else self.gm_capacity
        d.z_capacity = int(z_capacity)
        d.lmk_dim, d.meas_dim, d.pose_dim = self.D, self.D, 3
        d.device = int(device)
        d.precision = int(precision)
        self.ctx = C.c_void_p()
        rc = self.lib.rfsb200_create(C.byref(self.ctx), C.byref(d))
        if rc != 0:
            msg = self.lib.rfsb200_last_error(None)
            self.ctx = None
            raise RFSB200Error(f"rfsb200_create: {capi.ERRORS.get(rc, rc)}: {msg.decode() if msg else ''}")
        self._keep = []
        self._keep_pose = []

    # ---- lifetime -------------------------------------------------------------------------
    def close(self):
        if getattr(self, "ctx", None):
            self.lib.rfsb200_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- configuration ----------------------------------------------------------------------
    def set_model(self, md: dict):
        """Mirror of the live plugin configuration; call again whenever it changes (the Victoria Park
        lidar scan changes before every update, src/rbphdslam_VictoriaPark.cpp:582)."""
        d = capi.model_desc(md)
        _check(self.lib, self.ctx, self.lib.rfsb200_set_model(self.ctx, C.byref(d)), "set_model")

    def set_model_desc(self, d):
        """set_model with a prebuilt capi.ModelDesc (no dict conversion on the per-update path)."""
        _check(self.lib, self.ctx, self.lib.rfsb200_set_model(self.ctx, C.byref(d)), "set_model")

    def set_filter_cfg(self, fc: dict, brute_force_merge: bool = False):
        c = capi.filter_cfg(fc)
        c.reserved_i[0] = 1 if brute_force_merge else 0
        _check(self.lib, self.ctx, self.lib.rfsb200_set_filter_cfg(self.ctx, C.byref(c)), "set_filter_cfg")

    def set_stream(self, cuda_stream: int | None):
        """cuda_stream: a cudaStream_t handle (0 = the legacy default stream); None = the ctx-owned stream."""
        ext = 0 if cuda_stream is None else 1
        _check(self.lib, self.ctx, self.lib.rfsb200_set_stream(self.ctx, C.c_void_p(cuda_stream or 0), ext), "set_stream")

    def synchronize(self):
        _check(self.lib, self.ctx, self.lib.rfsb200_synchronize(self.ctx), "synchronize")

    # ---- state in --------------------------------------------------------------------------------
    def upload_maps(self, count, mean, cov, w):
        count = np.ascontiguousarray(count, dtype=np.int32)
        mean = np.ascontiguousarray(mean, dtype=np.float64)
        cov = np.ascontiguousarray(cov, dtype=np.float64)
        w = np.ascontiguousarray(w, dtype=np.float64)
        assert count.shape[0] == self.N
        self._keep = [count, mean, cov, w]  # async copies read these until the next sync
        _check(self.lib, self.ctx,
               self.lib.rfsb200_upload_maps(self.ctx, capi.ptr(count), capi.ptr(mean), capi.ptr(cov), capi.ptr(w)),
               "upload_maps")

    def append_gaussians(self, count, mean, cov, w):
        """GaussianMixture::addGaussian for host-decided births: count[i] packed Gaussians behind the map of particle i."""
        count = np.ascontiguousarray(count, dtype=np.int32)
        mean = np.ascontiguousarray(mean, dtype=np.float64)
        cov = np.ascontiguousarray(cov, dtype=np.float64)
        w = np.ascontiguousarray(w, dtype=np.float64)
        _check(self.lib, self.ctx,
               self.lib.rfsb200_append_gaussians(self.ctx, capi.ptr(count), capi.ptr(mean), capi.ptr(cov), capi.ptr(w)),
               "append_gaussians")

    def set_poses(self, pose, pose_cov=None, weight=None):
        """Particle poses (+ optional pose covariance, Q1) and weights; ascontiguousarray keeps a
        pinned float64 input as it is, so the H2D copy reads the caller's page-locked buffer."""
        pose = np.ascontiguousarray(pose, dtype=np.float64)
        mode = 0
        pc = None
        if pose_cov is not None:
            pc = np.ascontiguousarray(pose_cov, dtype=np.float64)
            mode = 1 if pc.ndim == 1 else 2
        wt = None if weight is None else np.ascontiguousarray(weight, dtype=np.float64)
        self._keep_pose = [pose, pc, wt]   # the async copies read these until the next sync; replaced, not accumulated
        _check(self.lib, self.ctx,
               self.lib.rfsb200_set_poses(self.ctx, capi.ptr(pose), capi.ptr(pc), mode, capi.ptr(wt)),
               "set_poses")

    def load_workload(self, wl):
        self.set_model(wl.model)
        self.set_filter_cfg(wl.cfg)
        self.upload_maps(wl.count, wl.mean, wl.cov, wl.w)
        self.set_poses(wl.pose, wl.pose_cov, wl.weight)

    # ---- the hot path ---------------------------------------------------------------------------
    def update(self, Z, flags: int = capi.UPDATE_DEFAULT, want_stats: bool = True):
        """RBPHDFilter::update(Z) for all particles.  Returns a StepOut (or None if async)."""
        Z = np.ascontiguousarray(Z, dtype=np.float64).reshape(-1, self.D)
        out = capi.StepOut() if want_stats else None
        rc = self.lib.rfsb200_update(self.ctx, capi.ptr(Z), Z.shape[0], flags,
                                     C.byref(out) if out is not None else None)
        _check(self.lib, self.ctx, rc, "update")
        return out

    def update_host(self, pose, pose_cov, weight, Z, flags: int = capi.UPDATE_DEFAULT, w_out=None, unused_out=None,
                    nfov_out=None, want_stats: bool = False):
        """set_poses + update + get_weights + get_unused in one ABI call with one synchronisation
        (rfsb200_update_host).  All arrays float64 / uint64 / int32, C-contiguous, ideally pinned_array()s."""
        mode = 0 if pose_cov is None else (1 if pose_cov.ndim == 1 else 2)
        out = capi.StepOut() if want_stats else None
        p = self._host_ptr   # building a ctypes pointer costs ~2.5 us per array: the per-step buffers are bound once
        rc = self.lib.rfsb200_update_host(self.ctx, p(pose), p(pose_cov), mode, p(weight), p(Z),
                                          Z.shape[0], flags, p(w_out), p(unused_out), p(nfov_out),
                                          C.byref(out) if out is not None else None)
        if rc != 0:
            _check(self.lib, self.ctx, rc, "update_host")
        return out

    def _host_ptr(self, a):
        """ctypes pointer of a host array, cached per array object (the cache holds a reference, so an id cannot be
        recycled while its pointer is cached; numpy never moves the data of a live array)."""
        if a is None:
            return None
        cache = self.__dict__.setdefault("_ptr_cache", {})
        hit = cache.get(id(a))
        if hit is not None and hit[0] is a:
            return hit[1]
        if len(cache) >= 64:
            cache.clear()
        ptr = capi.ptr(a)
        cache[id(a)] = (a, ptr)
        return ptr

    # ---- the callers either side (predict's map part, resampling's data movement) -----------------
    def predict_maps(self, Q_lmk=None, add_births: bool = True, birth_weight: float = 0.0):
        """RBPHDFilter::predict() minus the particle propagation: births, then P += Q."""
        q = None if Q_lmk is None else np.ascontiguousarray(Q_lmk, dtype=np.float64).reshape(self.NC)
        _check(self.lib, self.ctx,
               self.lib.rfsb200_predict_maps(self.ctx, capi.ptr(q), 1 if add_births else 0, float(birth_weight)),
               "predict_maps")

    def birth_candidates(self, birth_weight: float, support_dist: float, count_thr: int, check_thr: int, cur_count_thr: int,
                         parent=None):
        """addBirthGaussians() in its candidate-list form on the device (include/RBPHDFilter.hpp:1000-1080): the unused
        measurements of the last update feed the per-particle candidate lists kept by the ctx; candidates that become
        real are appended to the maps.  parent = the parent slots after a resampling (resampleOccured_), else None."""
        b = capi.BirthCfg()
        b.birth_weight, b.support_dist = float(birth_weight), float(support_dist)
        b.count_threshold, b.check_threshold, b.current_count_threshold = int(count_thr), int(check_thr), int(cur_count_thr)
        par = None if parent is None else np.ascontiguousarray(parent, dtype=np.int32)
        _check(self.lib, self.ctx, self.lib.rfsb200_birth_candidates(self.ctx, C.byref(b), capi.ptr(par)), "birth_candidates")

    def get_birth_candidates(self):
        """-> (n [N], mean [N][64][D], cov [N][64][NC], support [N][64], checks [N][64])"""
        cap = capi.BIRTH_CAND_CAP
        n = np.zeros(self.N, np.int32)
        mean = np.zeros((self.N, cap, self.D))
        cov = np.zeros((self.N, cap, self.NC))
        sup = np.zeros((self.N, cap), np.int32)
        chk = np.zeros((self.N, cap), np.int32)
        _check(self.lib, self.ctx, self.lib.rfsb200_get_birth_candidates(self.ctx, capi.ptr(n), capi.ptr(mean), capi.ptr(cov),
                                                                         capi.ptr(sup), capi.ptr(chk)), "get_birth_candidates")
        return n, mean, cov, sup, chk

    def set_birth_candidates(self, n, mean, cov, support, checks):
        cap = capi.BIRTH_CAND_CAP
        n = np.ascontiguousarray(n, dtype=np.int32)
        mean = np.ascontiguousarray(mean, dtype=np.float64).reshape(self.N, cap, self.D)
        cov = np.ascontiguousarray(cov, dtype=np.float64).reshape(self.N, cap, self.NC)
        sup = np.ascontiguousarray(support, dtype=np.int32).reshape(self.N, cap)
        chk = np.ascontiguousarray(checks, dtype=np.int32).reshape(self.N, cap)
        _check(self.lib, self.ctx, self.lib.rfsb200_set_birth_candidates(self.ctx, capi.ptr(n), capi.ptr(mean), capi.ptr(cov),
                                                                         capi.ptr(sup), capi.ptr(chk)), "set_birth_candidates")

    def propagate(self, model: str, u, *, Q=None, input_cov=None, dt: float = 0.0, use_model_noise: bool = True,
                  use_input_noise: bool = False, ackerman=(0.0, 1.0, 0.0, 0.0), seed: int = 0, step: int = 0):
        """ParticleFilter::propagate() on the device: ProcessModel::sample() of MotionModel_Odometry2d ("odometry2d",
        u = (dx, dy, dtheta)) or MotionModel_Ackerman2d ("ackerman2d", u = (velocity, steering), ackerman = (h, l,
        poi_dx, poi_dy)) for every particle, in place on the device poses."""
        m = capi.MotionDesc()
        m.model_id = capi.MOTION_ODOMETRY2D if model == "odometry2d" else capi.MOTION_ACKERMAN2D
        m.use_model_noise, m.use_input_noise = int(use_model_noise), int(use_input_noise)
        if Q is not None:
            for i, v in enumerate(np.asarray(Q, dtype=np.float64).reshape(9)):
                m.Q[i] = float(v)
        for i, v in enumerate(np.asarray(u, dtype=np.float64).ravel()):
            m.input[i] = float(v)
        if input_cov is not None:
            for i, v in enumerate(np.asarray(input_cov, dtype=np.float64).ravel()):
                m.input_cov[i] = float(v)
        m.dt = float(dt)
        m.ackerman_h, m.ackerman_l, m.ackerman_dx, m.ackerman_dy = (float(v) for v in ackerman)
        m.seed, m.step_counter = int(seed), int(step)
        _check(self.lib, self.ctx, self.lib.rfsb200_propagate(self.ctx, C.byref(m)), "propagate")

    def get_poses(self) -> np.ndarray:
        p = np.zeros((self.N, 3))
        _check(self.lib, self.ctx, self.lib.rfsb200_get_poses(self.ctx, capi.ptr(p)), "get_poses")
        return p

    def resample(self, map_src, aux_src=None, weight: float | None = 1.0):
        ms = np.ascontiguousarray(map_src, dtype=np.int32)
        au = None if aux_src is None else np.ascontiguousarray(aux_src, dtype=np.int32)
        wv = None if weight is None else np.array([weight], dtype=np.float64)
        _check(self.lib, self.ctx, self.lib.rfsb200_resample(self.ctx, capi.ptr(ms), capi.ptr(au), capi.ptr(wv)), "resample")

    # ---- cross-GPU particle exchange (records in device memory) ---------------------------------------
    def particle_record_bytes(self) -> int:
        b = C.c_int64()
        _check(self.lib, self.ctx, self.lib.rfsb200_particle_record_bytes(self.ctx, C.byref(b)), "particle_record_bytes")
        return int(b.value)

    def export_particles(self, idx, dev_ptr: int):
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        _check(self.lib, self.ctx, self.lib.rfsb200_export_particles(self.ctx, capi.ptr(idx), len(idx), C.c_void_p(dev_ptr)),
               "export_particles")

    def import_particles(self, slot, dev_ptr: int, weight: float = 1.0):
        slot = np.ascontiguousarray(slot, dtype=np.int32)
        _check(self.lib, self.ctx,
               self.lib.rfsb200_import_particles(self.ctx, capi.ptr(slot), len(slot), C.c_void_p(dev_ptr), float(weight)),
               "import_particles")

    def comm_export(self) -> bytes:
        h = (C.c_ubyte * 64)()
        _check(self.lib, self.ctx, self.lib.rfsb200_comm_export(self.ctx, C.cast(h, C.c_void_p)), "comm_export")
        return bytes(h)

    def comm_connect(self, rank: int, world: int, handles: list[bytes]):
        buf = (C.c_ubyte * (64 * world)).from_buffer_copy(b"".join(handles))
        _check(self.lib, self.ctx, self.lib.rfsb200_comm_connect(self.ctx, rank, world, C.cast(buf, C.c_void_p)), "comm_connect")

    def comm_barrier(self):
        """Barrier of the connected ranks on the ctx stream (rfsb200_comm_barrier)."""
        _check(self.lib, self.ctx, self.lib.rfsb200_comm_barrier(self.ctx), "comm_barrier")

    def comm_connect_local(self, rank: int, world: int, updaters):
        """Connects contexts of this process (rfsb200_comm_connect_local); updaters = the PHDUpdater of every rank, in order."""
        arr = (C.c_void_p * world)(*[u.ctx for u in updaters])
        _check(self.lib, self.ctx, self.lib.rfsb200_comm_connect_local(self.ctx, rank, world, C.cast(arr, C.c_void_p)), "comm_connect_local")

    def comm_resolve(self):
        """Runs the normalisation a deferred step (UPDATE_DEFER_NORMALIZE) left open, if any (rfsb200_comm_resolve)."""
        _check(self.lib, self.ctx, self.lib.rfsb200_comm_resolve(self.ctx), "comm_resolve")

    def comm_error(self) -> bool:
        f = C.c_int32()
        _check(self.lib, self.ctx, self.lib.rfsb200_comm_error(self.ctx, C.byref(f)), "comm_error")
        return bool(f.value)

    def weight_sums_device_ptr(self) -> int:
        p = C.c_void_p()
        _check(self.lib, self.ctx, self.lib.rfsb200_weight_sums_device(self.ctx, C.byref(p)), "weight_sums_device")
        return int(p.value)

    def normalize(self):
        _check(self.lib, self.ctx, self.lib.rfsb200_normalize(self.ctx), "normalize")

    # ---- state out -------------------------------------------------------------------------------
    def get_weights(self, which: int = 0, out: np.ndarray | None = None) -> np.ndarray:
        w = np.zeros(self.N) if out is None else out
        _check(self.lib, self.ctx, self.lib.rfsb200_get_weights(self.ctx, which, capi.ptr(w)), "get_weights")
        return w

    def get_gm_sizes(self, which: int = 0) -> np.ndarray:
        n = np.zeros(self.N, dtype=np.int32)
        _check(self.lib, self.ctx, self.lib.rfsb200_get_gm_sizes(self.ctx, which, capi.ptr(n)), "get_gm_sizes")
        return n

    def getGMSize(self, i: int, which: int = 0) -> int:
        if i < 0 or i >= self.N:
            return -1  # include/RBPHDFilter.hpp:1153-1159
        return int(self.get_gm_sizes(which)[i])

    def get_map(self, i: int, which: int = 0):
        cap = self.gm_capacity + 8
        n = C.c_int32()
        mean = np.zeros((cap, self.D)); cov = np.zeros((cap, self.NC)); w = np.zeros(cap)
        _check(self.lib, self.ctx,
               self.lib.rfsb200_get_map(self.ctx, which, i, cap, C.byref(n), capi.ptr(mean), capi.ptr(cov), capi.ptr(w)),
               "get_map")
        k = n.value
        return mean[:k].copy(), cov[:k].copy(), w[:k].copy()

    def getLandmark(self, i: int, m: int, which: int = 0):
        """(ok, mean[2], cov[2,2], w) — include/RBPHDFilter.hpp:1161-1178."""
        if i < 0 or i >= self.N:
            return False, None, None, None
        mean, cov, w = self.get_map(i, which)
        if m < 0 or m >= len(w):
            return False, None, None, None
        S = np.zeros((self.D, self.D))
        S[np.triu_indices(self.D)] = cov[m]
        S = S + S.T - np.diag(np.diag(S))
        return True, mean[m].copy(), S, float(w[m])

    def download_maps(self, which: int = 0):
        cap_total = self.N * (self.gm_capacity + 8)
        count = np.zeros(self.N, dtype=np.int32)
        mean = np.zeros((cap_total, self.D)); cov = np.zeros((cap_total, self.NC)); w = np.zeros(cap_total)
        _check(self.lib, self.ctx,
               self.lib.rfsb200_download_maps(self.ctx, which, cap_total, capi.ptr(count), capi.ptr(mean),
                                              capi.ptr(cov), capi.ptr(w)), "download_maps")
        t = int(count.sum())
        return count, mean[:t].copy(), cov[:t].copy(), w[:t].copy()

    def get_unused(self, out_mask: np.ndarray | None = None, out_nfov: np.ndarray | None = None):
        mask = np.zeros(self.N, dtype=np.uint64) if out_mask is None else out_mask
        nfov = np.zeros(self.N, dtype=np.int32) if out_nfov is None else out_nfov
        _check(self.lib, self.ctx, self.lib.rfsb200_get_unused(self.ctx, capi.ptr(mask), capi.ptr(nfov)), "get_unused")
        return mask, nfov

    def get_flags(self) -> np.ndarray:
        f = np.zeros(self.N, dtype=np.int32)
        _check(self.lib, self.ctx, self.lib.rfsb200_get_flags(self.ctx, capi.ptr(f)), "get_flags")
        return f

    def profile_begin(self, max_updates: int):
        _check(self.lib, self.ctx, self.lib.rfsb200_profile_begin(self.ctx, int(max_updates)), "profile_begin")

    def profile_read(self) -> np.ndarray:
        us = np.zeros(4096, dtype=np.float32)
        n = C.c_int32()
        _check(self.lib, self.ctx, self.lib.rfsb200_profile_read(self.ctx, capi.ptr(us), 4096, C.byref(n)), "profile_read")
        return us[:n.value].copy()

    def stage_times(self) -> dict:
        """Per-phase times of the last update that ran with capi.UPDATE_STAGE_TIMES (rfsb200_get_stage_times)."""
        st = capi.StageTimes()
        _check(self.lib, self.ctx, self.lib.rfsb200_get_stage_times(self.ctx, C.byref(st)), "stage_times")
        d = {k: getattr(st, k) for k, _ in capi.StageTimes._fields_ if not k.startswith("reserved")}
        d["merge_parts"] = dict(zip(("cell_sort", "pair_search", "clusters", "cluster_loops"), list(st.reserved)))
        return d

    def permanent(self, A) -> np.ndarray:
        A = np.ascontiguousarray(A, dtype=np.float64)
        if A.ndim == 2:
            A = A[None]
        out = np.zeros(A.shape[0])
        _check(self.lib, self.ctx,
               self.lib.rfsb200_permanent(self.ctx, capi.ptr(A), A.shape[1], A.shape[0], capi.ptr(out)), "permanent")
        return out

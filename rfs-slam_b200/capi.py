"""ctypes binding of include/rfsb200.h (the C-ABI shared library librfsb200.so).

This is plumbing only: the product is the CUDA library.  There is NO CPU fallback: if the
library is missing or no CUDA device is present, calls fail loudly.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "librfsb200.so")
if os.environ.get("RFSB200_LIB"):   # development aid: A/B builds of the same library
    LIB_PATH = os.environ["RFSB200_LIB"]

OK = 0
UPDATE_DEFAULT = 0
UPDATE_NO_COMMIT = 1
UPDATE_NO_NORMALIZE = 2
UPDATE_FUSED_ALLREDUCE = 4
UPDATE_STAGE_TIMES = 8
UPDATE_DEFER_NORMALIZE = 16

ERRORS = {0: "OK", -1: "EINVAL", -2: "ECUDA", -3: "ENOMEM", -4: "ECAPACITY", -5: "EUNSUPPORTED",
          -6: "ESTATE", -7: "ENODEVICE"}


class Dims(C.Structure):
    _fields_ = [("n_particles", C.c_int32), ("gm_capacity", C.c_int32), ("work_capacity", C.c_int32),
                ("z_capacity", C.c_int32), ("lmk_dim", C.c_int32), ("meas_dim", C.c_int32),
                ("pose_dim", C.c_int32), ("device", C.c_int32), ("precision", C.c_int32),
                ("reserved", C.c_int32 * 7)]


class ModelDesc(C.Structure):
    _fields_ = [("model_id", C.c_int32), ("reserved0", C.c_int32), ("R", C.c_double * 9),
                ("Pd", C.c_double), ("clutter_intensity", C.c_double), ("clutter_integral", C.c_double),
                ("range_min", C.c_double), ("range_max", C.c_double), ("range_buffer", C.c_double),
                ("innov_thr_range", C.c_double), ("innov_thr_bearing", C.c_double),
                # RFSB200_MODEL_VICTORIAPARK only
                ("bearing_min", C.c_double), ("bearing_max", C.c_double), ("Slb", C.c_double),
                ("buffer_zone_pd", C.c_double), ("pd_table", C.c_double * 16),
                ("pd_table_n", C.c_int32), ("scan_n", C.c_int32), ("scan", C.c_void_p),
                ("reserved", C.c_double * 4)]


class FilterCfg(C.Structure):
    _fields_ = [("birth_gaussian_weight", C.c_double),
                ("new_gaussian_create_innov_md_threshold", C.c_double),
                ("eval_point_gaussian_weight", C.c_double),
                ("meas_likelihood_md_threshold", C.c_double),
                ("merging_threshold", C.c_double),
                ("merging_cov_inflation_factor", C.c_double),
                ("pruning_threshold", C.c_double),
                ("eval_point_count", C.c_int32), ("use_cluster_process", C.c_int32),
                ("assignment_sum_method", C.c_int32), ("murty_compat", C.c_int32), ("reserved_i", C.c_int32 * 2),
                ("reserved", C.c_double * 6)]


MOTION_ODOMETRY2D, MOTION_ACKERMAN2D = 1, 2


class MotionDesc(C.Structure):
    _fields_ = [("model_id", C.c_int32), ("use_model_noise", C.c_int32), ("use_input_noise", C.c_int32),
                ("reserved_i", C.c_int32), ("Q", C.c_double * 9), ("input", C.c_double * 3),
                ("input_cov", C.c_double * 9), ("dt", C.c_double), ("ackerman_h", C.c_double),
                ("ackerman_l", C.c_double), ("ackerman_dx", C.c_double), ("ackerman_dy", C.c_double),
                ("seed", C.c_uint64), ("step_counter", C.c_uint64), ("reserved", C.c_double * 4)]


class StepOut(C.Structure):
    _fields_ = [("sum_w", C.c_double), ("sum_w2", C.c_double), ("n_eff", C.c_double),
                ("gm_total_in", C.c_int64), ("gm_total_out", C.c_int64), ("gm_max_out", C.c_int32),
                ("n_overflow", C.c_int32), ("n_murty", C.c_int32), ("n_launches", C.c_int32),
                ("elapsed_us", C.c_float), ("n_merge_redo", C.c_int32), ("reserved", C.c_int32 * 6)]


MODEL_RNGBRG, MODEL_VICTORIAPARK = 1, 2


def model_desc(md: dict) -> ModelDesc:
    """dict -> rfsb200_model_desc.  For the Victoria Park model the returned struct keeps the scan
    array alive (d._scan) because the descriptor only carries a pointer to it."""
    d = ModelDesc()
    d.model_id = int(md.get("model_id", 1))
    for i, v in enumerate(list(md["R"])):   # row-major meas_dim x meas_dim
        d.R[i] = float(v)
    for k in ("Pd", "clutter_intensity", "clutter_integral", "range_min", "range_max", "range_buffer",
              "innov_thr_range", "innov_thr_bearing"):
        setattr(d, k, float(md.get(k, 0.0)))
    if d.model_id == MODEL_VICTORIAPARK:
        for k in ("bearing_min", "bearing_max", "Slb", "buffer_zone_pd"):
            setattr(d, k, float(md[k]))
        tab = list(md["pd_table"])
        d.pd_table_n = len(tab)
        for i, v in enumerate(tab):
            d.pd_table[i] = float(v)
        scan = np.ascontiguousarray(md["scan"], dtype=np.float64)
        d._scan = scan
        d.scan_n = int(scan.shape[0])
        d.scan = scan.ctypes.data
    return d


def filter_cfg(fc: dict) -> FilterCfg:
    c = FilterCfg()
    for k in ("birth_gaussian_weight", "new_gaussian_create_innov_md_threshold",
              "eval_point_gaussian_weight", "meas_likelihood_md_threshold", "merging_threshold",
              "merging_cov_inflation_factor", "pruning_threshold"):
        setattr(c, k, float(fc[k]))
    for k in ("eval_point_count", "use_cluster_process", "assignment_sum_method", "murty_compat"):
        setattr(c, k, int(fc.get(k, 0)))
    return c


BIRTH_CAND_CAP = 64


class BirthCfg(C.Structure):
    """rfsb200_birth_cfg"""
    _fields_ = [("birth_weight", C.c_double), ("support_dist", C.c_double), ("count_threshold", C.c_uint32),
                ("check_threshold", C.c_uint32), ("current_count_threshold", C.c_uint32), ("reserved", C.c_uint32)]


class StageTimes(C.Structure):
    """rfsb200_stage_times"""
    _fields_ = [("kernel_us", C.c_double), ("setup_us", C.c_double), ("particles_us", C.c_double), ("epilogue_us", C.c_double),
                ("share_load", C.c_double), ("share_map_update_kf", C.c_double), ("share_weighting", C.c_double),
                ("share_merge", C.c_double), ("share_prune", C.c_double), ("warp_cycles", C.c_double),
                ("warps_per_cta", C.c_int32), ("reserved_i", C.c_int32), ("reserved", C.c_double * 4)]


# name -> (restype, argtypes); the not-gpu test checks every symbol of include/rfsb200.h is here
_P = C.c_void_p
_SIGS = {
    "rfsb200_abi_version": (C.c_int, []),
    "rfsb200_device_count": (C.c_int, []),
    "rfsb200_create": (C.c_int, [C.POINTER(_P), C.POINTER(Dims)]),
    "rfsb200_destroy": (C.c_int, [_P]),
    "rfsb200_last_error": (C.c_char_p, [_P]),
    "rfsb200_set_stream": (C.c_int, [_P, _P, C.c_int]),
    "rfsb200_synchronize": (C.c_int, [_P]),
    "rfsb200_set_model": (C.c_int, [_P, C.POINTER(ModelDesc)]),
    "rfsb200_set_filter_cfg": (C.c_int, [_P, C.POINTER(FilterCfg)]),
    "rfsb200_upload_maps": (C.c_int, [_P, _P, _P, _P, _P]),
    "rfsb200_set_poses": (C.c_int, [_P, _P, _P, C.c_int, _P]),
    "rfsb200_append_gaussians": (C.c_int, [_P, _P, _P, _P, _P]),
    "rfsb200_update": (C.c_int, [_P, _P, C.c_int32, C.c_uint32, C.POINTER(StepOut)]),
    "rfsb200_update_host": (C.c_int, [_P, _P, _P, C.c_int, _P, _P, C.c_int32, C.c_uint32, _P, _P, _P, C.POINTER(StepOut)]),
    "rfsb200_predict_maps": (C.c_int, [_P, _P, C.c_int32, C.c_double]),
    "rfsb200_birth_candidates": (C.c_int, [_P, C.POINTER(BirthCfg), _P]),
    "rfsb200_get_birth_candidates": (C.c_int, [_P, _P, _P, _P, _P, _P]),
    "rfsb200_set_birth_candidates": (C.c_int, [_P, _P, _P, _P, _P, _P]),
    "rfsb200_propagate": (C.c_int, [_P, C.POINTER(MotionDesc)]),
    "rfsb200_get_poses": (C.c_int, [_P, _P]),
    "rfsb200_resample": (C.c_int, [_P, _P, _P, _P]),
    "rfsb200_particle_record_bytes": (C.c_int, [_P, C.POINTER(C.c_int64)]),
    "rfsb200_export_particles": (C.c_int, [_P, _P, C.c_int32, _P]),
    "rfsb200_import_particles": (C.c_int, [_P, _P, C.c_int32, _P, C.c_double]),
    "rfsb200_comm_export": (C.c_int, [_P, _P]),
    "rfsb200_comm_connect": (C.c_int, [_P, C.c_int32, C.c_int32, _P]),
    "rfsb200_comm_error": (C.c_int, [_P, C.POINTER(C.c_int32)]),
    "rfsb200_comm_barrier": (C.c_int, [_P]),
    "rfsb200_comm_resolve": (C.c_int, [_P]),
    "rfsb200_comm_connect_local": (C.c_int, [_P, C.c_int32, C.c_int32, _P]),
    "rfsb200_weight_sums_device": (C.c_int, [_P, C.POINTER(_P)]),
    "rfsb200_normalize": (C.c_int, [_P]),
    "rfsb200_get_weights": (C.c_int, [_P, C.c_int, _P]),
    "rfsb200_get_gm_sizes": (C.c_int, [_P, C.c_int, _P]),
    "rfsb200_get_map": (C.c_int, [_P, C.c_int, C.c_int32, C.c_int32, C.POINTER(C.c_int32), _P, _P, _P]),
    "rfsb200_download_maps": (C.c_int, [_P, C.c_int, C.c_int64, _P, _P, _P, _P]),
    "rfsb200_get_unused": (C.c_int, [_P, _P, _P]),
    "rfsb200_get_flags": (C.c_int, [_P, _P]),
    "rfsb200_permanent": (C.c_int, [_P, _P, C.c_int32, C.c_int32, _P]),
    "rfsb200_profile_begin": (C.c_int, [_P, C.c_int32]),
    "rfsb200_profile_read": (C.c_int, [_P, _P, C.c_int32, C.POINTER(C.c_int32)]),
    "rfsb200_get_stage_times": (C.c_int, [_P, C.POINTER(StageTimes)]),
    "rfsb200_host_alloc": (C.c_int, [C.POINTER(_P), C.c_uint64]),
    "rfsb200_host_free": (C.c_int, [_P]),
}

_lib = None


def load_library(path: str | None = None):
    """dlopen librfsb200.so and bind every entry point.  Raises if it is not built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise RuntimeError(f"{p} is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback for the PHD update path)")
    lib = C.CDLL(p, mode=C.RTLD_GLOBAL)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


def ptr(a: np.ndarray | None):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)

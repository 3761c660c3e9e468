"""The C-ABI library loads on a CPU-only box and exports every symbol include/rfsb200.h declares;
the ctypes mirror agrees with the C structs.  No compute calls here (no GPU needed)."""
import ctypes as C
import os
import re
import subprocess

import pytest

import rfs_slam_b200  # noqa: F401
from rfs_slam_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "rfsb200.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rfsb200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = capi.load_library()
    names = _declared_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/rfsb200.h but not exported"


def test_ctypes_mirror_covers_header():
    assert sorted(capi._SIGS.keys()) == _declared_functions()


def test_abi_version_and_no_device_behaviour():
    lib = capi.load_library()
    assert lib.rfsb200_abi_version() == 2
    n = lib.rfsb200_device_count()
    assert n >= 0
    if n == 0:
        # product path must fail loudly without a GPU: no CPU fallback
        d = capi.Dims()
        d.n_particles, d.gm_capacity, d.work_capacity, d.z_capacity = 4, 32, 32, 8
        d.lmk_dim, d.meas_dim, d.pose_dim, d.device, d.precision = 2, 2, 3, 0, 32
        ctx = C.c_void_p()
        rc = lib.rfsb200_create(C.byref(ctx), C.byref(d))
        assert rc == -7  # RFSB200_ENODEVICE
        assert b"no CPU fallback" in lib.rfsb200_last_error(None)


def test_bad_arguments_are_rejected_without_a_device():
    lib = capi.load_library()
    assert lib.rfsb200_create(None, None) == -1
    assert lib.rfsb200_set_model(None, None) == -1
    assert lib.rfsb200_update(None, None, 0, 0, None) == -1


def test_struct_layout_matches_c(tmp_path):
    prog = tmp_path / "sz.c"
    prog.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "rfsb200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                    'sizeof(rfsb200_dims),sizeof(rfsb200_model_desc),sizeof(rfsb200_filter_cfg),sizeof(rfsb200_step_out),'
                    'offsetof(rfsb200_model_desc,Pd),offsetof(rfsb200_filter_cfg,eval_point_count),offsetof(rfsb200_step_out,elapsed_us),sizeof(rfsb200_stage_times),offsetof(rfsb200_stage_times,warps_per_cta));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    got = [C.sizeof(capi.Dims), C.sizeof(capi.ModelDesc), C.sizeof(capi.FilterCfg), C.sizeof(capi.StepOut),
           capi.ModelDesc.Pd.offset, capi.FilterCfg.eval_point_count.offset, capi.StepOut.elapsed_us.offset,
           C.sizeof(capi.StageTimes), capi.StageTimes.warps_per_cta.offset]
    assert [int(x) for x in out] == got


def test_header_is_plain_c():
    # the boundary must compile as C (no C++/torch types in the signatures)
    r = subprocess.run(["gcc", "-std=c99", "-fsyntax-only", "-x", "c", HEADER], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr

"""TEST INFRASTRUCTURE ONLY — builds tests/simt/_build/librfsb200_simt.so: the product's own sources
(rfs-slam_b200/csrc/rfsb200_abi.cu + the kernel headers it includes) compiled with g++ against the host
interpreter of tests/simt/simt.h, so that the CPU test suite can execute the kernel logic lane by lane.

Two purely syntactic rewrites are applied to a scratch copy of the sources (line numbers are kept, `#line`
points back at the originals):
  kernel<<<grid, block, smem, stream>>>(args)   ->  simt::Launcher(grid, block, smem, stream, "kernel")(kernel, args)
  extern __shared__ ... name[];                 ->  unsigned char* name = simt::dyn_smem();
Everything else — kernels, shared-memory carve-up, launch configuration, the ABI entry points — is the product code.
The package never loads this library (rfs-slam_b200/capi.py binds csrc/librfsb200.so only).
"""
from __future__ import annotations

import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "rfs-slam_b200", "csrc")
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "librfsb200_simt.so")
SOURCES = ["rfsb200_abi.cu", "phd_kernels.cuh", "phd_vp_kernels.cuh", "birth_kernels.cuh", "common.cuh", "murty_compat.hpp"]

_LAUNCH = re.compile(r"([A-Za-z_]\w*(?:<[^<>;()]*>)?)<<<(.*?)>>>\(")
_DYN_SMEM = re.compile(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?unsigned char (\w+)\[\];")


def rewrite(text: str) -> str:
    text = _LAUNCH.sub(lambda m: 'simt::Launcher(%s, "%s")(%s, ' % (m.group(2), m.group(1), m.group(1)), text)
    text = _DYN_SMEM.sub(lambda m: "unsigned char* %s = simt::dyn_smem();" % m.group(1), text)
    return text


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [os.path.join(ROOT, "include", "rfsb200.h")]
    deps += [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith(".h")] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Returns the path of the interpreter build of the ABI library (rebuilt when a source is newer)."""
    if not force and not _stale():
        return LIB
    src = os.path.join(OUT, "src")
    os.makedirs(src, exist_ok=True)
    for name in SOURCES:
        path = os.path.join(CSRC, name)
        text = open(path).read()
        if "<<<" in rewrite(text) or "extern __shared__" in rewrite(text):
            raise RuntimeError(f"{name}: a launch or a dynamic shared-memory declaration was not rewritten")
        text = rewrite(text).replace('#include "../../include/rfsb200.h"',
                                     '#include "%s"' % os.path.join(ROOT, "include", "rfsb200.h"))
        out = os.path.join(src, name.replace(".cu", ".cpp") if name.endswith(".cu") else name)
        with open(out, "w") as f:
            f.write('#line 1 "%s"\n' % path)
            f.write(text)
    cmd = [os.environ.get("CXX", "g++"), "-std=c++17", "-O1", "-g", "-fPIC", "-shared", "-fno-strict-aliasing",
           # the product library may already be loaded RTLD_GLOBAL in the same process (same symbol names):
           # bind this library's references to its own definitions
           "-Wl,-Bsymbolic", "-mtls-dialect=gnu2",
           "-DRFSB200_SIMT_HOST", "-Wno-unknown-pragmas", "-Wno-attributes", "-Wno-subobject-linkage",
           "-I", HERE, "-I", src, "-include", os.path.join(HERE, "simt.h"),
           os.path.join(src, "rfsb200_abi.cpp"), "-o", LIB + ".tmp"]
    if verbose:
        print("+", " ".join(cmd), flush=True)
    subprocess.run(cmd, check=True, cwd=src)
    os.replace(LIB + ".tmp", LIB)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))

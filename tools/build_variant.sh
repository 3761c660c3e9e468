#!/bin/bash
# build an A/B variant of the library from the working tree: bash tools/build_variant.sh <name> '<sed script for phd_kernels.cuh>' [extra nvcc flags]
#   ->  rfs-slam_b200/csrc/ab/<name>.so
set -e
NAME=$1; SED=$2; shift; shift
D=/tmp/variant_$NAME; rm -rf $D; mkdir -p $D
cp rfs-slam_b200/csrc/*.cuh rfs-slam_b200/csrc/*.hpp $D/
sed 's#"../../include/rfsb200.h"#"/root/repo/include/rfsb200.h"#' rfs-slam_b200/csrc/rfsb200_abi.cu > $D/rfsb200_abi.cu
if [ -n "$SED" ]; then sed -i "$SED" $D/phd_kernels.cuh; fi
mkdir -p rfs-slam_b200/csrc/ab
(cd $D && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --use_fast_math -Xcompiler -fPIC -shared -cudart static -Xptxas -v "$@" -o /root/repo/rfs-slam_b200/csrc/ab/$NAME.so rfsb200_abi.cu 2>&1 | grep -A2 "phd_update_kernelIfLb0ELi256ELb0" | grep "spill\|Used")

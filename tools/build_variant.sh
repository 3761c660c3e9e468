#!/bin/bash
# build an A/B variant of the library from the working tree with a sed script applied to phd_kernels.cuh:
#   bash tools/build_variant.sh <name> '<sed script>'   ->  rfs-slam_b200/csrc/ab/<name>.so
set -e
D=/tmp/variant_$1; rm -rf $D; mkdir -p $D
cp rfs-slam_b200/csrc/*.cuh $D/
sed 's#"../../include/rfsb200.h"#"/root/repo/include/rfsb200.h"#' rfs-slam_b200/csrc/rfsb200_abi.cu > $D/rfsb200_abi.cu
if [ -n "$2" ]; then sed -i "$2" $D/phd_kernels.cuh; fi
mkdir -p rfs-slam_b200/csrc/ab
(cd $D && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --use_fast_math -Xcompiler -fPIC -shared -cudart static -Xptxas -v -o /root/repo/rfs-slam_b200/csrc/ab/$1.so rfsb200_abi.cu 2>&1 | grep -A2 "phd_update_kernelIfLb0ELi256ELb0" | grep "spill\|Used")

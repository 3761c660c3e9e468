#!/bin/bash
# full ncu capture of the update kernel: bash tools/r2_ncu.sh <tag> [profile_step.py args]
TAG=$1; shift
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:phd_update -s 2 -c 1 -f -o gpurun_out/prof_$TAG \
  python tools/profile_step.py --steps 4 "$@" > gpurun_out/ncu_full_$TAG.log 2>&1
tail -3 gpurun_out/ncu_full_$TAG.log

#!/bin/bash
# A/B kernel timing of library builds: bash tools/r2_ab.sh <workload args for profile_step.py> -- lib1.so lib2.so ...
ARGS=()
while [ "$1" != "--" ] && [ $# -gt 0 ]; do ARGS+=("$1"); shift; done
shift
for L in "$@"; do
  for rep in 1 2; do
    echo -n "$L: "
    RFSB200_LIB=$PWD/$L timeout 300 python tools/profile_step.py --steps 12 "${ARGS[@]}" | awk '{print $3}' | sort -n | head -4 | tr '\n' ' '
    echo
  done
done

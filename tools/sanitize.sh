#!/bin/bash
# compute-sanitizer passes over small runs of every update kernel (memcheck, then racecheck on shared memory)
mkdir -p gpurun_out
for T in memcheck racecheck; do
  for ARGS in "--N 192 --nM 120 --nZ 20 --sc 1" "--N 192 --nM 120 --nZ 20 --sc 0" "--vp --N 128 --nM 100 --nZ 12 --sc 1" "--vp --N 128 --nM 100 --nZ 12 --sc 0"; do
    echo "== $T $ARGS"
    timeout 600 compute-sanitizer --tool $T --print-limit 5 python tools/profile_step.py --steps 2 $ARGS 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Invalid|hazard|out of bounds|=========     at|step 1" | head -8
  done
done

#!/bin/bash
# round 2: compute-sanitizer over the update kernels (tools/sanitize.sh) plus memcheck over the births, the deferred
# cross-GPU sums (contexts on one device) and the host-facing step
bash tools/sanitize.sh
for SEL in "tests/test_gpu_births.py" "tests/test_gpu_multi.py -k deferred" "tests/test_gpu_parity.py -k update_host"; do
  echo "== memcheck pytest $SEL"
  timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest $SEL -q -m gpu -p no:cacheprovider 2>&1 | grep -E "ERROR SUMMARY|Invalid|out of bounds|=========     at|passed|failed" | head -8
done

#!/bin/bash
# quick GPU check of a kernel change: parity subset, merge statistics, C3 bench line
# usage: bash tools/r2_quick.sh <tag> [full]
TAG=${1:-q}
mkdir -p gpurun_out
if [ "$2" == "full" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log
else
  timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log
fi
tail -4 gpurun_out/pytest_$TAG.log
timeout 120 python tools/merge_stats.py 2>&1 | tail -1
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$TAG.json"))
print("C3 ms_per_step %.4f kernel_us %.1f frac %.4f" % (d["ms_per_step"], d["roofline"]["kernel_us"], d["roofline"]["frac"]))
PY

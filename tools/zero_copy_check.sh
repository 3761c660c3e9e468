#!/bin/bash
# One short GPU call: the host-facing step with pinned caller buffers read / written by the kernels directly
# (RFSB200_ZERO_COPY=1) against the staged copies (=0): parity tests, then the C3 bench line of each, then the whole
# GPU suite with whatever time is left.  usage: gpurun -- bash tools/zero_copy_check.sh
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
t0=$(date +%s)
timeout 45 python -m pytest tests/test_gpu_parity.py -x -q -k "zero_copy or update_host or (golden and sc_dense)" > gpurun_out/zc_pytest.log 2>&1
echo "pytest rc=$? t=$(( $(date +%s) - t0 ))s" | tee -a gpurun_out/zc_pytest.log
tail -3 gpurun_out/zc_pytest.log
for zc in 1 0; do
  RFSB200_ZERO_COPY=$zc timeout 30 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/zc_bench_$zc.json 2> gpurun_out/zc_bench_$zc.err
  echo "bench zc=$zc rc=$? t=$(( $(date +%s) - t0 ))s"
  python - <<P
import json
try:
    d = json.loads(open("gpurun_out/zc_bench_$zc.json").read().strip().splitlines()[-1])
    print("zc=$zc value %.4g e2e %.4g e2e_ms %.4f kernel_us %.1f" % (d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["roofline"]["kernel_us"]))
except Exception as e:
    print("zc=$zc no bench line:", e)
P
done
timeout 70 python -m pytest tests -m gpu -x -q > gpurun_out/zc_pytest_all.log 2>&1
echo "full gpu suite rc=$? t=$(( $(date +%s) - t0 ))s" | tee -a gpurun_out/zc_pytest_all.log
tail -3 gpurun_out/zc_pytest_all.log

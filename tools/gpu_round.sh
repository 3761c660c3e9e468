#!/bin/bash
# One GPU-box pass: parity tests, bench, ncu launch list, ncu full capture of the update kernel.
# usage (under gpurun): bash tools/gpu_round.sh <tag> [tests|notests]
TAG=${1:-r1}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$TAG.txt 2>&1
if [ "$2" != "notests" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1
  echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log
  tail -5 gpurun_out/pytest_$TAG.log
fi
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench rc=$?"; cat gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
for C in C2 C3mf C5 N1k N64k C3sparse; do
  timeout 300 python bench.py --steps 20 --warmup 5 --config $C > gpurun_out/bench_${C}_$TAG.json 2>> gpurun_out/bench_$TAG.err
  cat gpurun_out/bench_${C}_$TAG.json
done
# the host-facing step staged through copies, for comparison with the default (kernels read / write the pinned buffers)
RFSB200_ZERO_COPY=0 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_staged_$TAG.json 2>> gpurun_out/bench_$TAG.err
cat gpurun_out/bench_staged_$TAG.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err
cat gpurun_out/bench_ref_$TAG.json
# launch list of the same bench command (cold-cache, serialised: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch_$TAG.log 2>&1
# full capture of the dominant kernel
timeout 900 ncu --set full --clock-control none --import-source on -k regex:phd_update -s 2 -c 1 -f -o gpurun_out/prof_$TAG \
  python tools/profile_step.py --steps 4 > gpurun_out/ncu_full_$TAG.log 2>&1
tail -3 gpurun_out/ncu_full_$TAG.log
python __graft_entry__.py smoke 2>&1 | tail -4

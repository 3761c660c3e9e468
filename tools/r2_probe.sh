#!/bin/bash
# round-2 baseline probe: bench lines + merge statistics of C3
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_base_C3.json 2> gpurun_out/r2_base_C3.err
cat gpurun_out/r2_base_C3.json
timeout 120 python tools/merge_stats.py > gpurun_out/r2_merge_stats.txt 2>&1; cat gpurun_out/r2_merge_stats.txt
for nw in 8 12; do
  RFSB200_WARPS_PER_CTA=$nw timeout 120 python tools/merge_stats.py 2>&1 | tail -1
done

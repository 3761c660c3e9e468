#!/bin/bash
# round-2 evidence for profiles/: ncu launch list of the bench command + full captures of the dominant kernels
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv \
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --extras none > gpurun_out/r2_ncu_launch.log 2>&1
bash tools/r2_ncu.sh r2_C3
bash tools/r2_ncu.sh r2_C3mf --sc 0
bash tools/r2_ncu.sh r2_C5 --vp --N 4000 --nM 150 --nZ 12 --sc 0 --cap 192
bash tools/r2_ncu.sh r2_C2 --N 1000 --nM 100 --nZ 20 --sc 0
# the four reports together exceed what gpurun brings back: summarise here, keep the C3 report only
python tools/ncu_summary.py r2 > gpurun_out/r2_ncu_summary.log 2>&1
mkdir -p gpurun_out/profiles
cp profiles/r2_ncu_*_metrics.txt profiles/r2_ncu_*_stages.txt profiles/traffic.json gpurun_out/profiles/
rm -f gpurun_out/prof_r2_C3mf.ncu-rep gpurun_out/prof_r2_C5.ncu-rep gpurun_out/prof_r2_C2.ncu-rep

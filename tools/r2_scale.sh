#!/bin/bash
# scaling lines for profiles/: bash tools/r2_scale.sh <ngpus>   (weak: 8000 particles per GPU; strong: 8000 and 64000 in total)
N=$1
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 --no-stages "$@" 2>>gpurun_out/r2_scale_$N.err; }
run > gpurun_out/r2_bench_C3_${N}gpu_weak.json
run --scaling strong > gpurun_out/r2_bench_C3_${N}gpu_strong8k.json
run --scaling strong --particles-total 64000 > gpurun_out/r2_bench_C3_${N}gpu_strong64k.json
run --no-fused > gpurun_out/r2_bench_C3_${N}gpu_weak_nccl.json
python - <<PY
import json
for t in ("weak","strong8k","strong64k","weak_nccl"):
    d=json.load(open("gpurun_out/r2_bench_C3_${N}gpu_%s.json"%t))
    print("${N} GPUs %-10s value %.4e ms %.4f e2e_ms %.4f kernel_us %.1f check %s err %s" % (t, d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["kernel_us"], d.get("collective_check"), d.get("comm_error")))
PY

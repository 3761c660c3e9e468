"""2-GPU probe: per-step device and host times of the eager and the deferred exchange (torchrun --nproc-per-node 2)."""
import os, sys, time
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rfs_slam_b200  # noqa
from rfs_slam_b200 import capi, synth
from rfs_slam_b200.phd import PHDUpdater
from rfs_slam_b200.dist import ShardedUpdater
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
wl = synth.make_config("C3", shard_id=rank)
up = PHDUpdater(wl.N, gm_capacity=256, z_capacity=32, device=local, precision=32)
up.load_workload(wl)
sh = ShardedUpdater(up, device=torch.device("cuda", local), fused=True)
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
F = capi.UPDATE_NO_COMMIT
for mode in ("eager", "defer", "eager", "defer"):
    for _ in range(3):
        sh.step(wl.Z, flags=F, defer=(mode == "defer"))
    sh.resolve()
    torch.cuda.synchronize(); dist.barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(8)]
    host = []
    up.profile_begin(8)
    for k in range(8):
        flush.zero_()
        up.comm_barrier()
        ev[k][0].record()
        t0 = time.perf_counter()
        sh.step(wl.Z, flags=F, defer=(mode == "defer"))
        host.append(1e6 * (time.perf_counter() - t0))
        ev[k][1].record()
    sh.resolve()
    torch.cuda.synchronize()
    ku = up.profile_read()
    print(rank, mode, "step us", ["%.0f" % (1e3 * a.elapsed_time(b)) for a, b in ev], "kernel us", ["%.0f" % x for x in ku], "host us", ["%.0f" % h for h in host], flush=True)
    dist.barrier()
up.close()
dist.destroy_process_group()

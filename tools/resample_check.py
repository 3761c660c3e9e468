"""torchrun --nproc-per-node G tools/resample_check.py : resampling over all shards (ShardedUpdater.resample_global: all-gather
of weights, the reference's placement on the global order, one all-to-all of packed particle records) against ONE
process holding all particles: every rank must end up with exactly the block of the single-process result."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
import rfs_slam_b200  # noqa
from rfs_slam_b200 import capi, synth
from rfs_slam_b200.dist import ShardedUpdater, block_range
from rfs_slam_b200.phd import PHDUpdater

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
stream = torch.cuda.Stream(dev)
torch.cuda.set_stream(stream)
ok_all = True
for vp in (False, True):
    N = 256 * world
    if vp:
        wl = synth.make_vp_workload(N=N, nM=80, nZ=10, use_cluster_process=1, config_id=9)
    else:
        wl = synth.make_workload(N=N, nM=90, nZ=16, use_cluster_process=1, config_id=8, ragged=0.2)
    sh_wl = wl.shard(rank, world)
    up = PHDUpdater(sh_wl.N, gm_capacity=160, z_capacity=32, device=local, lmk_dim=wl.dim)
    up.load_workload(sh_wl)
    sh = ShardedUpdater(up, device=dev, fused=True)
    sh.step(wl.Z)                                  # committed update, weights normalised over all ranks
    src = sh.resample_global(0.4321)
    cnt, mean, cov, w = up.download_maps()
    poses = up.get_poses()
    mask, nfov = up.get_unused()
    pw = up.get_weights()
    moved = int((np.abs(src - np.arange(*block_range(N, rank, world))) > 0).sum())
    lo, hi = block_range(N, rank, world)
    remote = int(((src < lo) | (src >= hi)).sum())
    # single process over all particles (every rank computes it on its own GPU)
    one = PHDUpdater(N, gm_capacity=160, z_capacity=32, device=local, lmk_dim=wl.dim)
    one.load_workload(wl)
    one.set_stream(stream.cuda_stream)
    one.update(wl.Z)
    from rfs_slam_b200.dist import reference_resample_sources
    w1 = one.get_weights()
    src1 = reference_resample_sources(w1 / w1.sum(), 0.4321)
    one.resample(src1.astype(np.int32), weight=1.0)
    c1, m1, v1, ww1 = one.download_maps()
    p1 = one.get_poses()
    k1, f1 = one.get_unused()
    off = np.zeros(N + 1, np.int64); np.cumsum(c1, out=off[1:])
    a, b = int(off[lo]), int(off[hi])
    ok = (np.array_equal(src, src1[lo:hi]) and np.array_equal(cnt, c1[lo:hi]) and np.array_equal(mean, m1[a:b]) and
          np.array_equal(cov, v1[a:b]) and np.array_equal(w, ww1[a:b]) and np.array_equal(poses, p1[lo:hi]) and
          np.array_equal(mask, k1[lo:hi]) and np.all(pw == 1.0))
    print(f"rank {rank} {'VP' if vp else '2-D'}: {moved} of {hi - lo} slots changed, {remote} copies came from another rank -> "
          f"{'identical to the single-process result' if ok else 'MISMATCH'}", flush=True)
    ok_all = ok_all and ok
    dist.barrier()
    up.close(); one.close()
dist.destroy_process_group()
sys.exit(0 if ok_all else 1)

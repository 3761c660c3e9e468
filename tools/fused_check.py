"""torchrun --nproc-per-node N tools/fused_check.py : the fused in-kernel all-reduce over peer memory
against the NCCL all-reduce path, on the same sharded workload (bit-identical sums expected for N=2,
1e-12 relative for N>2 where NCCL's reduction order may differ)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
import rfs_slam_b200  # noqa
from rfs_slam_b200 import capi, synth
from rfs_slam_b200.dist import ShardedUpdater
from rfs_slam_b200.phd import PHDUpdater

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
stream = torch.cuda.Stream(dev)
torch.cuda.set_stream(stream)
wl = synth.make_workload(N=2000, nM=120, nZ=24, use_cluster_process=1, config_id=5, shard_id=rank)
out = []
for fused in (False, True):
    up = PHDUpdater(wl.N, gm_capacity=192, z_capacity=32, device=local)
    up.load_workload(wl)
    sh = ShardedUpdater(up, device=dev, fused=fused)
    for _ in range(5):
        sh.step(wl.Z, flags=capi.UPDATE_NO_COMMIT)
    up.synchronize()
    sums = sh.sums.cpu().numpy().copy()
    w = up.get_weights(1)
    err = up.comm_error()
    out.append((sums, w, err))
    dist.barrier()
    up.close()
# the deferred consumption of the sums (RFSB200_UPDATE_DEFER_NORMALIZE) over the real peer mailboxes: COMMITTED steps (the
# next update applies the open normalisation while it loads the weights) against the eager fused steps, bit for bit —
# weights after every second step (closed by the read) and after the last one (closed by rfsb200_comm_resolve), and the maps
res = {}
for mode in ("eager", "deferred"):
    up = PHDUpdater(wl.N, gm_capacity=192, z_capacity=32, device=local)
    up.load_workload(wl)
    sh = ShardedUpdater(up, device=dev, fused=True)
    ws = []
    for k in range(6):
        Zk = wl.Z.reshape(-1, 2) + 0.003 * k
        sh.step(Zk, defer=(mode == "deferred"))
        if k % 2 == 1:
            ws.append(up.get_weights().copy())
    sh.step(wl.Z, defer=(mode == "deferred"))
    sh.resolve()
    ws.append(up.get_weights().copy())
    res[mode] = (ws, up.download_maps(), up.comm_error())
    dist.barrier()
    up.close()
same_w = all(np.array_equal(a, b) for a, b in zip(res["eager"][0], res["deferred"][0]))
same_m = all(np.array_equal(a, b) for a, b in zip(res["eager"][1], res["deferred"][1]))
defer_ok = same_w and same_m and not res["deferred"][2] and not res["eager"][2]
print(f"rank {rank}: deferred sums against eager over {len(res['eager'][0])} reads: weights identical {same_w}, maps identical {same_m}", flush=True)
tot = torch.tensor([out[1][1].sum()], dtype=torch.float64, device=dev)
dist.all_reduce(tot)
ok = (not out[1][2]) and np.allclose(out[0][0], out[1][0], rtol=1e-12) and np.allclose(out[0][1], out[1][1], rtol=1e-12) \
    and abs(tot.item() - 1.0) < 1e-12 and defer_ok
print(f"rank {rank}: nccl sums {out[0][0]} fused sums {out[1][0]} bit-identical {np.array_equal(out[0][0], out[1][0])} "
      f"weights match {np.allclose(out[0][1], out[1][1], rtol=1e-12)} global sum of weights {tot.item():.15f} -> {'OK' if ok else 'FAIL'}", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)

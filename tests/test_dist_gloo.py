"""N>1 path on CPU: two gloo ranks, particles block-partitioned, ONE all-reduce of [sum w, sum w^2]
per step (rfs_slam_b200.dist).  The per-shard posterior comes from the oracle here (no GPU in this
test); what is checked is the sharding + collective + normalisation logic that the GPU ranks run."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import rfs_slam_b200  # noqa: F401
    from rfs_slam_b200 import dist as rd, synth
    from oracle import binding as ob
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    wl = synth.make_workload(N=101, nM=40, nZ=10, use_cluster_process=1, config_id=77)
    lo, hi = rd.block_range(wl.N, rank, world)
    sh = wl.shard(rank, world)
    assert sh.N == hi - lo
    r = ob.run(sh, sort_mode=ob.SORT_STABLE)                      # this rank's particles only
    sums = torch.tensor([r.weight.sum(), (r.weight ** 2).sum()], dtype=torch.float64)
    rd.allreduce_sums(sums)                                        # the only data-path collective
    wn, ess = rd.normalise_and_ess(r.weight, sums)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), lo=lo, hi=hi, wn=wn, ess=ess, sums=sums.numpy(),
             count=r.count)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_update_matches_single_process(tmp_path, world):
    import rfs_slam_b200  # noqa: F401
    from rfs_slam_b200 import synth
    from oracle import binding as ob
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    wl = synth.make_workload(N=101, nM=40, nZ=10, use_cluster_process=1, config_id=77)
    ref = ob.run(wl, sort_mode=ob.SORT_STABLE)
    s1, s2 = ref.weight.sum(), (ref.weight ** 2).sum()
    wn_ref = ref.weight / s1
    got = np.zeros(wl.N)
    cnt = np.zeros(wl.N, dtype=np.int64)
    covered = 0
    for r in range(world):
        z = np.load(tmp_path / f"rank{r}.npz")
        lo, hi = int(z["lo"]), int(z["hi"])
        assert lo == covered
        covered = hi
        got[lo:hi] = z["wn"]
        cnt[lo:hi] = z["count"]
        assert np.allclose(z["sums"], [s1, s2], rtol=1e-12)      # every rank holds the global sums
        assert float(z["ess"]) == pytest.approx(s1 * s1 / s2, rel=1e-12)
    assert covered == wl.N
    assert np.allclose(got, wn_ref, rtol=1e-12)
    assert got.sum() == pytest.approx(1.0, abs=1e-12)
    assert np.array_equal(cnt, ref.count)                          # maps do not depend on the sharding


def test_block_range_covers_everything():
    import rfs_slam_b200  # noqa: F401
    from rfs_slam_b200 import dist as rd
    for n in (1, 7, 8000, 64000, 64001):
        for w in (1, 2, 3, 4, 8):
            edges = [rd.block_range(n, r, w) for r in range(w)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            for a, b in zip(edges[:-1], edges[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1

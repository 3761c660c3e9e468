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


# ---- resampling across shards: host logic (placement + exchange plan) ------------------------------------------
def _reference_resample_loop(w, r01):
    """the reference's loops, literally (include/ParticleFilter.hpp:419-479)"""
    n = len(w)
    interval = 1.0 / float(n)
    sample_point = interval * r01
    idx = 0
    cum = w[0]
    flag = [0] * n
    sampled = [0] * n
    for i in range(n):
        while sample_point > cum and idx < n - 1:
            idx += 1
            cum += w[idx]
        sampled[i] = idx
        flag[idx] = 1
        sample_point += interval
    src = list(range(n))
    idx_prev, nxt = 0, 0
    for i in range(n):
        idx = sampled[i]
        first = not (i > 0 and idx == idx_prev)
        idx_prev = idx
        if not (idx < n and first):
            while flag[nxt] == 1:
                nxt += 1
            src[nxt] = idx
            nxt += 1
    return np.array(src)


def test_resample_placement_matches_reference_loops():
    from rfs_slam_b200 import dist as rd
    rng = np.random.default_rng(3)
    for n in (1, 2, 7, 64, 500):
        for trial in range(4):
            w = rng.random(n) ** (1 + 3 * trial)
            w /= w.sum()
            r01 = float(rng.random())
            assert np.array_equal(rd.reference_resample_sources(w, r01), _reference_resample_loop(w, r01))


def _exchange_worker(rank, world, port, n_total, seed, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import rfs_slam_b200  # noqa: F401
    from rfs_slam_b200 import dist as rd
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(seed)
    w = rng.random(n_total) ** 4
    w /= w.sum()
    src = rd.reference_resample_sources(w, 0.37)
    lo, hi = rd.block_range(n_total, rank, world)
    payload = np.arange(lo, hi, dtype=np.int64) * 10 + 7          # a "particle" is its global id here
    local_src, send, recv = rd.exchange_plan(src, rank, world)
    sbuf = torch.from_numpy(np.concatenate([payload[s] for s in send]) if sum(map(len, send)) else np.zeros(0, np.int64))
    rbuf = torch.empty(sum(map(len, recv)), dtype=torch.int64)
    dist.all_to_all_single(rbuf, sbuf, [len(x) for x in recv], [len(x) for x in send])
    new = payload[local_src].copy()
    if len(rbuf):
        new[np.concatenate(recv)] = rbuf.numpy()
    ok = np.array_equal(new, src[lo:hi] * 10 + 7)
    q.put((rank, bool(ok), int(sum(map(len, send)))))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_exchange_plan_moves_every_copy_to_its_slot(world):
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    n_total = 60 * world
    ps = [ctx.Process(target=_exchange_worker, args=(r, world, port, n_total, 11, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in ps:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    assert sum(n for _, _, n in res) > 0          # some copies did change rank

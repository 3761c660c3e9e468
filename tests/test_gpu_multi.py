"""Multi-GPU paths on one node (skipped with fewer than 2 GPUs): the fused in-kernel cross-GPU weight sum against the
NCCL all-reduce, and resampling over all shards against one process holding all particles.  Each check is a
torchrun job of tools/*.py (one process per GPU, rendezvous on 127.0.0.1)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _torchrun(script, n, port):
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", script)],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    return r.returncode, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
def test_fused_allreduce_equals_nccl(cuda_required):
    rc, out = _torchrun("fused_check.py", 2, 29531)
    assert rc == 0, out


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
def test_global_resampling_is_independent_of_the_number_of_ranks(cuda_required):
    rc, out = _torchrun("resample_check.py", 2, 29532)
    assert rc == 0, out
    assert out.count("identical to the single-process result") == 4, out


@pytest.mark.parametrize("dim", [2, 3], ids=["rngbrg", "victoriapark"])
def test_particle_exchange_between_two_contexts_on_one_gpu(cuda_required, dim):
    """The cross-GPU resampling path with both "ranks" as contexts on ONE device (so that it runs, and is recorded, on a
    single-GPU box): the reference's placement on the global particle order (dist.reference_resample_sources), the
    exchange plan of each rank (dist.exchange_plan), packed particle records from rfsb200_export_particles into a device
    buffer, the local gather (rfsb200_resample), rfsb200_import_particles — against one context that holds all particles
    and resamples alone: maps, poses, unused-measurement masks and weights identical bit for bit."""
    import numpy as np
    import torch
    from rfs_slam_b200 import capi, synth
    from rfs_slam_b200.dist import block_range, exchange_plan, reference_resample_sources
    from rfs_slam_b200.phd import PHDUpdater
    mk = synth.make_vp_workload if dim == 3 else synth.make_workload
    wl = mk(N=96, nM=50, nZ=10, use_cluster_process=1, config_id=91 + dim)
    caps = dict(gm_capacity=128, z_capacity=16, lmk_dim=dim)

    def fresh(sub):
        up = PHDUpdater(sub.N, **caps)
        up.load_workload(sub)
        up.update(sub.Z, flags=capi.UPDATE_NO_NORMALIZE)   # maps, masks and weights of a real step
        return up

    one = fresh(wl)
    w = one.get_weights()
    w = w / w.sum()
    src = reference_resample_sources(w, 0.37)
    assert len(np.unique(src)) < wl.N, "the draw must duplicate some particles"
    one.resample(src.astype(np.int32), weight=1.0)
    ref = (one.download_maps(), one.get_poses(), one.get_unused(), one.get_weights())

    world = 2
    ranks = [fresh(wl.shard(r, world)) for r in range(world)]
    plans = [exchange_plan(src, r, world) for r in range(world)]
    rec = ranks[0].particle_record_bytes()
    assert rec == ranks[1].particle_record_bytes() and rec > 0
    bufs = {}
    for r in range(world):            # every rank packs what the others need, from its state BEFORE the local copies
        local_src, send, recv = plans[r]
        for dst in range(world):
            if len(send[dst]):
                b = torch.empty(len(send[dst]) * rec, dtype=torch.uint8, device="cuda")
                ranks[r].export_particles(send[dst], b.data_ptr())
                ranks[r].synchronize()
                bufs[(r, dst)] = b
    n_moved = 0
    for r in range(world):
        local_src, send, recv = plans[r]
        ranks[r].resample(local_src, weight=1.0)
        for s in range(world):
            if len(recv[s]):
                assert len(recv[s]) * rec == bufs[(s, r)].numel()
                ranks[r].import_particles(recv[s], bufs[(s, r)].data_ptr(), 1.0)
                n_moved += len(recv[s])
        ranks[r].synchronize()
    assert n_moved > 0, "no particle changed rank: the exchange was not exercised"
    got_cnt, got_mean, got_cov, got_w = [], [], [], []
    for r in range(world):
        c, m, cv, ww = ranks[r].download_maps()
        got_cnt.append(c); got_mean.append(m); got_cov.append(cv); got_w.append(ww)
    (rc, rm, rcv, rw), rpose, (rmask, rnfov), rweights = ref
    assert np.array_equal(np.concatenate(got_cnt), rc)
    assert np.array_equal(np.concatenate(got_mean), rm) and np.array_equal(np.concatenate(got_cov), rcv)
    assert np.array_equal(np.concatenate(got_w), rw)
    assert np.array_equal(np.concatenate([ranks[r].get_poses() for r in range(world)]), rpose)
    assert np.array_equal(np.concatenate([ranks[r].get_unused()[0] for r in range(world)]), rmask)
    assert np.array_equal(np.concatenate([ranks[r].get_weights() for r in range(world)]), rweights)
    for u in ranks + [one]:
        u.close()


@pytest.mark.parametrize("dim", [2, 3], ids=["rngbrg", "victoriapark"])
@pytest.mark.parametrize("world", [2, 3])
def test_deferred_cross_gpu_sums_between_contexts_on_one_gpu(cuda_required, world, dim):
    """RFSB200_UPDATE_DEFER_NORMALIZE with the "ranks" as contexts of this process on ONE device (rfsb200_comm_connect_local),
    driven one after the other — possible because a deferred step never waits for a peer that is running at the same time:
    it sends its [sum w, sum w^2] and ends; the next step finds the pairs of the previous epoch in its mailbox and divides
    the weights by their total while it loads them.  Against the same shards stepped with the normalisation done between
    the steps (local sums added in rank order on the host, w / total written back): weights and maps identical bit for
    bit after every step, whether the open normalisation is closed by the next update, by rfsb200_comm_resolve or by a
    reader of the weights."""
    import numpy as np
    from rfs_slam_b200 import capi, synth
    from rfs_slam_b200.phd import PHDUpdater
    mk = synth.make_vp_workload if dim == 3 else synth.make_workload
    wl = mk(N=90, nM=40, nZ=10, use_cluster_process=1, config_id=97)
    shards = [wl.shard(r, world) for r in range(world)]
    rng = np.random.default_rng(5)
    Zs = [wl.Z.reshape(-1, dim) + rng.normal(0, 0.01, (wl.nZ, dim)) for _ in range(5)]

    def fresh():
        ups = []
        for sh in shards:
            u = PHDUpdater(sh.N, gm_capacity=128, precision=32, z_capacity=16, lmk_dim=dim)
            u.load_workload(sh)
            ups.append(u)
        return ups

    # path B: eager arithmetic spelled out on the host
    B = fresh()
    want = []
    for Z in Zs:
        local = [u.update(Z, flags=capi.UPDATE_NO_NORMALIZE, want_stats=True).sum_w for u in B]
        total = 0.0
        for v in local:
            total += v
        ws = []
        for u, sh in zip(B, shards):
            w = u.get_weights() / total
            u.set_poses(sh.pose, sh.pose_cov, w)
            ws.append(w)
        want.append((np.concatenate(ws), [u.download_maps() for u in B]))

    A = fresh()
    for r, u in enumerate(A):
        u.comm_connect_local(r, world, A)
    F = capi.UPDATE_FUSED_ALLREDUCE | capi.UPDATE_DEFER_NORMALIZE
    for k, Z in enumerate(Zs):
        for u in A:
            u.update(Z, flags=F)          # step k of every rank; the previous step's sums are applied on the way in
        if k == 1:                        # closed explicitly ...
            for u in A:
                u.comm_resolve()
        if k in (1, 2, 4):                # ... or by reading the weights (k = 2, 4); k = 0, 3: by the next update
            got = np.concatenate([u.get_weights() for u in A])
            assert np.array_equal(got, want[k][0]), k
            assert got.sum() == pytest.approx(1.0, abs=1e-12)
        for u, (cnt, mean, cov, w) in zip(A, want[k][1]):
            c2, m2, cv2, w2 = u.download_maps()
            assert np.array_equal(c2, cnt) and np.array_equal(m2, mean) and np.array_equal(cv2, cov) and np.array_equal(w2, w), k
    assert not any(u.comm_error() for u in A)
    # an eager fused step cannot follow on one device (it would wait for a peer that cannot run): the flag combinations
    # that make no sense are refused
    with pytest.raises(RuntimeError):
        A[0].update(Zs[0], flags=capi.UPDATE_DEFER_NORMALIZE)
    # uncommitted deferred steps (what bench.py times): the open weights are those of the back buffer; the next step
    # overwrites them after picking the pairs up, a reader of the result closes them
    for u in A + B:
        u.comm_resolve()
    for k in range(3):
        for u in A:
            u.update(Zs[k], flags=F | capi.UPDATE_NO_COMMIT)
    got = np.concatenate([u.get_weights(1) for u in A])
    local = [u.update(Zs[2], flags=capi.UPDATE_NO_NORMALIZE | capi.UPDATE_NO_COMMIT, want_stats=True).sum_w for u in B]
    total = 0.0
    for v in local:
        total += v
    assert np.array_equal(got, np.concatenate([u.get_weights(1) / total for u in B]))
    assert not any(u.comm_error() for u in A)
    for u in A + B:
        u.close()

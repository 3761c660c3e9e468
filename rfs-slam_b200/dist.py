"""One process per GPU: particles are block-partitioned across ranks (each rank owns a
PHDUpdater on its own shard) and the only exchange per step is ONE all-reduce (SUM) of the two
fp64 scalars [sum w, sum w^2] of the unnormalised particle weights (replaces the
serial loops of ParticleFilter::normalizeWeights / the ESS test, include/ParticleFilter.hpp:352-363,
406-411).  torch.distributed is plumbing only: NCCL on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

import os

import numpy as np


def block_range(n_total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block partition of particle indices: rank g owns [g*N/G, (g+1)*N/G)."""
    return rank * n_total // world, (rank + 1) * n_total // world


def env_rank_world() -> tuple[int, int, int]:
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


class _CudaArrayView:
    """Wrap a raw device pointer so torch.as_tensor can alias it (no copy)."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = dict(shape=tuple(shape), typestr=typestr, data=(int(ptr), False), version=3)


def device_tensor_from_ptr(ptr: int, n: int, device):
    """Alias `n` fp64 values at device address `ptr` as a torch tensor on `device`."""
    import torch
    return torch.as_tensor(_CudaArrayView(ptr, (n,), "<f8"), device=device)


def allreduce_sums(sums, group=None):
    """In-place SUM all-reduce of the [sum w, sum w^2] pair (a torch tensor, CPU or CUDA)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    return sums


def normalise_and_ess(weights: np.ndarray, sums) -> tuple[np.ndarray, float]:
    """Host mirror of normalizeWeights + N_eff = 1 / sum(w_hat^2) = (sum w)^2 / sum w^2."""
    s1, s2 = float(sums[0]), float(sums[1])
    return weights / s1, (s1 * s1 / s2 if s2 > 0 else 0.0)


# ---- resampling across shards ---------------------------------------------------------
def reference_resample_sources(w: np.ndarray, r01: float) -> np.ndarray:
    """Placement of ParticleFilter::resample() (include/ParticleFilter.hpp:419-479) for normalised weights w and the
    single uniform draw r01: src[slot] = index of the particle whose copy ends up in `slot`.  Systematic sampling
    (sample_point += 1/n, idx advances while sample_point > cumulative weight); a sampled particle keeps its own slot
    the first time it is drawn, every further copy goes to the next slot nobody was drawn from, in ascending order."""
    w = np.asarray(w, dtype=np.float64)
    n = len(w)
    interval = 1.0 / float(n)
    steps = np.full(n, interval)
    steps[0] = interval * float(r01)
    sample_points = np.cumsum(steps)              # the reference accumulates sample_point += interval
    cum = np.cumsum(w)                            # ... and cumulative_weight += w[idx], in this order
    sampled = np.minimum(np.searchsorted(cum, sample_points, side="left"), n - 1)
    counts = np.bincount(sampled, minlength=n)
    src = np.arange(n, dtype=np.int64)
    free = np.nonzero(counts == 0)[0]
    extras = np.repeat(np.arange(n), np.maximum(counts - 1, 0))
    src[free] = extras                            # both ascending, as the reference's next_unsampled_idx walk
    return src


def exchange_plan(src_global: np.ndarray, rank: int, world: int):
    """What rank `rank` has to do for a global placement src_global[slot] with particles block-partitioned over
    `world` ranks: (local_src [n_local] with own index as placeholder for remote sources, send = list per destination
    rank of LOCAL particle indices to export in that order, recv = list per source rank of LOCAL slots to import into
    in that order).  Sender and receiver enumerate the slots of a (source rank, destination rank) pair in ascending
    slot order, so no indices travel with the payload."""
    n = len(src_global)
    bounds = [block_range(n, g, world) for g in range(world)]
    owner = np.zeros(n, dtype=np.int64)
    for g, (lo, hi) in enumerate(bounds):
        owner[lo:hi] = g
    lo, hi = bounds[rank]
    slots = np.arange(n)
    src_owner = owner[src_global]
    local_src = np.arange(hi - lo, dtype=np.int32)
    mine = slots[lo:hi]
    is_local = src_owner[lo:hi] == rank
    local_src[is_local] = (src_global[lo:hi][is_local] - lo).astype(np.int32)
    send, recv = [], []
    for g in range(world):
        if g == rank:
            send.append(np.zeros(0, np.int32)); recv.append(np.zeros(0, np.int32))
            continue
        glo, ghi = bounds[g]
        to_g = slots[glo:ghi][src_owner[glo:ghi] == rank]          # slots of rank g fed by my particles
        send.append((src_global[to_g] - lo).astype(np.int32))
        from_g = mine[src_owner[lo:hi] == g]                        # my slots fed by rank g's particles
        recv.append((from_g - lo).astype(np.int32))
    return local_src, send, recv


class ShardedUpdater:
    """RBPHDFilter::update() over particles sharded across ranks.

    step(Z): local fused update kernel (no normalisation) -> all-reduce of the two weight sums on
    the same CUDA stream -> local normalisation kernel.  No other data-path collective exists."""

    def __init__(self, updater, device=None, group=None, fused: bool = False):
        import torch
        import torch.distributed as dist
        self.up = updater
        self.group = group
        self.fused = False
        if fused and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            # exchange the CUDA IPC handles of the mailboxes once; afterwards the update kernel does the
            # cross-GPU sum itself through peer memory (RFSB200_UPDATE_FUSED_ALLREDUCE)
            world, rank = dist.get_world_size(group), dist.get_rank(group)
            handles = [None] * world
            dist.all_gather_object(handles, updater.comm_export(), group=group)
            ok = 1
            try:
                updater.comm_connect(rank, world, handles)
            except Exception as e:   # e.g. no peer access between two of the GPUs
                import sys
                print(f"[rfs_slam_b200.dist] rank {rank}: peer mailboxes unavailable ({e}); using the NCCL all-reduce", file=sys.stderr)
                ok = 0
            flag = torch.tensor([ok], dtype=torch.int32, device=device if device is not None else "cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)   # all ranks take the same path
            fused = bool(flag.item())
        self.fused = bool(fused)   # on one rank the flag still saves the normalisation launch
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.sums = device_tensor_from_ptr(updater.weight_sums_device_ptr(), 2, self.device)
        # run the library on torch's current stream so kernels and the collective are ordered
        self.up.set_stream(torch.cuda.current_stream(self.device).cuda_stream)

    def step(self, Z, flags: int = 0, want_stats: bool = False, defer: bool = False):
        """defer (fused path only): the kernel sends its sums and ends without waiting for the peers'; the next step() picks
        them up while it loads the weights (bit-identical results), resolve() or any reader of the weights closes the last
        one.  For the steps between two resampling decisions, where nobody looks at the weights."""
        from . import capi
        if self.fused:   # one launch: update + cross-GPU sum over NVLink + normalisation
            if defer and not want_stats:
                return self.up.update(Z, flags=flags | capi.UPDATE_FUSED_ALLREDUCE | capi.UPDATE_DEFER_NORMALIZE, want_stats=False)
            so = self.up.update(Z, flags=flags | capi.UPDATE_FUSED_ALLREDUCE, want_stats=want_stats)
            if so is not None and not np.isfinite(so.sum_w):
                self.check_comm()
            return so
        so = self.up.update(Z, flags=flags | capi.UPDATE_NO_NORMALIZE, want_stats=want_stats)
        allreduce_sums(self.sums, self.group)
        self.up.normalize()
        return so

    def step_host(self, pose, pose_cov, weight, Z, flags: int = 0, w_out=None, unused_out=None, nfov_out=None):
        """The host-facing step (what the drop-in RBPHDFilter::update() moves per step): poses / weights / Z in,
        normalised weights / unused-measurement masks / in-FOV counts out.  Fused path: ONE ABI call with one
        synchronisation; NCCL path: the same data through the separate calls around the collective."""
        from . import capi
        if self.fused:
            so = self.up.update_host(pose, pose_cov, weight, Z, flags=flags | capi.UPDATE_FUSED_ALLREDUCE,
                                     w_out=w_out, unused_out=unused_out, nfov_out=nfov_out)
            if w_out is not None and len(w_out) and w_out[0] != w_out[0]:   # NaN weights: a peer missed the exchange
                self.check_comm()
            return so
        self.up.set_poses(pose, pose_cov, weight)
        self.step(Z, flags=flags)
        which = 1 if (flags & capi.UPDATE_NO_COMMIT) else 0
        if w_out is not None:
            self.up.get_weights(which, out=w_out)
        if unused_out is not None or nfov_out is not None:
            self.up.get_unused(unused_out, nfov_out)
        return None


    def resolve(self, check: bool = False):
        """Closes a deferred step: afterwards the weights are normalised and the sums are the global ones.  check: wait for
        the stream and raise if a peer missed one of the exchanges since the last check (deferred steps are asynchronous,
        so nothing else looks at rfsb200_comm_error in between)."""
        if self.fused:
            self.up.comm_resolve()
            if check:
                self.up.synchronize()
                self.check_comm()

    def check_comm(self):
        """Raises if a peer did not arrive in some fused update since the last check (the sums of that step are NaN).
        step() / step_host() call it whenever they see non-finite results; an asynchronous caller (step() without
        statistics) should call it before it consumes the weights."""
        if self.fused and self.up.comm_error():
            raise RuntimeError("rfs_slam_b200.dist: a peer GPU did not reach the in-kernel weight-sum exchange in time "
                               "(RFSB200_COMM_TIMEOUT_MS); the particle weights of that step are NaN — rerun the step with "
                               "fused=False (NCCL all-reduce) or raise the timeout")

    # ---- resampling over ALL shards -----------------------------------------------------------------------------
    def resample_global(self, r01: float, neff_threshold: float | None = None):
        """ParticleFilter::resample() over the particles of all ranks: all-gather of the normalised weights, the
        reference's systematic-sampling placement on the global index order (identical to what one process holding all
        particles would do, so the result does not depend on the number of ranks), local copies on the device and ONE
        all-to-all of packed particle records (maps, pose, unused-measurement mask) for the copies that change rank.
        r01 must be the same number on every rank (one drand48() on rank 0, broadcast).  Returns the global source
        index of every local slot, or None if N_eff is above neff_threshold (then the weights are left normalised)."""
        import torch
        import torch.distributed as dist
        up = self.up
        world = dist.get_world_size(self.group) if (dist.is_available() and dist.is_initialized()) else 1
        rank = dist.get_rank(self.group) if world > 1 else 0
        w_local = up.get_weights()
        if world > 1:
            t = torch.from_numpy(w_local).to(self.device)
            parts = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(parts, t, group=self.group)          # equal shard sizes (block partition of a multiple of world)
            w = torch.cat(parts).cpu().numpy()
        else:
            w = w_local
        w = w / w.sum()
        if neff_threshold is not None and 1.0 / float((w * w).sum()) > neff_threshold:
            return None
        src = reference_resample_sources(w, r01)
        n_local = up.N
        if world == 1:
            up.resample(src.astype(np.int32), weight=1.0)
            return src
        assert len(w) == n_local * world, "resample_global needs equal shards"
        local_src, send, recv = exchange_plan(src, rank, world)
        rec = up.particle_record_bytes()
        n_send = sum(len(x) for x in send)
        n_recv = sum(len(x) for x in recv)
        sbuf = torch.empty(max(1, n_send) * rec, dtype=torch.uint8, device=self.device)
        rbuf = torch.empty(max(1, n_recv) * rec, dtype=torch.uint8, device=self.device)
        if n_send:
            up.export_particles(np.concatenate(send), sbuf.data_ptr())     # from the state BEFORE the local copies
        dist.all_to_all_single(rbuf[:n_recv * rec], sbuf[:n_send * rec], [len(x) * rec for x in recv],
                               [len(x) * rec for x in send], group=self.group)
        up.resample(local_src, weight=1.0)
        if n_recv:
            up.import_particles(np.concatenate(recv), rbuf.data_ptr(), 1.0)
        lo, hi = block_range(len(w), rank, world)
        return src[lo:hi]

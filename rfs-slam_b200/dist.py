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

    def step(self, Z, flags: int = 0, want_stats: bool = False):
        from . import capi
        if self.fused:   # one launch: update + cross-GPU sum over NVLink + normalisation
            return self.up.update(Z, flags=flags | capi.UPDATE_FUSED_ALLREDUCE, want_stats=want_stats)
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
            return self.up.update_host(pose, pose_cov, weight, Z, flags=flags | capi.UPDATE_FUSED_ALLREDUCE,
                                       w_out=w_out, unused_out=unused_out, nfov_out=nfov_out)
        self.up.set_poses(pose, pose_cov, weight)
        self.step(Z, flags=flags)
        which = 1 if (flags & capi.UPDATE_NO_COMMIT) else 0
        if w_out is not None:
            self.up.get_weights(which, out=w_out)
        if unused_out is not None or nfov_out is not None:
            self.up.get_unused(unused_out, nfov_out)
        return None

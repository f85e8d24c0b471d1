"""Sharding and gathers for the multi-GPU path.

The hot path shards with no exchange: every configuration / trajectory is independent and the robot constants are a few
KB, so each rank (one process per GPU) evaluates a contiguous block of whole trajectories.  Collectives appear only where
the path has a real exchange step: gathering the per-trajectory costs (all-gather, ceil(B/G) doubles per rank) and the
minimum sphere distance (all-reduce(min)).  torch.distributed is plumbing here (NCCL over NVLink on the GPUs, gloo in the
CPU tests); the tensors handed in wrap the device buffers the C-ABI wrote.
"""
import numpy as np


def shard_range(n_units, rank, world_size):
    """Contiguous block [begin, end) of the n_units trajectories (or configurations) that `rank` evaluates.
    Blocks have ceil(n / world) units except the trailing ranks, which may hold fewer (or none)."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank %r / world size %r" % (rank, world_size))
    per = -(-int(n_units) // world_size)
    begin = min(int(n_units), rank * per)
    return begin, min(int(n_units), begin + per)


def shard_evaluations(n_traj, T, rank, world_size):
    """Evaluation range [e0, e1) of a rank for trajectories of T timesteps: trajectories are never split, so the
    transition term x[t] - x[t-1] stays local to a rank."""
    b0, b1 = shard_range(n_traj, rank, world_size)
    return b0 * T, b1 * T


def gather_per_trajectory(local, n_traj, group=None):
    """All-gather the per-trajectory values (costs) of every rank into one [n_traj] tensor on every rank.
    `local`: 1-D torch tensor with this rank's block (device tensor under NCCL, CPU tensor under gloo)."""
    import torch
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized():
        return local.clone()
    world = dist.get_world_size(group)
    per = -(-int(n_traj) // world)
    padded = torch.zeros(per, dtype=local.dtype, device=local.device)
    padded[:local.numel()] = local
    out = torch.empty(per * world, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    return out[:n_traj]


def allreduce_min(local_min, group=None):
    """Global minimum (e.g. of the sphere distance) over all ranks; ranks without data pass +inf."""
    import torch
    import torch.distributed as dist

    t = local_min if isinstance(local_min, torch.Tensor) else torch.tensor([float(local_min)], dtype=torch.float64)
    t = t.clone()
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    return t


def split_states(X, T, rank, world_size):
    """X[B][T][N] (or [B*T][N]) -> this rank's block of whole trajectories, as a view."""
    X = np.asarray(X)
    flat = X.reshape(-1, X.shape[-1])
    e0, e1 = shard_evaluations(flat.shape[0] // T, T, rank, world_size)
    return flat[e0:e1]

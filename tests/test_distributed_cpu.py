"""Multi-rank host logic on CPU: world_size 2 over gloo (the GPU path uses the same helpers over NCCL)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from exotica_b200 import distributed as exd


def test_shard_ranges_cover_everything_without_splitting_trajectories():
    for n, world in ((4096, 8), (10, 4), (3, 8), (0, 2), (7, 1)):
        blocks = [exd.shard_range(n, r, world) for r in range(world)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        for (a0, a1), (b0, b1) in zip(blocks, blocks[1:]):
            assert a1 == b0 and a0 <= a1
    assert exd.shard_evaluations(10, 100, 1, 4) == (300, 600)
    with pytest.raises(ValueError):
        exd.shard_range(5, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_traj, T, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # every rank "evaluates" its block: cost of trajectory b = b + 0.5, min distance of (b, t) = 10 - b + 0.01 t
        b0, b1 = exd.shard_range(n_traj, rank, world)
        local_cost = torch.arange(b0, b1, dtype=torch.float64) + 0.5
        costs = exd.gather_per_trajectory(local_cost, n_traj)
        bt = 10.0 - torch.arange(b0, b1, dtype=torch.float64)[:, None] + 0.01 * torch.arange(T, dtype=torch.float64)[None, :]
        local_min = bt.min() if bt.numel() else torch.tensor(float("inf"), dtype=torch.float64)
        gmin = exd.allreduce_min(local_min.reshape(1))
        X = np.arange(n_traj * T * 3, dtype=np.float64).reshape(n_traj, T, 3)
        mine = exd.split_states(X, T, rank, world)
        ok = (costs.numpy().tolist() == [b + 0.5 for b in range(n_traj)] and abs(float(gmin) - (10.0 - (n_traj - 1))) < 1e-12 and
              mine.shape == ((b1 - b0) * T, 3) and (mine.size == 0 or mine[0, 0] == b0 * T * 3))
        out[rank] = 1 if ok else 0
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_traj", [7, 2, 1])
def test_gathers_world_size_2_gloo(n_traj):
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Array("i", [0] * world)
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_traj, 5, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert list(out) == [1, 1]

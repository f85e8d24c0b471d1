"""Multi-rank host logic on CPU, world_size 2: the sharding helpers, the slot layout of the gathered arrays (checked with a
gloo all-gather standing in for exo_comm_allgather, which needs GPUs), and the file rendezvous that ships the NCCL unique id
from rank 0 to the other ranks."""
import multiprocessing
import os
import socket
import tempfile

import numpy as np
import pytest

from exotica_b200 import distributed as exd


def test_shard_ranges_cover_everything_without_splitting_trajectories():
    for n, world in ((4096, 8), (10, 4), (3, 8), (0, 2), (7, 1)):
        blocks = [exd.shard_range(n, r, world) for r in range(world)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        for (a0, a1), (b0, b1) in zip(blocks, blocks[1:]):
            assert a1 == b0 and a0 <= a1
        assert all(b1 - b0 <= exd.units_per_rank(n, world) for b0, b1 in blocks)
    assert exd.shard_evaluations(10, 100, 1, 4) == (300, 600)
    with pytest.raises(ValueError):
        exd.shard_range(5, 2, 2)


def test_product_package_has_no_torch():
    """north_star: the Python bindings forward to the C-ABI, no PyTorch anywhere in the product package."""
    import re
    pkg = os.path.dirname(exd.__file__)
    for name in os.listdir(pkg):
        if name.endswith(".py"):
            text = open(os.path.join(pkg, name)).read()
            assert not re.search(r"^\s*(import|from)\s+torch\b", text, flags=re.M), name


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_traj, T, out):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # every rank "evaluates" its block: cost of trajectory b = b + 0.5, min distance of (b, t) = 10 - b + 0.01 t.  The gathered
        # array has one slot of ceil(B / world) entries per rank (exo_problem_cost_allgather pads with 0, the distance with +inf)
        b0, b1 = exd.shard_range(n_traj, rank, world)
        per = exd.units_per_rank(n_traj, world)
        slot = torch.zeros(per, dtype=torch.float64)
        slot[:b1 - b0] = torch.arange(b0, b1, dtype=torch.float64) + 0.5
        gathered = torch.empty(per * world, dtype=torch.float64)
        dist.all_gather_into_tensor(gathered, slot)
        costs = gathered[:n_traj]
        bt = 10.0 - torch.arange(b0, b1, dtype=torch.float64)[:, None] + 0.01 * torch.arange(T, dtype=torch.float64)[None, :]
        dslot = torch.full((per,), float("inf"), dtype=torch.float64)
        if bt.numel():
            dslot[:b1 - b0] = bt.min(dim=1).values
        dall = torch.empty(per * world, dtype=torch.float64)
        dist.all_gather_into_tensor(dall, dslot)
        gmin = dall.min()
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
    ctx = multiprocessing.get_context("spawn")
    out = ctx.Array("i", [0] * world)
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_traj, 5, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert list(out) == [1, 1]


def _rendezvous_worker(rank, world, directory, port, out):
    os.environ["MASTER_PORT"] = str(port)
    uid = exd.exchange_unique_id(rank, world, lambda: bytes(range(128)), directory=directory, timeout=60.0)
    out[rank] = 1 if uid == bytes(range(128)) else 0


def test_unique_id_rendezvous_world_size_2():
    """Both ranks are children of this process: same launch token, rank 0 publishes, rank 1 polls (started first, so it waits)."""
    ctx = multiprocessing.get_context("fork")
    out = ctx.Array("i", [0, 0])
    with tempfile.TemporaryDirectory() as d:
        port = _free_port()
        p1 = ctx.Process(target=_rendezvous_worker, args=(1, 2, d, port, out))
        p1.start()
        p0 = ctx.Process(target=_rendezvous_worker, args=(0, 2, d, port, out))
        p0.start()
        for p in (p0, p1):
            p.join(90)
            assert p.exitcode == 0
        assert len(os.listdir(d)) == 1  # the id file, no leftover temporaries
    assert list(out) == [1, 1]


def test_single_rank_needs_no_rendezvous():
    assert exd.exchange_unique_id(0, 1, lambda: b"x" * 128) is None

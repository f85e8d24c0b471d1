"""Sharding and rendezvous for the multi-GPU path.  No torch: the collectives are the C-ABI's
(exo_comm_* in include/exotica_b200.h, NCCL over NVLink / NVSwitch on the caller's stream), this module only decides which
rank evaluates what and gets the NCCL unique id from rank 0 to the other ranks.

The hot path shards with no exchange: every configuration / trajectory is independent and the robot constants are a few
KB, so each rank (one process per GPU) evaluates a contiguous block of whole trajectories.  Collectives appear only where
the path has a real exchange step: gathering the per-trajectory costs (all-gather, ceil(B/G) doubles per rank) and the
minimum sphere distance (all-gather / all-reduce(min)).
"""
import os
import tempfile
import time

import numpy as np


def shard_range(n_units, rank, world_size):
    """Contiguous block [begin, end) of the n_units trajectories (or configurations) that `rank` evaluates.
    Blocks have ceil(n / world) units except the trailing ranks, which may hold fewer (or none)."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank %r / world size %r" % (rank, world_size))
    per = -(-int(n_units) // world_size)
    begin = min(int(n_units), rank * per)
    return begin, min(int(n_units), begin + per)


def units_per_rank(n_units, world_size):
    """Slot size of the gathered arrays: every rank owns ceil(n / world) entries (trailing ranks pad)."""
    return -(-int(n_units) // int(world_size))


def shard_evaluations(n_traj, T, rank, world_size):
    """Evaluation range [e0, e1) of a rank for trajectories of T timesteps: trajectories are never split, so the
    transition term x[t] - x[t-1] stays local to a rank."""
    b0, b1 = shard_range(n_traj, rank, world_size)
    return b0 * T, b1 * T


def split_states(X, T, rank, world_size):
    """X[B][T][N] (or [B*T][N]) -> this rank's block of whole trajectories, as a view."""
    X = np.asarray(X)
    flat = X.reshape(-1, X.shape[-1])
    e0, e1 = shard_evaluations(flat.shape[0] // T, T, rank, world_size)
    return flat[e0:e1]


# ---- rendezvous: ship rank 0's unique id to the other ranks of one node ---------------------------------------------------
def _launch_token():
    """Something all ranks of one launch share and no other launch does: the launcher process (torchrun's agent, mpirun,
    a test's parent) with its start time, plus the rendezvous port."""
    ppid = os.getppid()
    start = "0"
    try:
        with open("/proc/%d/stat" % ppid) as f:
            start = f.read().rsplit(")", 1)[1].split()[19]  # field 22: start time in clock ticks since boot
    except Exception:
        pass
    return "%d_%s_%s_%s" % (ppid, start, os.environ.get("MASTER_PORT", "0"), os.environ.get("TORCHELASTIC_RESTART_COUNT", "0"))


def exchange_unique_id(rank, world_size, make_id, token=None, directory=None, timeout=300.0):
    """Rank 0 calls make_id() and publishes the bytes through a file (atomic rename) that the other ranks of the same launch
    poll for; returns the id on every rank.  One node only (the ranks share a filesystem), which is what one process per GPU
    of one box needs; multi-node launchers pass the id themselves (MPI_Bcast ...) to _capi.Comm."""
    if world_size == 1:
        return None
    directory = directory or os.environ.get("EXO_RENDEZVOUS_DIR") or tempfile.gettempdir()
    path = os.path.join(directory, "exotica_b200_ncclid_%s" % (token or _launch_token()))
    if rank == 0:
        uid = make_id()
        tmp = "%s.%d.tmp" % (path, os.getpid())
        with open(tmp, "wb") as f:
            f.write(uid)
        os.replace(tmp, path)
        return uid
    deadline = time.time() + timeout
    while time.time() < deadline:
        try:
            with open(path, "rb") as f:
                uid = f.read()
            if len(uid) >= 128:
                return uid[:128]
        except FileNotFoundError:
            pass
        time.sleep(0.01)
    raise RuntimeError("rank %d: no unique id from rank 0 after %.0f s (%s)" % (rank, timeout, path))


def release_unique_id(rank, token=None, directory=None):
    """Rank 0 removes the rendezvous file once every rank has its communicator (after the first barrier)."""
    if rank != 0:
        return
    directory = directory or os.environ.get("EXO_RENDEZVOUS_DIR") or tempfile.gettempdir()
    try:
        os.remove(os.path.join(directory, "exotica_b200_ncclid_%s" % (token or _launch_token())))
    except OSError:
        pass


def bind_to_gpu_numa_node(device_index):
    """Pins the calling process to the CPUs of the NUMA node its GPU hangs off (sysfs, through the GPU's PCI bus id from NVML), so that
    the pinned host pools allocated afterwards -- and the threads that fill them -- sit on the memory controller next to that GPU's
    PCIe root port.  With eight ranks on one node every host<->device copy otherwise crosses the socket interconnect for half of the
    GPUs.  Returns {"numa_node", "cpus"} or None when the topology cannot be read (single-node hosts, containers without sysfs)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(device_index)).busId
        if isinstance(bus, bytes):
            bus = bus.decode()
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:  # NVML prints an 8-digit PCI domain, sysfs a 4-digit one
            bus = bus[4:]
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return {"numa_node": node, "cpus": len(cpus)}
    except Exception:
        return None


def init_from_env():
    """One process per GPU, launched torchrun-style (RANK, WORLD_SIZE, LOCAL_RANK in the environment; a plain `python` run is
    rank 0 of 1): selects the rank's device, builds the communicator, returns (comm, rank, world, local_rank)."""
    from . import _capi
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", str(rank)))
    _capi.check(_capi.lib().exo_set_device(local_rank))
    uid = exchange_unique_id(rank, world, _capi.Comm.unique_id)
    comm = _capi.Comm(rank, world, uid)
    comm.barrier()
    release_unique_id(rank)
    return comm, rank, world, local_rank

#!/usr/bin/env python3
"""Multi-rank check of the C-ABI collectives (run under torchrun, one rank per GPU): all-gather / all-reduce results against what every
rank can compute from the seeds, for message sizes on both sides of the peer-window limit, back-to-back (the two-parity protocol) and
replayed from a CUDA graph; then the latency of a 512-double all-gather through the peer windows and through NCCL (EXO_COMM_PEER=0
in a second run).  Rank 0 prints one JSON line."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from exotica_b200 import _capi as capi  # noqa: E402
from exotica_b200 import distributed as exd  # noqa: E402


def data(rank, n, it):
    return np.random.default_rng(1000 * it + rank).normal(size=n)


def main():
    comm, rank, world, local_rank = exd.init_from_env()
    L = capi.lib()
    stream = capi.Stream()
    worst = 0.0
    for n in (1, 7, 512, 8192, 8193, 20000):
        d_send, d_recv, d_red = capi.DeviceBuffer(n), capi.DeviceBuffer(n * world), capi.DeviceBuffer(n)
        for it in range(25):
            d_send.upload(data(rank, n, it))
            comm.allgather(d_send.ptr, d_recv.ptr, n, stream=stream)
            comm.allreduce(d_send.ptr, d_red.ptr, n, capi.REDUCE_MIN if it % 3 == 0 else (capi.REDUCE_SUM if it % 3 == 1 else capi.REDUCE_MAX), stream=stream)
            stream.synchronize()
            ref = np.stack([data(r, n, it) for r in range(world)])
            worst = max(worst, float(np.max(np.abs(d_recv.download((world, n)) - ref))))
            got = d_red.download((n,))
            if it % 3 == 0:
                worst = max(worst, float(np.max(np.abs(got - ref.min(axis=0)))))
            elif it % 3 == 2:
                worst = max(worst, float(np.max(np.abs(got - ref.max(axis=0)))))
            else:
                acc = ref[0].copy()
                for r in range(1, world):
                    acc = acc + ref[r]
                worst = max(worst, float(np.max(np.abs(got - acc)) / max(1.0, float(np.max(np.abs(acc))))))
    # in place, many back to back without host synchronisation, then replayed from a graph
    n = 512
    d_all = capi.DeviceBuffer(n * world)
    mine = data(rank, n, 99)
    L.exo_memcpy_h2d(d_all.ptr + rank * n * 8, mine.ctypes.data, n * 8, None)
    L.exo_stream_synchronize(None)
    for _ in range(300):
        comm.allgather(d_all.ptr + rank * n * 8, d_all.ptr, n, stream=stream)
    stream.synchronize()
    worst = max(worst, float(np.max(np.abs(d_all.download((world, n)) - np.stack([data(r, n, 99) for r in range(world)])))))
    with capi.Graph.capture(stream) as graph:
        for _ in range(10):
            comm.allgather(d_all.ptr + rank * n * 8, d_all.ptr, n, stream=stream)
    comm.barrier(stream)
    ev0, ev1 = capi.Event(), capi.Event()
    for _ in range(5):
        graph.launch(stream)
    comm.barrier(stream)
    ev0.record(stream)
    reps = 50
    for _ in range(reps):
        graph.launch(stream)
    ev1.record(stream)
    comm.barrier(stream)
    us = ev0.elapsed_ms(ev1) * 1e3 / (reps * 10)
    worst = max(worst, float(np.max(np.abs(d_all.download((world, n)) - np.stack([data(r, n, 99) for r in range(world)])))))
    if rank == 0:
        print(json.dumps({"check": "collectives", "world": world, "max_abs_err": worst, "allgather_512_doubles_us": us,
                          "peer_collectives": comm.peer_collective_count(), "nccl_collectives": comm.collective_count(),
                          "path": "peer windows" if comm.peer_collective_count() > 0 else "NCCL"}), flush=True)
    graph.__del__()  # graphs that captured NCCL kernels go before their communicator
    comm.barrier(stream)
    comm.destroy()
    if worst > 1e-12:
        sys.exit(1)


if __name__ == "__main__":
    main()

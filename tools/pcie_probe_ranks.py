#!/usr/bin/env python3
"""What can N ranks on one node move over PCIe at the same time?  Every rank copies what one e2e step of bench.py moves (402.7 MB
device -> pinned host over four streams + 58.7 MB pinned host -> device), all ranks between the same two barriers; rank 0 prints the
per-rank and aggregate rates with and without binding each rank to its GPU's NUMA node.  No torch: exotica_b200's C-ABI only.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 tools/pcie_probe_ranks.py
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import exotica_b200 as exo  # noqa: E402
from exotica_b200 import _capi as capi  # noqa: E402
from exotica_b200 import distributed as exd  # noqa: E402


def main():
    comm, rank, world, local_rank = exd.init_from_env()
    L = capi.lib()
    n_out, n_in, K = 402653184 // 8, 58720256 // 8, 4
    d_out, d_in = capi.DeviceBuffer(n_out), capi.DeviceBuffer(n_in)
    streams = [capi.Stream() for _ in range(K + 1)]
    results = {}
    for mode in ("unbound", "numa"):
        binding = exd.bind_to_gpu_numa_node(local_rank) if mode == "numa" else None
        h_out, h_in = exo.PinnedArray((n_out,)), exo.PinnedArray((n_in,))  # allocated (first touched) after the binding
        h_out.array[:] = 0.0
        h_in.array[:] = 1.0

        def step():
            capi.check(L.exo_memcpy_h2d(d_in.ptr, h_in.array.ctypes.data, n_in * 8, streams[K].ptr))
            per = n_out // K
            for i in range(K):
                capi.check(L.exo_memcpy_d2h(h_out.array.ctypes.data + i * per * 8, d_out.ptr + i * per * 8, per * 8, streams[i].ptr))

        def sync():
            for s in streams:
                s.synchronize()

        for _ in range(2):
            step()
        sync()
        comm.barrier()
        t0 = time.perf_counter()
        reps = 8
        for _ in range(reps):
            step()
        sync()
        comm.barrier()
        dt = (time.perf_counter() - t0) / reps
        mine = np.array([(n_out + n_in) * 8 / dt / 1e9])
        scalar = capi.DeviceBuffer(1)
        scalar.upload(mine)
        comm.allreduce(scalar.ptr, scalar.ptr, 1, capi.REDUCE_SUM)
        capi.check(L.exo_device_synchronize())
        total = float(scalar.download((1,))[0])
        scalar.upload(mine)
        comm.allreduce(scalar.ptr, scalar.ptr, 1, capi.REDUCE_MIN)
        capi.check(L.exo_device_synchronize())
        worst = float(scalar.download((1,))[0])
        results[mode] = {"ms_per_step": dt * 1e3, "rank0_GBps": float(mine[0]), "sum_of_rank_rates_GBps": total, "slowest_rank_GBps": worst,
                         "aggregate_GBps_by_wall_clock": world * (n_out + n_in) * 8 / dt / 1e9, "binding_rank0": binding}
    comm.barrier()
    comm.destroy()
    if rank == 0:
        print(json.dumps({"probe": "pcie, %d rank(s) at once, per step 402.7 MB D2H over 4 streams + 58.7 MB H2D" % world, "world": world, **results}), flush=True)


if __name__ == "__main__":
    main()

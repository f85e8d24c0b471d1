#!/usr/bin/env python3
"""Multi-GPU runs of the sharded configurations with their collectives; one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/bench_multi.py

cfg3: TimeIndexedProblem T=100, EffFrame + JointLimit, 4096 trajectories sharded by trajectory; each rank evaluates Phi and
      J of its block, reduces them to per-trajectory costs and all ranks all-gather the 4096 costs over NCCL.
cfg5: SphereCollision (64 spheres, 1024 pairs), 4096 x T=100; per-trajectory min distance, all-reduce(min) over NCCL.
Strong scaling (total work fixed).  Timed on the device (CUDA events on the stream that carries kernels and collective),
max over ranks; rank 0 prints one JSON line per configuration and checks the gathered costs against the CPU oracle.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import exotica_b200 as exo  # noqa: E402
from exotica_b200 import distributed as exd  # noqa: E402
import bench_configs  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--traj", type=int, default=4096)
    args = ap.parse_args()
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    exo._capi.check(exo._capi.lib().exo_set_device(local))
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group(backend="nccl", device_id=dev)
    stream = torch.cuda.current_stream()
    sptr = stream.cuda_stream
    for name, make in (("cfg3", bench_configs.cfg3), ("cfg5", bench_configs.cfg5)):
        init, _, bytes_per_eval = make()
        prob = exo.Setup.create_problem(init)
        N, LP, LJ, T = prob.N, prob.length_Phi, prob.length_jacobian, prob.T
        B = args.traj
        b0, b1 = exd.shard_range(B, rank, world)
        nb = b1 - b0
        lim = prob.get_scene().get_kinematic_tree().get_joint_limits()
        rng = np.random.default_rng(1234)
        q0 = rng.uniform(lim[:, 0], lim[:, 1], size=(B, 1, N))
        X = np.clip(q0 + np.cumsum(rng.normal(0, 0.05, size=(B, T, N)), axis=1), lim[:, 0] - 0.1, lim[:, 1] + 0.1)
        mine = np.ascontiguousarray(X[b0:b1].reshape(nb * T, N).T)                      # SoA block of whole trajectories
        d_x = torch.from_numpy(mine).to(dev)
        d_phi = torch.empty((LP, max(nb * T, 1)), dtype=torch.float64, device=dev)
        d_jac = torch.empty((LJ, N, max(nb * T, 1)), dtype=torch.float64, device=dev)
        d_cost = torch.zeros(max(nb, 1), dtype=torch.float64, device=dev)
        d_min_eval = torch.empty(max(nb * T, 1), dtype=torch.float64, device=dev)
        d_min = torch.empty(max(nb, 1), dtype=torch.float64, device=dev)
        backend = exo._host.lib().exo_host_problem_backend(prob._h)
        if name == "cfg5":
            exo._capi.check(exo._capi.lib().exo_problem_set_min_distance_output(backend, d_min_eval.data_ptr()))

        def step():
            if nb > 0:
                prob.update_batch_device(d_x.data_ptr(), nb * T, d_phi.data_ptr(), d_jac.data_ptr(), layout=exo.LAYOUT_SOA, stream=sptr)
                prob.cost_batch_device(d_x.data_ptr(), d_phi.data_ptr(), nb, d_cost.data_ptr(), layout=exo.LAYOUT_SOA, stream=sptr)
            costs = exd.gather_per_trajectory(d_cost[:nb], B) if world > 1 else d_cost[:nb]
            gmin = None
            if name == "cfg5":
                if nb > 0:
                    exo._capi.check(exo._capi.lib().exo_trajectory_reduce(d_min_eval.data_ptr(), nb, T, 0, 0, d_min.data_ptr(), sptr))
                local_min = d_min[:nb].min().reshape(1) if nb > 0 else torch.full((1,), float("inf"), dtype=torch.float64, device=dev)
                gmin = exd.allreduce_min(local_min) if world > 1 else local_min
            return costs, gmin

        for _ in range(3):
            costs, gmin = step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        for _ in range(args.reps):
            costs, gmin = step()
        ev1.record(stream)
        torch.cuda.synchronize()
        t = torch.tensor([ev0.elapsed_time(ev1) / args.reps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        if rank == 0:
            # parity of the gathered costs on a sample of trajectories (CPU oracle)
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            err = None
            if name == "cfg3":
                from oracle import oracle
                robots = os.path.join(ROOT, "tests", "fixtures", "robots")
                orc = oracle.Oracle(oracle.load_robot(os.path.join(robots, "lwr_simplified.urdf"), os.path.join(robots, "lwr_simplified.srdf"), "arm"))
                orc.set_problem([{"type": "EffFrame", "EndEffector": [{"Link": "lwr_arm_7_link"}]}, {"type": "JointLimit"}])
                sel = list(range(0, B, max(1, B // 16)))
                ref = np.array([orc.trajectory_cost(X[b], [1.0, 1.0], ct=1.0 / 0.05 / T, W=np.ones(N))[0] for b in sel])
                got = costs.cpu().numpy()[sel]
                err = float(np.max(np.abs(got - ref) / np.maximum(1.0, np.abs(ref))))
            print(json.dumps({"config": name, "n_gpus": world, "trajectories": B, "T": T, "ms_per_step": ms, "evals_per_s": B * T / (ms * 1e-3),
                              "algorithmic_GBps_per_gpu": bytes_per_eval * B * T / world / (ms * 1e-3) / 1e9, "scaling": "strong",
                              "collective": "all_gather(%d costs)%s" % (B, " + all_reduce(min distance)" if name == "cfg5" else ""),
                              "gathered_cost_rel_err_vs_oracle": err, "min_distance": (float(gmin.item()) if gmin is not None else None)}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""cfg3's solver-facing kernels with one goal for all timesteps (T_goal = 1, what the host mirror pushes when the goals are equal)
against a goal per timestep (T_goal = T, what a binding that forwards TimeIndexedTask::y pushes): device times."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from exotica_b200 import _capi  # noqa: E402
from helpers import gpu_problem_from_maps  # noqa: E402
from oracle import oracle  # noqa: E402

robots = os.path.join(ROOT, "tests", "fixtures", "robots")
orc = oracle.Oracle(oracle.load_robot(os.path.join(robots, "lwr_simplified.urdf"), os.path.join(robots, "lwr_simplified.srdf"), "arm"))
maps = [{"type": "EffFrame", "EndEffector": [{"Link": "lwr_arm_7_link"}]}, {"type": "JointLimit"}]
orc.set_problem(maps)
B, T, N = 4096, 100, 7
E = B * T
rng = np.random.default_rng(3)
X = rng.uniform(orc.lower, orc.upper, size=(E, N))
goal1 = orc.zero_goal()
goal1[:6] = [0.4, 0.1, 0.6, 0.2, -0.1, 0.3]
stream = _capi.Stream()


def timed(fn, reps=30):
    for _ in range(3):
        fn()
    ev0, ev1 = _capi.Event(), _capi.Event()
    ev0.record(stream)
    for _ in range(reps):
        fn()
    ev1.record(stream)
    ev1.synchronize()
    return ev0.elapsed_ms(ev1) / reps


for name, goal in (("T_goal = 1", goal1), ("T_goal = T", np.tile(goal1, (T, 1)))):
    tree, prob = gpu_problem_from_maps(orc, maps)
    LP, LJ = prob.length_phi, prob.length_jacobian
    dx, dphi, djac, dg, dc = (_capi.DeviceBuffer(n) for n in (E * N, E * LP, E * LJ * N, E * N, B))
    dx.upload(X.T)
    prob.set_cost(np.ones((T, 2)), T=T, goal=goal, ct=0.2, W=np.ones(N))
    prob.update_batch_device(dx.ptr, E, dphi.ptr, djac.ptr, T=T, stream=stream.ptr)
    out = {"goals": name}
    out["cost_kernel_us"] = 1e3 * timed(lambda: prob.cost_batch_device(dx.ptr, dphi.ptr, B, dc.ptr, None, stream=stream.ptr))
    out["cost_gradient_kernel_us"] = 1e3 * timed(lambda: prob.cost_jacobian_batch_device(dx.ptr, dphi.ptr, djac.ptr, B, dg.ptr, stream=stream.ptr))
    out["fused_costgrad_us"] = 1e3 * timed(lambda: prob.cost_gradient_batch_device(dx.ptr, B, dg.ptr, dc.ptr, stream=stream.ptr))
    out["update_cost_us"] = 1e3 * timed(lambda: prob.update_cost_batch_device(dx.ptr, B, dphi.ptr, djac.ptr, dc.ptr, stream=stream.ptr))
    out["cost0"] = float(dc.download((B,))[0])
    print(json.dumps(out), flush=True)

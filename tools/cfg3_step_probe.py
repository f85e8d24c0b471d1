#!/usr/bin/env python3
"""The kernels of one rank's share of the cfg3 step at eight GPUs (512 of the 4096 trajectories, T = 100), back to back on one stream:
update (Phi, J) + cost kernel.  Run under `ncu --metrics gpu__time_duration.sum` for the per-kernel durations DESIGN.md section 6 quotes
(the collective is measured by bench.py itself at N > 1)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import exotica_b200 as exo  # noqa: E402
from exotica_b200 import _capi  # noqa: E402
from bench_configs import cfg3  # noqa: E402

init, _, _ = cfg3()
prob = exo.Setup.create_problem(init)
N, LP, LJ, T = prob.N, prob.length_Phi, prob.length_jacobian, prob.T
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
E = B * T
lim = prob.get_scene().get_kinematic_tree().get_joint_limits()
X = np.random.default_rng(3).uniform(lim[:, 0], lim[:, 1], size=(B, T, N))
prob.cost_batch(X[:2])
dx, dphi, djac, dc = (_capi.DeviceBuffer(n) for n in (E * N, E * LP, E * LJ * N, B))
dx.upload(X.reshape(E, N).T)
stream = _capi.Stream()
for _ in range(6):
    prob.update_batch_device(dx.ptr, E, dphi.ptr, djac.ptr, layout=exo.LAYOUT_SOA, stream=stream.ptr)
    prob.cost_batch_device(dx.ptr, dphi.ptr, B, dc.ptr, layout=exo.LAYOUT_SOA, stream=stream.ptr)
stream.synchronize()
ev0, ev1 = _capi.Event(), _capi.Event()
ev0.record(stream)
for _ in range(50):
    prob.update_batch_device(dx.ptr, E, dphi.ptr, djac.ptr, layout=exo.LAYOUT_SOA, stream=stream.ptr)
    prob.cost_batch_device(dx.ptr, dphi.ptr, B, dc.ptr, layout=exo.LAYOUT_SOA, stream=stream.ptr)
ev1.record(stream)
stream.synchronize()
print("B = %d trajectories x T = %d: update + cost kernel back to back %.2f us per step (no graph, no collective)" % (B, T, ev0.elapsed_ms(ev1) * 1e3 / 50))

"""Whole-body multi-start IK only (38 DoF, four end-effector frames = 24 task rows, 65 536 perturbed starts), for a kernel launch
list under ncu:  ncu --metrics gpu__time_duration.sum --csv python tools/wbik_probe.py"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import exotica_b200 as exo  # noqa: E402
from bench_configs import frame, scene  # noqa: E402

maps = [("exotica/EffFrame", {"Name": "Eff", "EndEffector": [frame("leftFoot"), frame("rightFoot"), frame("leftPalm"), frame("rightPalm")]})]
init = ("exotica/UnconstrainedEndPoseProblem", {"Name": "wb", "PlanningScene": scene("valkyrie_sim", "whole_body"), "Maps": maps,
                                                "Cost": [("exotica/Task", {"Task": "Eff", "Rho": "1"})]})
wb = exo.Setup.create_problem(init)
lim = wb.get_scene().get_kinematic_tree().get_joint_limits()
lo = np.where(lim[:, 0] < -1e19, -1.0, lim[:, 0])
hi = np.where(lim[:, 1] > 1e19, 1.0, lim[:, 1])
q_goal = np.random.default_rng(10).uniform(np.maximum(lo, -0.5), np.minimum(hi, 0.5)) * 0.5
wb.update(q_goal)
goal = " ".join("%.17g" % v for v in np.array(wb.Phi)[:24])
ik_init = ("exotica/UnconstrainedEndPoseProblem", {"Name": "wbik", "PlanningScene": scene("valkyrie_sim", "whole_body"), "Maps": maps,
                                                   "Cost": [("exotica/Task", {"Task": "Eff", "Rho": "1", "Goal": goal})]})
solver = exo.IKSolver({"MaxIterations": 20})
solver.specify_problem(exo.Setup.create_problem(ik_init))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 16
Q0 = q_goal + np.random.default_rng(11).normal(0, 0.05, size=(n, wb.N))
solver.solve_batch(Q0)
for rep in range(int(os.environ.get("WBIK_REPS", "1"))):
    t0 = time.perf_counter()
    Q, cost, iters, term = solver.solve_batch(Q0)
    print("whole-body IK %.2f ms for %d starts, mean iterations %.2f, max %d" % ((time.perf_counter() - t0) * 1e3, n, iters.mean(), iters.max()))

"""Multi-start IK only (65 536 LWR starts), for kernel-level timing under ncu."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import exotica_b200 as exo  # noqa: E402
from bench_configs import frame, scene  # noqa: E402

maps = [("exotica/EffFrame", {"Name": "Eff", "EndEffector": [frame("lwr_arm_6_link")]})]
init = ("exotica/UnconstrainedEndPoseProblem", {"Name": "ik", "PlanningScene": scene("lwr_simplified", "arm"), "Maps": maps,
                                                "Cost": [("exotica/Task", {"Task": "Eff", "Goal": "0.4 0.1 0.6 0 0 0"})]})
prob = exo.Setup.create_problem(init)
lim = prob.get_scene().get_kinematic_tree().get_joint_limits()
solver = exo.IKSolver({"MaxIterations": int(sys.argv[1]) if len(sys.argv) > 1 else 100})
solver.specify_problem(prob)
Q0 = np.random.default_rng(7).uniform(lim[:, 0], lim[:, 1], size=(1 << 16, prob.N))
solver.solve_batch(Q0)
for rep in range(int(os.environ.get("IK_PROBE_REPS", "1"))):
    t0 = time.perf_counter()
    Q, cost, it, term = solver.solve_batch(Q0)
    print("IK %.2f ms, mean iterations %.1f, kernel %s" % ((time.perf_counter() - t0) * 1e3, it.mean(), prob.last_kernel()))

#!/usr/bin/env python3
"""Times the kernels of SURVEY.md section 8 (f) -- task cost, cost gradient, task-map Hessians, batched IK -- and the two
host-buffer paths of a time-indexed problem (Phi + J back over PCIe vs. cost + gradient back) on BASELINE.json's cfg3
(LWR, EffFrame + JointLimit, T = 100, 4096 trajectories).

    python tools/bench_next.py [--reps 10]

Device kernels are timed as the mean of `reps` back-to-back launches between two device synchronisations (wall clock; every
launch is > 50 us of work); host paths by wall clock around the blocking call on pinned-free numpy buffers.  One JSON line each.
Parity for all of them is covered by tests/; this script only measures.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import exotica_b200 as exo  # noqa: E402
from exotica_b200 import _capi, _host  # noqa: E402
from bench_configs import cfg3, frame, scene  # noqa: E402


def sync():
    _capi.check(_capi.lib().exo_device_synchronize())


WARMUPS = 3


def timed(fn, reps):
    for _ in range(WARMUPS):
        fn()
    sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    sync()
    return (time.perf_counter() - t0) / reps * 1e3


def wb_goal(problem, q):
    """Phi of the whole-body problem's EffFrame map at state q (the IK goal)."""
    problem.update(q)
    return np.array(problem.Phi)[:24]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--profile", action="store_true", help="one warm-up and one launch per kernel, no host paths, 4 IK iterations (for ncu)")
    args = ap.parse_args()
    if args.profile:
        global WARMUPS
        WARMUPS, args.reps = 1, 1
    peak = 6549.4
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak = float(json.load(f)["hbm_gbs"])
    except Exception:
        pass

    init, E, _ = cfg3()
    prob = exo.Setup.create_problem(init)
    N, LP, LJ, T = prob.N, prob.length_Phi, prob.length_jacobian, prob.T
    B = E // T
    lim = prob.get_scene().get_kinematic_tree().get_joint_limits()
    rng = np.random.default_rng(7)
    q0 = rng.uniform(lim[:, 0], lim[:, 1], size=(B, 1, N))
    X = np.clip(q0 + np.cumsum(rng.normal(0, 0.02, size=(B, T, N)), axis=1), lim[:, 0], lim[:, 1])
    backend = _capi.Problem.__new__(_capi.Problem)  # a view of the host problem's backend handle: borrowed, released below
    backend.h = _host.lib().exo_host_problem_backend(prob._h)
    dx, dphi, djac, dg, dc = (_capi.DeviceBuffer(n) for n in (E * N, E * LP, E * LJ * N, E * N, B))
    dx.upload(X.reshape(E, N).T)
    prob.cost_batch(X[:2])  # pushes rho / goal / W of the XML-style initializer to the device

    def line(name, ms, units, unit, bytes_per_unit=None, **kw):
        d = {"bench": name, "ms": ms, "units": units, unit + "_per_s": units / (ms * 1e-3)}
        if bytes_per_unit:
            d["algorithmic_GBps"] = bytes_per_unit * units / (ms * 1e-3) / 1e9
            d["frac_of_measured_hbm"] = d["algorithmic_GBps"] / peak
        d.update(kw)
        print(json.dumps(d), flush=True)

    ms = timed(lambda: prob.update_batch_device(dx.ptr, E, dphi.ptr, djac.ptr, layout=exo.LAYOUT_SOA), args.reps)
    line("update_batch (FK + J), device", ms, E, "evals", 8 * (N + LP + LJ * N))
    ms = timed(lambda: backend.cost_batch_device(dx.ptr, dphi.ptr, B, dc.ptr, None), args.reps)
    line("cost kernel (reads x, Phi), device", ms, E, "evals", 8 * (N + LP))
    ms = timed(lambda: backend.update_cost_batch_device(dx.ptr, B, dphi.ptr, djac.ptr, dc.ptr, None), args.reps)
    line("update + cost in one call (%s + per-trajectory sum), device" % backend.last_kernel(), ms, E, "evals", 8 * (N + LP + LJ * N + 1))
    ms = timed(lambda: backend.cost_jacobian_batch_device(dx.ptr, dphi.ptr, djac.ptr, B, dg.ptr), args.reps)
    line("cost-gradient kernel (reads x, Phi, J; writes N per eval), device", ms, E, "evals", 8 * (2 * N + LP + LJ * N))

    ms = timed(lambda: backend.cost_gradient_batch_device(dx.ptr, B, dg.ptr, dc.ptr), args.reps)
    line("fused FK + cost + gradient (exo_chain_costgrad_kernel + per-trajectory sum; J never materialised), device", ms, E, "evals", 8 * (2 * N + 1),
         kernel=backend.last_kernel())

    dR, dr, drh, dcc = (_capi.DeviceBuffer(n) for n in (E * N * N, E * N, E, E))
    ms = timed(lambda: backend.aico_task_messages_device(dx.ptr, dphi.ptr, djac.ptr, B, dR.ptr, dr.ptr, drh.ptr, dcc.ptr), args.reps)
    line("AICO task messages R, r, rhat, cost (reads x, Phi, J; writes N^2 + N + 2 per eval), device", ms, E, "evals", 8 * (N + LP + LJ * N + N * N + N + 2))
    for b in (dR, dr, drh, dcc):
        b.free()

    # Hessians: length_jacobian x N x N doubles per evaluation
    EH = 1 << 16
    dh = _capi.DeviceBuffer(EH * LJ * N * N)
    ms = timed(lambda: backend.hessian_batch_device(dx.ptr, djac.ptr, EH, dh.ptr, T=1), max(3, args.reps // 2))
    line("task-map Hessians (raw-frame update + Hessian kernel), device", ms, EH, "evals", 8 * (LJ * N * N + LJ * N + N))
    dh.free()

    if args.profile:
        maps = [("exotica/EffFrame", {"Name": "Eff", "EndEffector": [frame("lwr_arm_6_link")]})]
        ik_init = ("exotica/UnconstrainedEndPoseProblem", {"Name": "ik", "PlanningScene": scene("lwr_simplified", "arm"), "Maps": maps,
                                                           "Cost": [("exotica/Task", {"Task": "Eff", "Goal": "0.4 0.1 0.6 0 0 0"})]})
        solver = exo.IKSolver({"MaxIterations": 4})
        solver.specify_problem(exo.Setup.create_problem(ik_init))
        solver.solve_batch(rng.uniform(lim[:, 0], lim[:, 1], size=(1 << 16, N)))
        backend.h = None
        return

    # host-buffer paths of the same problem
    px, pphi, pjac = exo.PinnedArray((B, T, N)), exo.PinnedArray((E, LP)), exo.PinnedArray((E, LJ, N))
    px.array[:] = X
    t0 = time.perf_counter()
    prob.update_batch(px.array, out_phi=pphi.array, out_jacobian=pjac.array)
    t1 = time.perf_counter()
    prob.update_batch(px.array, out_phi=pphi.array, out_jacobian=pjac.array)
    t2 = time.perf_counter()
    line("host path (pinned buffers): X in, Phi + J out", (t2 - t1) * 1e3, E, "evals", None, d2h_bytes=8 * E * (LP + LJ * N), first_call_ms=(t1 - t0) * 1e3)
    X = px.array
    prob.cost_jacobian_batch(X)
    t1 = time.perf_counter()
    prob.cost_jacobian_batch(X)
    t2 = time.perf_counter()
    line("host path: X in (pinned), cost + gradient out into pageable arrays (Phi, J stay in HBM)", (t2 - t1) * 1e3, E, "evals", None, d2h_bytes=8 * (E * N + B))
    pcost, pgrad = exo.PinnedArray((B,)), exo.PinnedArray((B, T, N))
    prob.cost_jacobian_batch(X, out_cost=pcost.array, out_jacobian=pgrad.array)
    best = 1e9
    for _ in range(5):
        t1 = time.perf_counter()
        prob.cost_jacobian_batch(X, out_cost=pcost.array, out_jacobian=pgrad.array)
        best = min(best, time.perf_counter() - t1)
    c_ref, g_ref = prob.cost_jacobian_batch(X)
    assert np.array_equal(c_ref, pcost.array) and np.array_equal(g_ref, pgrad.array)
    line("host path: X in, cost + gradient out, all buffers pinned (best of 5)", best * 1e3, E, "evals", None, d2h_bytes=8 * (E * N + B),
         h2d_bytes=8 * E * N, link_GBps=8 * (2 * E * N + B) / best / 1e9)
    prob.cost_batch(X)
    t1 = time.perf_counter()
    prob.cost_batch(X)
    t2 = time.perf_counter()
    line("host path: X in, cost out (FK only)", (t2 - t1) * 1e3, E, "evals", None, d2h_bytes=8 * B)

    backend.h = None  # borrowed handle: the host problem owns (and destroys) it
    for b in (dx, dphi, djac, dg, dc):
        b.free()

    # whole-body problem (38-DoF humanoid, four end-effector frames + joint limits): cost and gradient of 2^18 configurations through
    # the generated fused kernel of tree programs and through the Phi / J round trip (update + cost + cost-gradient kernels)
    from bench_configs import frame, scene
    maps = [("exotica/EffFrame", {"Name": "Eff", "EndEffector": [frame("leftFoot"), frame("rightFoot"), frame("leftPalm"), frame("rightPalm")]}),
            ("exotica/JointLimit", {"Name": "JL"})]
    init = ("exotica/UnconstrainedEndPoseProblem", {"Name": "wb", "PlanningScene": scene("valkyrie_sim", "whole_body"), "Maps": maps,
                                                    "Cost": [("exotica/Task", {"Task": "Eff", "Rho": "10"}), ("exotica/Task", {"Task": "JL", "Rho": "100"})]})
    wb = exo.Setup.create_problem(init)
    E2, N2 = 1 << 18, wb.N
    lim2 = wb.get_scene().get_kinematic_tree().get_joint_limits()
    lo = np.where(lim2[:, 0] < -1e19, -1.0, lim2[:, 0])
    hi = np.where(lim2[:, 1] > 1e19, 1.0, lim2[:, 1])
    X2 = np.random.default_rng(9).uniform(lo, hi, size=(E2, N2))
    wb.cost_batch(X2[:2])
    b2 = _capi.Problem.__new__(_capi.Problem)
    b2.h = _host.lib().exo_host_problem_backend(wb._h)
    dx2, dg2, dc2 = _capi.DeviceBuffer(E2 * N2), _capi.DeviceBuffer(E2 * N2), _capi.DeviceBuffer(E2)
    dx2.upload(X2.T)
    ms = timed(lambda: b2.cost_gradient_batch_device(dx2.ptr, E2, dg2.ptr, dc2.ptr), args.reps)
    g_fused = dg2.download((N2, E2))
    line("whole-body cost + gradient, %s (J never materialised), device" % b2.last_kernel(), ms, E2, "evals", 8 * (2 * N2 + 1))
    b2.set_kernel(1)
    ms = timed(lambda: b2.cost_gradient_batch_device(dx2.ptr, E2, dg2.ptr, dc2.ptr), args.reps)
    line("whole-body cost + gradient, Phi / J round trip (update + cost + %s), device" % b2.last_kernel(), ms, E2, "evals", None,
         max_abs_diff_vs_fused=float(np.max(np.abs(dg2.download((N2, E2)) - g_fused))))
    # whole-body multi-start IK (BASELINE config 4's shape: 38 DoF, four end-effector frames = 24 task rows): 65 536 start states
    q_goal = np.random.default_rng(10).uniform(np.maximum(lo, -0.5), np.minimum(hi, 0.5)) * 0.5
    goal = " ".join("%.17g" % v for v in wb_goal(wb, q_goal))
    ik_init = ("exotica/UnconstrainedEndPoseProblem", {"Name": "wbik", "PlanningScene": scene("valkyrie_sim", "whole_body"), "Maps": maps[:1],
                                                       "Cost": [("exotica/Task", {"Task": "Eff", "Rho": "1", "Goal": goal})]})
    solver = exo.IKSolver({"MaxIterations": 20})
    solver.specify_problem(exo.Setup.create_problem(ik_init))
    Q0 = q_goal + np.random.default_rng(11).normal(0, 0.05, size=(1 << 16, N2))
    solver.solve_batch(Q0)  # sizes the workspace
    times = []
    for _ in range(3):
        t1 = time.perf_counter()
        Q, cost, iters, term = solver.solve_batch(Q0)
        times.append((time.perf_counter() - t1) * 1e3)
    line("whole-body batched IK (host call, 65536 starts, <= 20 iterations, 24 task rows x 38 joints; median of 3)", sorted(times)[1], Q0.shape[0],
         "solves", None, mean_iterations=float(iters.mean()), cost_below_1e_4=float((cost < 1e-4).mean()), ms_all=[round(t, 2) for t in times])
    b2.h = None
    for b in (dx2, dg2, dc2):
        b.free()

    # multi-start IK on example_ik.xml's problem shape
    maps = [("exotica/EffFrame", {"Name": "Eff", "EndEffector": [frame("lwr_arm_6_link")]})]
    ik_init = ("exotica/UnconstrainedEndPoseProblem", {"Name": "ik", "PlanningScene": scene("lwr_simplified", "arm"), "Maps": maps,
                                                       "Cost": [("exotica/Task", {"Task": "Eff", "Goal": "0.4 0.1 0.6 0 0 0"})]})
    ikp = exo.Setup.create_problem(ik_init)
    solver = exo.IKSolver({"MaxIterations": 100})
    solver.specify_problem(ikp)
    Q0 = rng.uniform(lim[:, 0], lim[:, 1], size=(1 << 16, N))
    solver.solve_batch(Q0)
    times = []
    for _ in range(5):
        t1 = time.perf_counter()
        Q, cost, it, term = solver.solve_batch(Q0)
        times.append((time.perf_counter() - t1) * 1e3)
    line("batched IK (host call, 65536 starts, <= 100 iterations; median of 5)", sorted(times)[2], Q0.shape[0], "solves", None,
         mean_iterations=float(it.mean()), cost_below_1e_4=float((cost < 1e-4).mean()), ms_all=[round(t, 2) for t in times])


if __name__ == "__main__":
    main()

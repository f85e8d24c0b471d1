#!/usr/bin/env python3
"""Times every BASELINE.json configuration (device-resident SoA buffers, CUDA events around each launch).

    python tools/bench_configs.py [--reps 20] [--only cfg2a,cfg4]

Prints one JSON line per configuration: evals/s, ms per launch, algorithmic GB/s and fraction of the measured HBM peak.
Parity for these configurations is covered by tests/test_gpu_parity.py; this script only measures.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import exotica_b200 as exo  # noqa: E402
from exotica_b200 import _capi, _host  # noqa: E402

ROBOTS = os.path.join(ROOT, "tests", "fixtures", "robots")


def scene(robot, group):
    return ("exotica/Scene", {"Name": "Scene", "JointGroup": group, "URDF": os.path.join(ROBOTS, robot + ".urdf"),
                              "SRDF": os.path.join(ROBOTS, robot + ".srdf")})


def frame(link, **kw):
    d = {"Link": link}
    d.update(kw)
    return ("exotica/Frame", d)


def cfg2a():
    maps = [("exotica/EffFrame", {"Name": "Eff", "EndEffector": [frame("lwr_arm_7_link")]})]
    return ("exotica/UnconstrainedEndPoseProblem", {"Name": "cfg2a", "PlanningScene": scene("lwr_simplified", "arm"), "Maps": maps,
                                                    "Cost": [("exotica/Task", {"Task": "Eff"})]}), 1 << 20, 440


def cfg2b():
    maps = [("exotica/EffPosition", {"Name": "Eff", "EndEffector": [frame("lwr_arm_7_link")]})]
    return ("exotica/UnconstrainedEndPoseProblem", {"Name": "cfg2b", "PlanningScene": scene("lwr_simplified", "arm"), "Maps": maps,
                                                    "Cost": [("exotica/Task", {"Task": "Eff"})]}), 1 << 20, 248


def cfg3():
    maps = [("exotica/EffFrame", {"Name": "Eff", "EndEffector": [frame("lwr_arm_7_link")]}), ("exotica/JointLimit", {"Name": "JL"})]
    return ("exotica/TimeIndexedProblem", {"Name": "cfg3", "PlanningScene": scene("lwr_simplified", "arm"), "Maps": maps, "T": 100, "tau": 0.05,
                                           "Cost": [("exotica/Task", {"Task": "Eff"}), ("exotica/Task", {"Task": "JL"})]}), 4096 * 100, 888


def cfg4():
    maps = [("exotica/EffFrame", {"Name": "Eff", "EndEffector": [frame("leftFoot"), frame("rightFoot"), frame("leftPalm"), frame("rightPalm")]})]
    return ("exotica/UnconstrainedEndPoseProblem", {"Name": "cfg4", "PlanningScene": scene("valkyrie_sim", "whole_body"), "Maps": maps,
                                                    "Cost": [("exotica/Task", {"Task": "Eff"})]}), 1 << 18, 7792


def cfg5():
    # 32 robot spheres on links 1-7 (offsets along the link), 32 obstacle spheres fixed in the world: 1024 pairs
    eff = []
    rng = np.random.default_rng(5)
    for i in range(32):
        eff.append(frame("lwr_arm_%d_link" % (1 + i % 7), LinkOffset=[0.0, 0.0, 0.02 * (i // 7)], Radius=0.05, Group="Robot"))
    for i in range(32):
        eff.append(frame("", LinkOffset=list(rng.uniform([-0.8, -0.8, 0.0], [0.8, 0.8, 1.2])), Radius=0.05, Group="Obstacle"))
    maps = [("exotica/SphereCollision", {"Name": "Col", "Precision": 0.1, "EndEffector": eff})]
    return ("exotica/TimeIndexedProblem", {"Name": "cfg5", "PlanningScene": scene("lwr_simplified", "arm"), "Maps": maps, "T": 100, "tau": 0.05,
                                           "Cost": [("exotica/Task", {"Task": "Col"})]}), 4096 * 100, 120


CONFIGS = (("cfg2a", cfg2a), ("cfg2b", cfg2b), ("cfg3", cfg3), ("cfg4", cfg4), ("cfg5", cfg5))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--only", default="")
    ap.add_argument("--kernel", type=int, default=0, help="0 auto, 1 interpreter, 2 chain regs, 3 chain two-pass, 4 / 5 interpreter one- / two-pass program")
    args = ap.parse_args()
    peak = 6549.4
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak = float(json.load(f)["hbm_gbs"])
    except Exception:
        pass
    for name, make in CONFIGS:
        if args.only and name not in args.only.split(","):
            continue
        init, E, bytes_per_eval = make()
        prob = exo.Setup.create_problem(init)
        N, LP, LJ = prob.N, prob.length_Phi, prob.length_jacobian
        lim = prob.get_scene().get_kinematic_tree().get_joint_limits()
        lo = np.where(lim[:, 0] < -1e19, -1.0, lim[:, 0])
        hi = np.where(lim[:, 1] > 1e19, 1.0, lim[:, 1])
        X = np.random.default_rng(1234).uniform(lo, hi, size=(E, N))
        dx, dphi, djac = _capi.DeviceBuffer(E * N), _capi.DeviceBuffer(E * LP), _capi.DeviceBuffer(E * LJ * N)
        dx.upload(X.T)
        backend = _host.lib().exo_host_problem_backend(prob._h)
        if args.kernel:
            _capi.check(_capi.lib().exo_problem_set_kernel(backend, args.kernel))
        times = []
        for r in range(args.reps + 3):
            try:
                prob.update_batch_device(dx.ptr, E, dphi.ptr, djac.ptr, layout=exo.LAYOUT_SOA)
            except Exception as ex:  # e.g. chain kernel forced on a tree program
                print(json.dumps({"config": name, "kernel_choice": args.kernel, "skipped": str(ex)}), flush=True)
                times = None
                break
            ms = prob.last_kernel_ms()
            if r >= 3:
                times.append(ms)
        if times is None:
            continue
        kernel = _capi.lib().exo_problem_last_kernel(backend).decode()
        med = float(np.median(times))
        gbs = bytes_per_eval * E / (med * 1e-3) / 1e9
        print(json.dumps({"config": name, "evals": E, "N": N, "length_Phi": LP, "length_jacobian": LJ, "kernel": kernel, "kernel_choice": args.kernel, "ms_median": med,
                          "ms_best": float(np.min(times)), "evals_per_s": E / (med * 1e-3), "algorithmic_GBps": gbs,
                          "frac_of_measured_hbm": gbs / peak}), flush=True)
        dx.free()
        dphi.free()
        djac.free()


if __name__ == "__main__":
    main()

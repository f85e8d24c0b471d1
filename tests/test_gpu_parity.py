"""GPU parity: the CUDA path (through the C-ABI) against the CPU oracle on the same seeded inputs.

Bar (BASELINE.json north_star): |Phi_gpu - Phi_oracle| and |J_gpu - J_oracle| <= 1e-10 absolute, FP64.
Euler-angle encodings are compared modulo 2 pi (the LWR base yaw sits on the atan2
cut); the rotation-matrix encoding, which has no cuts, is checked without wrapping.
"""
import os

import numpy as np
import pytest

from conftest import robot_paths
from exotica_b200 import _capi
from helpers import angle_entries, gpu_problem_from_maps, max_abs_err, random_states, record, run_gpu

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _oracle(oracle_lib, robot, group, **kw):
    return oracle_lib.Oracle(oracle_lib.load_robot(*robot_paths(robot), joint_group=group, **kw))


def _check(orc, maps, X, layout=_capi.LAYOUT_SOA, T=1, generated=None, tol=TOL):
    orc.set_problem(maps)
    tree, prob = gpu_problem_from_maps(orc, maps)
    if generated is not None:
        prob.set_generated(generated)
    assert prob.length_phi == orc.length_phi and prob.length_jacobian == orc.length_jacobian
    phi_ref, jac_ref = orc.problem_update_batch(X, T=T, generated=generated)
    wrap = angle_entries(maps, orc)
    kernels = set()
    # 1: general interpreter, 0: auto, 2 / 3: the chain kernel's variants (serial-chain programs only),
    # 4 / 5: the interpreter with the one-pass / two-pass program, 6: SphereCollision fused into the interpreter
    # (the default runs it as a second launch, exo_spheres_kernel), 7 / 8: the chain kernel with TMA stores (64 / 128 evaluations
    # per tile; SoA buffers with an even row pitch and chains that own every column only -- auto takes it when it applies)
    has_spheres = any(m["type"] == "SphereCollision" for m in maps)
    # 9: the per-robot generated kernel for tree programs (run-time compiled; EffFrame / EffPosition / JointLimit maps with
    # constant base frames, SoA buffers with an even row pitch)
    for choice in (1, 0, 2, 3, 4, 5, 6, 7, 8, 9):
        prob.set_kernel(choice)
        if choice in (2, 3, 7, 8) and not any(k.startswith("exo_chain") for k in kernels):
            continue
        if choice == 6 and not has_spheres:
            continue
        try:
            phi, jac = run_gpu(prob, X, layout=layout, T=T)
        except _capi.ExoticaError as err:
            if choice in (7, 8) and "TMA output path" in str(err):
                continue
            if choice == 9 and "generated tree kernel not applicable" in str(err) and "compilation failed" not in str(err):
                continue
            raise
        kernels.add(prob.last_kernel())
        e_phi = max_abs_err(phi, phi_ref, wrap)
        e_jac = max_abs_err(jac, jac_ref)
        assert e_phi <= tol, "Phi error %g (%s)" % (e_phi, prob.last_kernel())
        assert e_jac <= tol, "J error %g (%s)" % (e_jac, prob.last_kernel())
    prob.kernels_used = kernels
    return prob, phi, jac


@pytest.mark.parametrize("rot", ["RPY", "Quaternion", "ZYX", "ZYZ", "AngleAxis", "Matrix"])
def test_lwr_eff_frame_all_rotation_types(oracle_lib, rot):
    orc = _oracle(oracle_lib, "lwr_simplified", "arm")
    X = random_states(orc, 4096, seed=1234)
    _check(orc, [{"type": "EffFrame", "Type": rot, "EndEffector": [{"Link": "lwr_arm_7_link"}]}], X)


@pytest.mark.parametrize("layout", [_capi.LAYOUT_SOA, _capi.LAYOUT_AOS_ROWMAJOR, _capi.LAYOUT_AOS_COLMAJOR])
def test_lwr_layouts(oracle_lib, layout):
    orc = _oracle(oracle_lib, "lwr_simplified", "arm")
    X = random_states(orc, 3001, seed=7)  # ragged: not a multiple of the CTA size
    maps = [{"type": "EffFrame", "Type": "Matrix", "EndEffector": [{"Link": "lwr_arm_7_link"}]},
            {"type": "JointLimit", "SafePercentage": 0.1}]
    _check(orc, maps, X, layout=layout)


def test_lwr_example_ik_config(oracle_lib):
    # cfg 1: example_ik.xml -- EffFrame(RPY) on lwr_arm_6_link with a quaternion LinkOffset
    orc = _oracle(oracle_lib, "lwr_simplified", "arm")
    maps = [{"type": "EffFrame", "EndEffector": [{"Link": "lwr_arm_6_link",
             "LinkOffset": [0, 0, 0, 0.7071067811865476, -4.3297802811774664e-17, 0.7071067811865475, 4.3297802811774664e-17]}]}]
    X = np.vstack([np.zeros((1, 7)), random_states(orc, 1023, seed=3)])
    prob, phi, jac = _check(orc, maps, X)
    # link 6 leaves column 6 to no chain joint: the per-lane store kernel runs it, the chain kernel's TMA path declines
    assert prob.kernels_used == {"exo_eval_kernel", "exo_chain_kernel", "exo_tree_kernel"}
    assert np.all(jac[:, :, 6] == 0.0)  # joint 6 moves link 7 only


def test_lwr_eff_position_and_single_eval(oracle_lib):
    orc = _oracle(oracle_lib, "lwr_simplified", "arm")
    maps = [{"type": "EffPosition", "EndEffector": [{"Link": "lwr_arm_7_link"}, {"Link": "lwr_arm_3_link", "LinkOffset": [0.1, 0.2, 0.3]}]}]
    _check(orc, maps, random_states(orc, 1, seed=11))
    _check(orc, maps, random_states(orc, 2048, seed=12))


def test_relative_frames_and_base_offsets(oracle_lib):
    # frame B != root, with offsets on both sides: exercises the subtracting B-chain walk (kinematic_tree.cpp:934-944)
    orc = _oracle(oracle_lib, "lwr_simplified", "arm")
    maps = [{"type": "EffFrame", "Type": "Matrix", "EndEffector": [
        {"Link": "lwr_arm_7_link", "LinkOffset": [0.05, 0, 0.1, 0.1, 0.2, 0.3], "Base": "lwr_arm_2_link", "BaseOffset": [0, 0.1, 0, 0.3, -0.2, 0.1]},
        {"Link": "lwr_arm_1_link", "Base": "lwr_arm_5_link"},
        {"Link": "lwr_arm_4_link", "Base": "", "BaseOffset": [0.3, 0.2, 0.1, 0, 0, 0.5]},
        {"Link": "base", "Base": "lwr_arm_3_link"}]}]
    _check(orc, maps, random_states(orc, 2048, seed=21))


def test_raw_frames_map(oracle_lib):
    orc = _oracle(oracle_lib, "lwr_simplified", "arm")
    frames = [{"Link": "lwr_arm_%d_link" % i} for i in (7, 5, 2)]
    orc.request_frames(frames)
    X = random_states(orc, 257, seed=5)
    tree, prob = gpu_problem_from_maps(orc, [{"type": "Frames", "EndEffector": frames}])
    phi, jac = run_gpu(prob, X)
    for e in range(X.shape[0]):
        orc.update(X[e])
        assert max_abs_err(phi[e].reshape(3, 12), orc.phi()) <= TOL
        assert max_abs_err(jac[e].reshape(3, 6, 7), orc.jacobian()) <= TOL


def test_joint_limit_branches(oracle_lib):
    orc = _oracle(oracle_lib, "lwr_simplified", "arm")
    X = random_states(orc, 4096, seed=31, spread=0.2)
    X[0] = orc.lower  # exactly on the limit: phi 0, J 0 (App. C #3)
    X[1] = orc.upper
    for safe in (0.0, 0.2):
        prob, phi, jac = _check(orc, [{"type": "JointLimit", "SafePercentage": safe}], X)
    assert (np.abs(phi) > 0).any() and (phi == 0).any()


def test_time_indexed_style_problem_with_unused_map(oracle_lib):
    # cfg 3 shape: EffFrame + JointLimit per (b,t); plus an unused quaternion EffFrame that only gets the SetZero fill
    orc = _oracle(oracle_lib, "lwr_simplified", "arm")
    maps = [{"type": "EffFrame", "EndEffector": [{"Link": "lwr_arm_7_link"}]},
            {"type": "EffFrame", "Type": "Quaternion", "used": False, "EndEffector": [{"Link": "lwr_arm_5_link"}]},
            {"type": "JointLimit"}]
    B, T = 16, 100
    rng = np.random.default_rng(99)
    q0 = random_states(orc, B, seed=98)
    X = np.clip(q0[:, None, :] + np.cumsum(rng.normal(0, 0.05, size=(B, T, 7)), axis=1), orc.lower - 0.1, orc.upper + 0.1)
    prob, phi, jac = _check(orc, maps, X.reshape(B * T, 7), T=T)
    assert np.all(phi[:, 6:13] == np.array([0, 0, 0, 0, 0, 0, 1.0]))  # unused map: identity quaternion (App. C #2)


def test_three_dof_goldens_on_gpu(oracle_lib):
    # the reference's own EffPosition golden inputs (test_maps.cpp:340-347), GPU vs oracle at 1e-10
    orc = _oracle(oracle_lib, "test_robot", "arm")
    X = np.array([0.680375, -0.211234, 0.566198, 0.59688, 0.823295, -0.604897, -0.329554, 0.536459, -0.444451, 0.10794,
                  -0.0452059, 0.257742, -0.270431, 0.0268018, 0.904459]).reshape(5, 3)
    prob, phi, jac = _check(orc, [{"type": "EffPosition", "EndEffector": [{"Link": "endeff"}]}], X)
    assert abs(phi[0, 2] - 1.21417) < 1e-5


def test_valkyrie_multi_end_effector(oracle_lib):
    # cfg 4 shape: 38-DoF floating base, 4 end-effectors, branched tree
    orc = _oracle(oracle_lib, "valkyrie_sim", "whole_body")
    maps = [{"type": "EffFrame", "Type": "Matrix", "EndEffector": [
        {"Link": "leftFoot"}, {"Link": "rightFoot"}, {"Link": "leftPalm"}, {"Link": "rightPalm", "LinkOffset": [0, 0.05, 0.02]}]}]
    X = random_states(orc, 1024, seed=41)
    X[:, 3:6] = np.random.default_rng(42).uniform(-np.pi, np.pi, size=(1024, 3))
    prob, phi, jac = _check(orc, maps, X)
    info = prob.program_info()
    assert info["joint_ops"] == 6 + 6 + 6 + 3 + 7 + 7  # neck is not needed by any frame
    # relative frame across branches (hand w.r.t. the other hand), RPY away from gimbal lock
    maps = [{"type": "EffFrame", "EndEffector": [{"Link": "leftPalm", "Base": "rightPalm"}]},
            {"type": "EffPosition", "EndEffector": [{"Link": "head"}]}]
    _check(orc, maps, X[:256])


@pytest.mark.parametrize("E", [32, 3002, 4097, 8191 + 1])
def test_generated_tree_kernel(oracle_lib, E):
    """Per-robot generated kernel (tree_codegen.cpp + NVRTC, exo_problem_set_kernel(9)) on the 38-DoF humanoid: every rotation
    encoding, EffPosition and JointLimit beside the EffFrame map, ragged tiles (the TMA unit clips them), a constant base offset,
    an angle beyond the fast trig range, odd batch sizes falling back to the interpreter."""
    orc = _oracle(oracle_lib, "valkyrie_sim", "whole_body")
    maps = [{"type": "EffFrame", "Type": "Quaternion" if E == 32 else ("AngleAxis" if E == 3002 else "RPY"), "EndEffector": [
                {"Link": "leftFoot"}, {"Link": "rightFoot", "LinkOffset": [0.01, 0.02, 0.03, 0.1, 0.2, 0.3]},
                {"Link": "leftPalm", "BaseOffset": [0.2, 0.1, 0.3, 0.0, 0.3, 0.1]}, {"Link": "rightPalm"}]},
            {"type": "EffPosition", "EndEffector": [{"Link": "head"}, {"Link": "torso"}]},
            {"type": "JointLimit", "SafePercentage": 0.05}]
    if E == 8192:  # an unused map between the used ones: its SetZero values per lane, its Jacobian rows from the zero tile
        maps.insert(1, {"type": "EffFrame", "Type": "Quaternion", "used": False, "EndEffector": [{"Link": "leftForearmLink"}, {"Link": "pelvis"}]})
    X = random_states(orc, E, seed=500 + E)
    X[:, 3:6] = np.random.default_rng(E).uniform(-1.3, 1.3, size=(E, 3))
    X[min(E - 1, 40), 20] = 3.0e5  # beyond the fast sin / cos reduction: the cold path of that joint
    orc.set_problem(maps)
    tree, prob = gpu_problem_from_maps(orc, maps)
    phi_ref, jac_ref = orc.problem_update_batch(X)
    wrap = angle_entries(maps, orc)
    prob.set_kernel(9)
    if E % 2:
        with pytest.raises(_capi.ExoticaError, match="even, 16-byte aligned row pitch"):
            run_gpu(prob, X)
        prob.set_kernel(0)  # auto: the interpreter takes the odd batch
        phi, jac = run_gpu(prob, X)
        assert prob.last_kernel() == "exo_eval_kernel"
    else:
        phi, jac = run_gpu(prob, X)
        assert prob.last_kernel() == "exo_tree_kernel"
        phi_only, _ = run_gpu(prob, X, want_jacobian=False)
        assert np.array_equal(phi_only, phi)
    assert record("generated_tree_phi", max_abs_err(phi, phi_ref, wrap)) <= TOL
    assert record("generated_tree_jac", max_abs_err(jac, jac_ref)) <= TOL
    # a solver loop with a small batch: the interpreter serves the first calls, the generated kernel the later ones
    if E == 3002:
        prob.set_kernel(0)
        seen = []
        for _ in range(18):
            phi_s, jac_s = run_gpu(prob, X[:512])
            seen.append(prob.last_kernel())
        assert seen[0] == "exo_eval_kernel" and seen[-1] == "exo_tree_kernel"
        assert max_abs_err(phi_s, phi_ref[:512], wrap) <= TOL and max_abs_err(jac_s, jac_ref[:512]) <= TOL
    if E == 3002:
        prob.set_kernel(9)
        src = prob.generated_kernel_source()
        assert "exo_tree_kernel" in src and "cp.async.bulk.tensor.3d" in src


def test_generated_tree_kernel_sibling_maps(oracle_lib):
    """EffOrientation (angular rows), EffPositionXY (two linear rows) and JointPose beside an EffFrame map on the humanoid: all inside
    the generated kernel's scope; a Distance map is not (its program has no two-pass form) and says so."""
    orc = _oracle(oracle_lib, "valkyrie_sim", "whole_body")
    maps = [{"type": "EffOrientation", "Type": "ZYZ", "EndEffector": [{"Link": "leftPalm", "LinkOffset": [0, 0, 0.05, 0.2, 0.1, -0.3]}, {"Link": "head"}]},
            {"type": "EffPositionXY", "EndEffector": [{"Link": "rightFoot"}, {"Link": "torso", "BaseOffset": [0.1, 0.2, 0.0]}]},
            {"type": "JointPose", "JointMap": [7, 20, 3, 37], "JointRef": [0.1, -0.2, 0.3, 0.05]},
            {"type": "EffFrame", "Type": "Matrix", "EndEffector": [{"Link": "rightPalm"}]}]
    E = 4096
    X = random_states(orc, E, seed=611)
    X[:, 3:6] = np.random.default_rng(612).uniform(-1.2, 1.2, size=(E, 3))
    orc.set_problem(maps)
    tree, prob = gpu_problem_from_maps(orc, maps)
    phi_ref, jac_ref = orc.problem_update_batch(X)
    prob.set_kernel(9)
    phi, jac = run_gpu(prob, X)
    assert prob.last_kernel() == "exo_tree_kernel"
    assert record("generated_siblings_phi", max_abs_err(phi, phi_ref, angle_entries(maps, orc))) <= TOL
    assert record("generated_siblings_jac", max_abs_err(jac, jac_ref)) <= TOL
    prob.set_kernel(1)
    phi_i, jac_i = run_gpu(prob, X[:512])
    assert max_abs_err(phi_i, phi[:512], angle_entries(maps, orc)) <= 1e-12 and max_abs_err(jac_i, jac[:512]) <= 1e-12
    maps = [{"type": "Distance", "EndEffector": [{"Link": "leftPalm", "Base": "rightPalm"}]}]
    orc.set_problem(maps)
    tree2, prob2 = gpu_problem_from_maps(orc, maps)
    prob2.set_kernel(9)
    with pytest.raises(_capi.ExoticaError, match="generated tree kernel not applicable"):
        run_gpu(prob2, X[:64])


def test_gimbal_lock_set(oracle_lib):
    # pitch within 1e-13 of +-pi/2 on the floating base: the |pitch| > pi/2 - 1e-12 branch of GetRPY must agree
    orc = _oracle(oracle_lib, "valkyrie_sim", "whole_body")
    X = np.zeros((64, 38))
    rng = np.random.default_rng(5)
    X[:, 3] = rng.uniform(-1, 1, 64)
    X[:, 5] = rng.uniform(-1, 1, 64)
    X[:32, 4] = np.pi / 2
    X[32:, 4] = -np.pi / 2
    maps = [{"type": "EffFrame", "EndEffector": [{"Link": "pelvis"}]}]
    orc.set_problem(maps)
    tree, prob = gpu_problem_from_maps(orc, maps)
    phi_ref, jac_ref = orc.problem_update_batch(X)
    phi, jac = run_gpu(prob, X)
    assert max_abs_err(jac, jac_ref) <= TOL
    assert max_abs_err(phi[:, :3], phi_ref[:, :3]) <= TOL
    # at the singularity roll/yaw are only defined as a sum; both sides take the same branch (roll == 0)
    same_branch = (phi[:, 3] == 0) == (phi_ref[:, 3] == 0)
    assert same_branch.all()
    assert record("gimbal_lock_rpy", max_abs_err(phi[:, 3:6], phi_ref[:, 3:6], np.arange(3))) <= TOL


def test_sphere_collision_goldens_and_fixture(oracle_lib):
    orc = _oracle(oracle_lib, "test_robot", "arm")
    X = np.array([-0.590167, 1.2309, 1.67611, -1.72098, 1.79731, 0.103981, -1.65578, -1.23114, 0.652908, 1.56093,
                  -0.604428, -1.74331, -1.91991, -0.169193, -1.74762]).reshape(5, 3)
    maps = [{"type": "SphereCollision", "Precision": 1e-2,
             "EndEffector": [{"Link": "base", "Radius": 0.3, "Group": "base"}, {"Link": "endeff", "Radius": 0.3, "Group": "eff"}]}]
    prob, phi, jac = _check(orc, maps, X)
    assert abs(jac[0, 0, 1] - 0.431392) < 1e-5  # reference golden (test_maps.cpp:667)


def _lwr_sphere_maps():
    robot = [("lwr_arm_7_link", None), ("lwr_arm_6_link", None)]
    for k in (5, 4, 3, 2, 1):
        robot += [("lwr_arm_%d_link" % k, [0, 0, 0.1]), ("lwr_arm_%d_link" % k, None)]
    frames = []
    for link, off in robot:
        f = {"Link": link, "Radius": 0.1, "Group": "Robot"}
        if off:
            f["LinkOffset"] = off
        frames.append(f)
    for y in (-0.3, -0.2, -0.1, 0.0, 0.1, 0.2, 0.3):
        frames.append({"Link": "Obstacle", "Radius": 0.15, "Group": "Obstacle", "LinkOffset": [0, y, 0]})
    return [{"type": "SphereCollision", "Precision": 0.1, "EndEffector": frames}]


def test_sphere_collision_lwr_moving_obstacle(oracle_lib):
    # cfg 5 fixture shape (example_sphere_collision.xml: 12 robot + 7 obstacle spheres, obstacle link trajectory-driven)
    tree = oracle_lib.load_robot(*robot_paths("lwr_simplified"), joint_group="arm")
    tree.add_environment_element("Obstacle", generated_slot=0)
    orc = oracle_lib.Oracle(tree)
    B, T = 8, 25
    gen = np.zeros((T, 1, 12))
    for t in range(T):
        gen[t, 0] = oracle_lib.get_frame([0.5 - 0.02 * t, 0.1, 0.6 + 0.01 * t, 0, 0, 0.1 * t])
    X = random_states(orc, B * T, seed=77)
    maps = _lwr_sphere_maps() + [{"type": "EffPosition", "EndEffector": [{"Link": "Obstacle", "LinkOffset": [0, 0.3, 0]}]}]
    prob, phi, jac = _check(orc, maps, X, T=T, generated=gen)
    assert np.all(jac[:, 1:, :] == 0)  # the obstacle has no controlled ancestors


def test_sphere_collision_split_launch_in_chunks(oracle_lib, monkeypatch):
    # two SphereCollision maps (two parked slot sets per evaluation), a trajectory-driven obstacle, and a scratch buffer so small
    # that the batch goes through the interpreter + exo_spheres_kernel pair in several chunks of whole trajectories
    monkeypatch.setenv("EXO_SPH_DUMP_MIB", "1")
    tree = oracle_lib.load_robot(*robot_paths("lwr_simplified"), joint_group="arm")
    tree.add_environment_element("Obstacle", generated_slot=0)
    orc = oracle_lib.Oracle(tree)
    B, T = 120, 25
    gen = np.zeros((T, 1, 12))
    for t in range(T):
        gen[t, 0] = oracle_lib.get_frame([0.5 - 0.02 * t, 0.1, 0.6 + 0.01 * t, 0, 0, 0.1 * t])
    X = random_states(orc, B * T - 7, seed=79)  # ragged: the last chunk ends inside a trajectory
    second = _lwr_sphere_maps()
    second[0]["Precision"] = 0.02
    maps = _lwr_sphere_maps() + [{"type": "EffFrame", "EndEffector": [{"Link": "lwr_arm_7_link"}]}] + second
    prob, phi, jac = _check(orc, maps, X, T=T, generated=gen)
    assert "exo_spheres_kernel" in prob.kernels_used and "exo_eval_kernel" in prob.kernels_used
    # minimum surface distance per evaluation through the chunked path
    E = X.shape[0]
    dmin = _capi.DeviceBuffer(E)
    prob.set_kernel(0)
    prob.set_min_distance_output(dmin.ptr)
    run_gpu(prob, X, T=T)
    d_split = dmin.download((E,)).copy()
    prob.set_kernel(6)
    run_gpu(prob, X, T=T)
    assert np.array_equal(d_split, dmin.download((E,)))
    prob.set_min_distance_output(0)


@pytest.mark.parametrize("robot", ["lwr", "valkyrie"])
def test_sphere_collision_generated_front_half(oracle_lib, robot, monkeypatch):
    """Batches of >= EXO_TREE_MIN_EVALS evaluations of a SphereCollision-only problem: the rows exo_spheres_kernel reads are parked by
    the generated straight-line kernel (tree_park_codegen) instead of the interpreter -- a serial arm with a trajectory-driven
    obstacle, and a branched robot whose limbs collide with each other (frame slots read back by later joints).  The same
    Phi / J as with the interpreter's front half to 1e-12, and the oracle's to 1e-10."""
    if robot == "lwr":
        tree = oracle_lib.load_robot(*robot_paths("lwr_simplified"), joint_group="arm")
        tree.add_environment_element("Obstacle", generated_slot=0)
        orc = oracle_lib.Oracle(tree)
        B, T = 200, 25
        gen = np.zeros((T, 1, 12))
        for t in range(T):
            gen[t, 0] = oracle_lib.get_frame([0.5 - 0.02 * t, 0.1, 0.6 + 0.01 * t, 0, 0, 0.1 * t])
        maps = _lwr_sphere_maps()
        X = random_states(orc, B * T - 3, seed=179)
    else:
        orc = _oracle(oracle_lib, "valkyrie_sim", "whole_body")
        frames = [{"Link": link, "Radius": 0.08, "Group": grp, "LinkOffset": [0.01, 0.02, 0.03]}
                  for link, grp in (("leftPalm", "a"), ("leftForearmLink", "a"), ("rightPalm", "b"), ("rightForearmLink", "b"), ("torso", "c"), ("pelvis", "c"),
                                    ("leftFoot", "b"))]
        maps = [{"type": "SphereCollision", "Precision": 0.05, "EndEffector": frames}]
        T, gen = 1, None
        X = random_states(orc, 4200, seed=180)
    orc.set_problem(maps)
    tree_gpu, prob = gpu_problem_from_maps(orc, maps)
    if gen is not None:
        prob.set_generated(gen)
    if os.path.exists("/tmp/exo_park_test.cu"):
        os.remove("/tmp/exo_park_test.cu")
    monkeypatch.setenv("EXO_TREE_DUMP", "/tmp/exo_park_test")  # the park kernel really was compiled for this problem
    phi, jac = run_gpu(prob, X, T=T)
    assert prob.last_kernel() == "exo_spheres_kernel" and prob.launch_count() == 2
    assert "exo_tree_park_kernel" in open("/tmp/exo_park_test.cu").read()
    monkeypatch.delenv("EXO_TREE_DUMP")
    sel = np.random.default_rng(3).choice(X.shape[0], 600, replace=False)
    sel -= sel % T  # whole timesteps of the generated table: evaluation e is timestep e % T
    sel = np.unique(np.concatenate([sel + t for t in range(T)]))[:1500]
    phi_ref, jac_ref = zip(*[orc.problem_update_batch(X[e:e + 1], T=1, generated=None if gen is None else gen[e % T:e % T + 1]) for e in sel[:300]])
    assert record("park_phi", max_abs_err(phi[sel[:300]], np.concatenate(phi_ref))) <= TOL
    assert record("park_jac", max_abs_err(jac[sel[:300]], np.concatenate(jac_ref))) <= TOL
    # the interpreter's front half (small batches take it): the same numbers up to the rounding of the folded constant products
    n = 1024 - (1024 % T)
    phi_i, jac_i = run_gpu(prob, X[:n], T=T)
    assert max_abs_err(phi_i, phi[:n]) <= 1e-12 and max_abs_err(jac_i, jac[:n]) <= 1e-12


def test_sphere_collision_three_groups_robot_vs_robot(oracle_lib):
    orc = _oracle(oracle_lib, "valkyrie_sim", "whole_body")
    frames = []
    for link, grp in (("leftPalm", "a"), ("leftForearmLink", "a"), ("rightPalm", "b"), ("rightForearmLink", "b"),
                      ("torso", "c"), ("pelvis", "c"), ("leftFoot", "b")):
        frames.append({"Link": link, "Radius": 0.08, "Group": grp, "LinkOffset": [0.01, 0.02, 0.03]})
    maps = [{"type": "SphereCollision", "Precision": 0.05, "EndEffector": frames}]
    X = random_states(orc, 512, seed=78)
    prob, phi, jac = _check(orc, maps, X)
    assert phi.shape[1] == 3


def test_mimic_and_uncontrolled_joints(oracle_lib):
    # a second planning group leaves joints uncontrolled at a non-zero model state; add a mimic joint by hand
    tree = oracle_lib.load_robot(*robot_paths("lwr_simplified"), joint_group="arm")
    # make joint 5 mimic joint 1 (App. C #5: moves in FK, no Jacobian column of its own)
    e5 = tree.link_index["lwr_arm_6_link"]
    e1 = tree.link_index["lwr_arm_2_link"]
    tree.elements[e5].mimic_elem = e1
    tree.elements[e5].mimic_mult = -0.5
    tree.elements[e5].mimic_off = 0.3
    # and joint 3 uncontrolled, frozen at 0.7
    e3 = tree.link_index["lwr_arm_4_link"]
    for e in tree.elements:
        if e.control_id > 3:
            e.control_id -= 1
    tree.elements[e3].control_id = -1
    tree.controlled_joint_names.pop(3)
    tree.n_controlled = 6
    tree.lower = np.delete(tree.lower, 3)
    tree.upper = np.delete(tree.upper, 3)
    tree.tree_state[e3] = 0.7
    orc = oracle_lib.Oracle(tree)
    maps = [{"type": "EffFrame", "Type": "Matrix", "EndEffector": [{"Link": "lwr_arm_7_link"}]}]
    prob, phi, jac = _check(orc, maps, random_states(orc, 512, seed=55))
    assert prob.program_info()["joint_ops"] == 6


@pytest.mark.parametrize("k", [1, 2, 3, 4, 5, 6, 7])
def test_chain_kernels_of_every_length(oracle_lib, k):
    # a frame on link k of the LWR makes a chain of k joints: every NJ instantiation of the chain kernel and of the fused
    # cost-gradient kernel (the columns of the joints above link k are structural zeros)
    orc = _oracle(oracle_lib, "lwr_simplified", "arm")
    maps = [{"type": "EffFrame", "Type": "ZYX", "EndEffector": [{"Link": "lwr_arm_%d_link" % k, "LinkOffset": [0.02, 0.0, 0.05]}]}]
    X = random_states(orc, 257, seed=90 + k)
    prob, phi, jac = _check(orc, maps, X)
    assert prob.chain_info()["joints"] == k and not jac[:, :, k:].any()
    rho, goal = [3.0], np.array([0.2, 0.1, 0.5, 0.1, -0.2, 0.3])
    prob.set_kernel(0)
    prob.set_cost(rho, T=1, goal=goal)
    cost, grad = _gpu_cost_gradient_fused(prob, X, 257, 1)
    assert prob.last_kernel() == "exo_chain_costgrad_kernel"
    for e in range(0, 257, 16):
        gref = orc.trajectory_cost_jacobian(X[e:e + 1], rho, goal=goal)[0]
        assert np.max(np.abs(grad[e] - gref)) <= 1e-10 * max(1.0, np.abs(gref).max())
        assert abs(cost[e] - orc.trajectory_cost(X[e:e + 1], rho, goal=goal)[0]) <= 1e-10 * max(1.0, cost[e])


def test_joint_angles_far_outside_the_fast_trig_range(oracle_lib):
    # |theta| > 1e5 rad leaves the Cody-Waite reduction of the kernels' own sincos: the chain kernel takes its cold path through
    # CUDA's sincos (Payne-Hanek), the interpreter and the fused kernels branch per joint.  Mixed batch: every 7th state is huge.
    orc = _oracle(oracle_lib, "lwr_simplified", "arm")
    maps = [{"type": "EffFrame", "Type": "Matrix", "EndEffector": [{"Link": "lwr_arm_7_link"}]}, {"type": "JointLimit"}]
    X = random_states(orc, 700, seed=95)
    X[::7] += np.random.default_rng(96).choice([-1.0, 1.0], size=X[::7].shape) * 10.0 ** np.random.default_rng(97).uniform(5.2, 9.0, size=X[::7].shape)
    # sin / cos of a 1e9 argument amplify the last bit of q: absolute parity holds because both sides see the same doubles
    prob, phi, jac = _check(orc, maps, X, tol=1e-9)
    orc.set_problem(maps)
    rho, goal = [2.0, 0.0], orc.zero_goal()
    prob.set_kernel(0)
    prob.set_cost(rho, T=1, goal=goal)
    cost, grad = _gpu_cost_gradient_fused(prob, X, 700, 1)
    assert prob.last_kernel() == "exo_chain_costgrad_kernel"
    for e in range(0, 700, 7):
        gref = orc.trajectory_cost_jacobian(X[e:e + 1], rho, goal=goal)[0]
        assert np.max(np.abs(grad[e] - gref)) <= 1e-8 * max(1.0, np.abs(gref).max())


def test_prismatic_joints(oracle_lib, monkeypatch):
    # no fixture robot has a sliding joint: turn two LWR joints (and the test robot's middle joint) into prismatic ones.  Covers the
    # not-all-revolute instantiations of the chain kernels (update, fused cost / gradient, IK line search) and the interpreter's
    # prismatic branch, including Jacobian columns (w = 0, v = axis) and the Hessian of a mixed chain.
    from oracle import robot_model
    tree = oracle_lib.load_robot(*robot_paths("lwr_simplified"), joint_group="arm")
    for name in ("lwr_arm_3_link", "lwr_arm_6_link"):
        tree.elements[tree.link_index[name]].joint_type = robot_model.TRANS_AXIS
    tree.lower = np.where(np.isin(np.arange(7), (2, 5)), -0.2, tree.lower)
    tree.upper = np.where(np.isin(np.arange(7), (2, 5)), 0.2, tree.upper)
    orc = oracle_lib.Oracle(tree)
    maps = [{"type": "EffFrame", "Type": "Quaternion", "EndEffector": [{"Link": "lwr_arm_7_link", "LinkOffset": [0.0, 0.05, 0.1]}]},
            {"type": "JointLimit", "SafePercentage": 0.1},
            {"type": "Distance", "EndEffector": [{"Link": "lwr_arm_7_link", "BaseOffset": [0.3, 0.2, 0.4]}]}]
    X = random_states(orc, 500, seed=81)
    prob, phi, jac = _check(orc, maps, X)
    assert "exo_chain_kernel" in prob.kernels_used and prob.chain_info()["all_revolute"] is False
    assert not jac[:, 3:6, 2].any() and not jac[:, 3:6, 5].any()  # a sliding joint does not rotate the end-effector
    # fused cost + gradient against the oracle on the same mixed chain
    maps2 = maps[:2]
    orc.set_problem(maps2)
    tree2, prob2 = gpu_problem_from_maps(orc, maps2)
    B, T = 6, 9
    Xt = np.clip(random_states(orc, B, seed=82)[:, None, :] + np.cumsum(np.random.default_rng(83).normal(0, 0.03, size=(B, T, 7)), axis=1), orc.lower, orc.upper)
    rho = np.ones((T, 2)) * [4.0, 30.0]
    goal = orc.zero_goal()
    goal[:7] = [0.3, 0.1, 0.7, 0.0, 0.0, 0.0, 1.0]
    W = np.linspace(1.0, 2.0, 7)
    prob2.set_cost(rho, T=T, goal=goal, ct=0.5, W=W)
    cost, grad = _gpu_cost_gradient_fused(prob2, Xt.reshape(B * T, 7), B, T)
    assert prob2.last_kernel() == "exo_chain_costgrad_kernel"
    for b in range(B):
        gref = orc.trajectory_cost_jacobian(Xt[b], rho, goal=goal, ct=0.5, W=W)
        assert np.max(np.abs(grad.reshape(B, T, 7)[b] - gref)) <= 1e-10 * max(1.0, np.abs(gref).max())
        assert abs(cost[b] - orc.trajectory_cost(Xt[b], rho, goal=goal, ct=0.5, W=W)[0]) <= 1e-10 * max(1.0, cost[b])
    # Hessians of the mixed chain
    orc.set_problem(maps[:1])
    tree3, prob3 = gpu_problem_from_maps(orc, maps[:1])
    Hs = _gpu_hessian(prob3, X[:40])
    for e in range(40):
        assert np.max(np.abs(Hs[e] - orc.problem_hessian(X[e]))) <= 1e-10


def test_host_pipeline_matches_device_path(oracle_lib):
    orc = _oracle(oracle_lib, "lwr_simplified", "arm")
    maps = [{"type": "EffFrame", "EndEffector": [{"Link": "lwr_arm_7_link"}]}, {"type": "JointLimit"}]
    orc.set_problem(maps)
    tree, prob = gpu_problem_from_maps(orc, maps)
    E = 200_003
    X = random_states(orc, E, seed=61)
    phi_d, jac_d = run_gpu(prob, X, layout=_capi.LAYOUT_AOS_ROWMAJOR)
    phi_h, jac_h = prob.update_batch_host(X, layout=_capi.LAYOUT_AOS_ROWMAJOR)
    assert np.array_equal(phi_d, phi_h) and np.array_equal(jac_d, jac_h)
    phi_s, jac_s = prob.update_batch_host(np.ascontiguousarray(X.T), layout=_capi.LAYOUT_SOA)
    assert np.array_equal(phi_s.T, phi_h) and np.array_equal(jac_s.transpose(2, 0, 1), jac_h)
    phi_c, jac_c = prob.update_batch_host(X, layout=_capi.LAYOUT_AOS_COLMAJOR)
    assert np.array_equal(jac_c.transpose(0, 2, 1), jac_h)
    # spot-check against the oracle
    sel = np.random.default_rng(0).integers(0, E, 512)
    phi_ref, jac_ref = orc.problem_update_batch(X[sel])
    assert max_abs_err(phi_h[sel], phi_ref, angle_entries(maps, orc)) <= TOL
    assert max_abs_err(jac_h[sel], jac_ref) <= TOL


def test_full_size_properties_1m(oracle_lib):
    """BASELINE cfg 2a at full size (2^20 configurations): size-independent properties + sampled oracle parity."""
    orc = _oracle(oracle_lib, "lwr_simplified", "arm")
    maps = [{"type": "EffFrame", "Type": "Matrix", "EndEffector": [{"Link": "lwr_arm_7_link"}]}]
    orc.set_problem(maps)
    tree, prob = gpu_problem_from_maps(orc, maps)
    E = 1 << 20
    X = random_states(orc, E, seed=1234 + 2)
    phi, jac = run_gpu(prob, X)
    R = phi[:, 3:12].reshape(E, 3, 3)
    # rotation blocks stay orthonormal with det +1
    assert np.abs(np.einsum("eij,ekj->eik", R, R) - np.eye(3)).max() < 1e-12
    assert np.abs(np.linalg.det(R) - 1).max() < 1e-12
    # reach bound of the arm: |p - shoulder| <= sum of link lengths
    assert np.linalg.norm(phi[:, :3] - np.array([0, 0, 0.3305]), axis=1).max() <= 0.2 * 3 + 0.19 + 0.078 + 1e-12
    # last joint axis is the tool z axis: angular Jacobian column 6 == R[:, :, 2]
    assert np.abs(jac[:, 3:6, 6] - R[:, :, 2]).max() < 1e-12
    # permutation equivariance: evaluating a shuffled batch shuffles the outputs, bit for bit
    perm = np.random.default_rng(1).permutation(E)[: 1 << 16]
    phi_p, jac_p = run_gpu(prob, X[perm])
    assert np.array_equal(phi_p, phi[perm]) and np.array_equal(jac_p, jac[perm])
    sel = np.random.default_rng(2).integers(0, E, 4096)
    phi_ref, jac_ref = orc.problem_update_batch(X[sel])
    assert max_abs_err(phi[sel], phi_ref) <= TOL and max_abs_err(jac[sel], jac_ref) <= TOL


def test_errors_are_loud(oracle_lib):
    orc = _oracle(oracle_lib, "lwr_simplified", "arm")
    tree, prob = gpu_problem_from_maps(orc, [{"type": "JointLimit"}])
    with pytest.raises(_capi.ExoticaError):
        prob.update_batch_host(np.zeros((4, 6)))  # wrong state size
    p2 = _capi.Problem(tree)
    with pytest.raises(_capi.ExoticaError):
        p2.add_task_map(_capi.MAP_EFF_POSITION, elem_a=[999])  # unknown link
    with pytest.raises(_capi.ExoticaError):
        p2.update_batch_device(None, 1, None, None)  # not finalized


# ---- task cost and per-trajectory reductions (SURVEY.md section 8 rows e, f1) ------------------------------------------------
def _gpu_costs(prob, X, B, T, rho, goal=None, ct=1.0, W=None, layout=_capi.LAYOUT_SOA):
    """Runs the update kernel, then the cost kernel on its device outputs; returns (cost[B], cost_eval[B*T])."""
    E, N = X.shape
    LP, LJ = prob.length_phi, prob.length_jacobian
    dx, dphi, djac = _capi.DeviceBuffer(E * N), _capi.DeviceBuffer(E * LP), _capi.DeviceBuffer(E * LJ * N)
    dct, dce = _capi.DeviceBuffer(B), _capi.DeviceBuffer(E)
    dx.upload(X.T if layout == _capi.LAYOUT_SOA else X)
    prob.update_batch_device(dx.ptr, E, dphi.ptr, djac.ptr, T=T, layout=layout)
    prob.set_cost(rho, T=T, goal=goal, ct=ct, W=W)
    prob.cost_batch_device(dx.ptr, dphi.ptr, B, dct.ptr, dce.ptr, layout=layout)
    _capi.check(_capi.lib().exo_device_synchronize())
    _check_update_cost(prob, dx, B, T, layout, dphi, djac, dct, dce)
    return dct.download((B,)), dce.download((E,))


def _check_update_cost(prob, dx, B, T, layout, dphi, djac, dct, dce):
    """exo_problem_update_cost_batch (Update + GetCost in one call) against the two separate calls whose results sit in dphi / djac /
    dct / dce: through the generated kernel with the cost epilogue where the problem and the layout qualify, through the two-call
    path otherwise."""
    E = B * T
    LP, LJ, N = prob.length_phi, prob.length_jacobian, prob.n
    fused = False
    if layout == _capi.LAYOUT_SOA and E % 2 == 0:
        try:
            fused = "exo_tree_update_cost_kernel" in prob.generated_kernel_source(update_cost=True)
        except _capi.ExoticaError:
            fused = False
    dphi2, djac2, dct2, dce2 = (_capi.DeviceBuffer(n) for n in (E * LP, E * LJ * N, B, E))
    for b in (dphi2, djac2, dct2, dce2):
        b.fill_bytes(0xFF)
    if fused:
        prob.set_kernel(9)  # small batches: ask for the generated kernel (choice 0 keeps them on the two-call path)
    try:
        prob.update_cost_batch_device(dx.ptr, B, dphi2.ptr, djac2.ptr, dct2.ptr, dce2.ptr, layout=layout)
        _capi.check(_capi.lib().exo_device_synchronize())
        assert prob.last_kernel() == ("exo_tree_update_cost_kernel" if fused else "exo_cost_kernel")
        # without the per-evaluation costs: the library's own scratch
        dct3 = _capi.DeviceBuffer(B)
        dct3.fill_bytes(0xFF)
        prob.update_cost_batch_device(dx.ptr, B, dphi2.ptr, djac2.ptr, dct3.ptr, None, layout=layout)
        _capi.check(_capi.lib().exo_device_synchronize())
    finally:
        if fused:
            prob.set_kernel(0)
    phi_a, phi_b = dphi.download((E * LP,)), dphi2.download((E * LP,))
    jac_a, jac_b = djac.download((E * LJ * N,)), djac2.download((E * LJ * N,))
    assert record("update_cost_phi", np.max(np.abs(phi_a - phi_b))) <= 1e-12 and record("update_cost_jac", np.max(np.abs(jac_a - jac_b))) <= 1e-12
    ce_a, ce_b = dce.download((E,)), dce2.download((E,))
    ct_a, ct_b = dct.download((B,)), dct2.download((B,))
    assert record("update_cost_eval_rel", np.max(np.abs(ce_a - ce_b)) / max(1.0, np.max(np.abs(ce_a)))) <= 1e-12
    assert record("update_cost_traj_rel", np.max(np.abs(ct_a - ct_b)) / max(1.0, np.max(np.abs(ct_a)))) <= 1e-12
    assert np.array_equal(dct3.download((B,)), ct_b)


@pytest.mark.parametrize("rot", ["RPY", "Quaternion", "ZYZ", "AngleAxis", "Matrix", "ZYX"])
def test_trajectory_cost_matches_oracle(oracle_lib, rot):
    # cfg 3 shape with costs: EffFrame + JointLimit, goal with a rotated target, rho per timestep, transition cost W
    orc = _oracle(oracle_lib, "lwr_simplified", "arm")
    maps = [{"type": "EffFrame", "Type": rot, "EndEffector": [{"Link": "lwr_arm_7_link"}, {"Link": "lwr_arm_4_link", "LinkOffset": [0, 0, 0.1]}]},
            {"type": "JointLimit", "SafePercentage": 0.1}]
    orc.set_problem(maps)
    tree, prob = gpu_problem_from_maps(orc, maps)
    B, T = 12, 40
    rng = np.random.default_rng(5)
    q0 = random_states(orc, B, seed=6)
    X = np.clip(q0[:, None, :] + np.cumsum(rng.normal(0, 0.05, size=(B, T, 7)), axis=1), orc.lower - 0.1, orc.upper + 0.1)
    rho = np.ones((T, 2))
    rho[:, 0] = np.linspace(0.0, 10.0, T)
    rho[3, 1] = 0.0
    goal = orc.zero_goal()
    rl = _capi.ROTATION_LENGTH[_capi.ROTATION_TYPES[rot]]
    target = oracle_lib.get_frame([0.4, 0.1, 0.6, 0.3, -0.2, 0.5])
    for f in range(2):
        goal[f * (3 + rl):f * (3 + rl) + 3] = target[9:]
        goal[f * (3 + rl) + 3:(f + 1) * (3 + rl)] = oracle_lib.set_rotation(target[:9], rot)
    W = np.linspace(1.0, 2.0, 7)
    ct = 1.0 / 0.05 / T
    cost, cost_eval = _gpu_costs(prob, X.reshape(B * T, 7), B, T, rho, goal=goal, ct=ct, W=W)
    for b in range(B):
        ref, ref_t = orc.trajectory_cost(X[b], rho, goal=goal, ct=ct, W=W)
        assert record("cost_eval_rel_%s" % rot, np.max(np.abs(cost_eval[b * T:(b + 1) * T] - ref_t)) / max(1.0, np.max(ref_t))) <= TOL
        assert record("cost_traj_rel_%s" % rot, abs(cost[b] - ref) / max(1.0, ref)) <= TOL
        assert cost_eval[b * T] == 0.0  # GetCost starts at t = 1


@pytest.mark.parametrize("robot", ["lwr", "valkyrie"])
def test_goal_per_timestep(oracle_lib, robot):
    """TimeIndexedTask keeps one goal y[t] per timestep (tasks.cpp:166-244); a binding that forwards them pushes goal[T][length_Phi]
    (exo_cost_desc.T_goal = T).  A target that moves along the trajectory: cost, its gradient (three-kernel path, fused chain kernel /
    generated tree kernel) and Update + GetCost in one call, all against the oracle; identical rows collapse to one without changing
    a bit."""
    if robot == "lwr":
        orc = _oracle(oracle_lib, "lwr_simplified", "arm")
        maps = [{"type": "EffFrame", "Type": "Quaternion", "EndEffector": [{"Link": "lwr_arm_7_link"}]}, {"type": "JointLimit", "SafePercentage": 0.1}]
        stride, lead = 7, 3
    else:
        orc = _oracle(oracle_lib, "valkyrie_sim", "whole_body")
        maps = [{"type": "EffFrame", "Type": "Quaternion", "EndEffector": [{"Link": "leftPalm"}, {"Link": "rightFoot"}]}, {"type": "JointLimit"}]
        stride, lead = 7, 3
    orc.set_problem(maps)
    N = orc.n
    B, T = 8, 16
    E = B * T
    rng = np.random.default_rng(17)
    lo, hi = np.maximum(orc.lower, -1.0), np.minimum(orc.upper, 1.0)
    X = np.clip(rng.uniform(lo, hi, size=(B, 1, N)) + np.cumsum(rng.normal(0, 0.03, size=(B, T, N)), axis=1), lo - 0.05, hi + 0.05)
    n_frames = len(maps[0]["EndEffector"])
    goals = np.tile(orc.zero_goal(), (T, 1))
    for t in range(T):
        for f in range(n_frames):
            target = oracle_lib.get_frame([0.4 + 0.01 * t, 0.1 - 0.02 * f, 0.6 - 0.005 * t, 0.3 + 0.02 * t, -0.2, 0.5 - 0.03 * t])
            goals[t, f * stride:f * stride + lead] = target[9:]
            goals[t, f * stride + lead:(f + 1) * stride] = oracle_lib.set_rotation(target[:9], "Quaternion")
    rho = np.ones((T, 2)) * [4.0, 20.0]
    W = np.linspace(1.0, 2.0, N)
    ct = 1.0 / 0.05 / T
    tree, prob = gpu_problem_from_maps(orc, maps)
    cost, cost_eval = _gpu_costs(prob, X.reshape(E, N), B, T, rho, goal=goals, ct=ct, W=W)  # also runs exo_problem_update_cost_batch
    grad = _gpu_cost_gradient(prob, X.reshape(E, N), B, T, rho, goal=goals, ct=ct, W=W).reshape(B, T, N)
    cost_f, grad_f = _gpu_cost_gradient_fused(prob, X.reshape(E, N), B, T)
    assert prob.last_kernel() == ("exo_chain_costgrad_kernel" if robot == "lwr" else "exo_tree_costgrad_kernel")
    grad_f = grad_f.reshape(B, T, N)
    for b in range(B):
        ref, ref_t = orc.trajectory_cost(X[b], rho, goal=goals, ct=ct, W=W)
        gref = orc.trajectory_cost_jacobian(X[b], rho, goal=goals, ct=ct, W=W)
        assert np.max(np.abs(cost_eval[b * T:(b + 1) * T] - ref_t)) <= TOL * max(1.0, np.max(ref_t))
        assert abs(cost[b] - ref) <= TOL * max(1.0, ref) and abs(cost_f[b] - ref) <= TOL * max(1.0, ref)
        assert np.max(np.abs(grad[b] - gref)) <= TOL * max(1.0, np.abs(gref).max())
        assert np.max(np.abs(grad_f[b] - gref)) <= TOL * max(1.0, np.abs(gref).max())
    # the same goal in every row is stored once: bit-identical results to pushing the single row
    tree, p1 = gpu_problem_from_maps(orc, maps)
    tree, pT = gpu_problem_from_maps(orc, maps)
    c1, e1 = _gpu_costs(p1, X.reshape(E, N), B, T, rho, goal=goals[3], ct=ct, W=W)
    cT, eT = _gpu_costs(pT, X.reshape(E, N), B, T, rho, goal=np.tile(goals[3], (T, 1)), ct=ct, W=W)
    assert np.array_equal(c1, cT) and np.array_equal(e1, eT)


def test_end_pose_cost_and_layouts(oracle_lib):
    orc = _oracle(oracle_lib, "lwr_simplified", "arm")
    maps = [{"type": "EffPosition", "EndEffector": [{"Link": "lwr_arm_7_link"}]}, {"type": "EffFrame", "EndEffector": [{"Link": "lwr_arm_6_link"}]}]
    orc.set_problem(maps)
    X = random_states(orc, 257, seed=9)
    for layout in (_capi.LAYOUT_SOA, _capi.LAYOUT_AOS_ROWMAJOR):
        tree, prob = gpu_problem_from_maps(orc, maps)
        cost, cost_eval = _gpu_costs(prob, X, 257, 1, [2.0, 0.5], layout=layout)
        ref = np.array([orc.trajectory_cost(X[e:e + 1], [2.0, 0.5])[0] for e in range(257)])
        assert np.max(np.abs(cost - ref)) <= 1e-10 * max(1.0, ref.max())
        assert np.array_equal(cost, cost_eval)
    # an even batch: Update + GetScalarCost in one call takes the generated kernel with the cost epilogue (T = 1, one goal, no W)
    tree, prob = gpu_problem_from_maps(orc, maps)
    cost, cost_eval = _gpu_costs(prob, X[:256], 256, 1, [2.0, 0.5])
    assert np.max(np.abs(cost - ref[:256])) <= 1e-10 * max(1.0, ref.max())


def _gpu_cost_gradient(prob, X, B, T, rho, goal=None, ct=1.0, W=None, layout=_capi.LAYOUT_SOA):
    E, N = X.shape
    LP, LJ = prob.length_phi, prob.length_jacobian
    dx, dphi, djac, dg = (_capi.DeviceBuffer(n) for n in (E * N, E * LP, E * LJ * N, E * N))
    dg.fill_bytes(0xFF)
    dx.upload(X.T if layout == _capi.LAYOUT_SOA else X)
    prob.update_batch_device(dx.ptr, E, dphi.ptr, djac.ptr, T=T, layout=layout)
    prob.set_cost(rho, T=T, goal=goal, ct=ct, W=W)
    prob.cost_jacobian_batch_device(dx.ptr, dphi.ptr, djac.ptr, B, dg.ptr, layout=layout)
    _capi.check(_capi.lib().exo_device_synchronize())
    return dg.download((N, E)).T if layout == _capi.LAYOUT_SOA else dg.download((E, N))


@pytest.mark.parametrize("layout", [_capi.LAYOUT_SOA, _capi.LAYOUT_AOS_ROWMAJOR])
def test_trajectory_cost_jacobian_matches_oracle(oracle_lib, layout):
    # GetCostJacobian: 2 ct J^T S ydiff + transition terms of t and t + 1, per (trajectory, timestep)
    orc = _oracle(oracle_lib, "lwr_simplified", "arm")
    maps = [{"type": "EffFrame", "Type": "Quaternion", "EndEffector": [{"Link": "lwr_arm_7_link"}, {"Link": "lwr_arm_4_link", "LinkOffset": [0, 0, 0.1]}]},
            {"type": "JointLimit", "SafePercentage": 0.1},
            {"type": "Distance", "EndEffector": [{"Link": "lwr_arm_6_link", "BaseOffset": [0.3, 0.2, 0.4]}]}]
    orc.set_problem(maps)
    tree, prob = gpu_problem_from_maps(orc, maps)
    B, T = 9, 33
    rng = np.random.default_rng(15)
    q0 = random_states(orc, B, seed=16)
    X = np.clip(q0[:, None, :] + np.cumsum(rng.normal(0, 0.05, size=(B, T, 7)), axis=1), orc.lower - 0.1, orc.upper + 0.1)
    rho = np.ones((T, 3))
    rho[:, 0] = np.linspace(0.0, 10.0, T)
    rho[3, 1] = 0.0
    rho[:, 2] = 3.0
    goal = orc.zero_goal()
    target = oracle_lib.get_frame([0.4, 0.1, 0.6, 0.3, -0.2, 0.5])
    for f in range(2):
        goal[f * 7:f * 7 + 3] = target[9:]
        goal[f * 7 + 3:(f + 1) * 7] = oracle_lib.set_rotation(target[:9], "Quaternion")
    W = np.linspace(1.0, 2.0, 7)
    ct = 1.0 / 0.05 / T
    grad = _gpu_cost_gradient(prob, X.reshape(B * T, 7), B, T, rho, goal=goal, ct=ct, W=W, layout=layout).reshape(B, T, 7)
    for b in range(B):
        ref = orc.trajectory_cost_jacobian(X[b], rho, goal=goal, ct=ct, W=W)
        assert np.max(np.abs(grad[b] - ref)) <= 1e-10 * max(1.0, np.abs(ref).max())
        assert not grad[b, 0].any()  # the reference's row vector starts at t = 1
    # the same through the host pipeline: only N doubles per evaluation (and one cost per trajectory) come back
    cost_h, grad_h = prob.cost_jacobian_batch_host(X)
    assert np.array_equal(grad_h, grad) or np.max(np.abs(grad_h - grad)) <= 1e-12 * np.abs(grad).max()
    ref_cost = np.array([orc.trajectory_cost(X[b], rho, goal=goal, ct=ct, W=W)[0] for b in range(B)])
    assert record("cost_host_rel", np.max(np.abs(cost_h - ref_cost)) / ref_cost.max()) <= TOL


def _gpu_cost_gradient_fused(prob, X, B, T, layout=_capi.LAYOUT_SOA):
    E, N = X.shape
    dx, dg, dc = _capi.DeviceBuffer(E * N), _capi.DeviceBuffer(E * N), _capi.DeviceBuffer(B)
    dg.fill_bytes(0xFF)
    dx.upload(X.T if layout == _capi.LAYOUT_SOA else X)
    prob.cost_gradient_batch_device(dx.ptr, B, dg.ptr, dc.ptr, layout=layout)
    _capi.check(_capi.lib().exo_device_synchronize())
    grad = dg.download((N, E)).T if layout == _capi.LAYOUT_SOA else dg.download((E, N))
    return dc.download((B,)), np.ascontiguousarray(grad)


@pytest.mark.parametrize("case", ["frame+limit", "base+position", "unused", "end-pose"])
def test_fused_cost_gradient_kernel(oracle_lib, case):
    # states in, cost + gradient out: the fused chain kernel (J never reaches HBM) against the oracle and the three-kernel path
    orc = _oracle(oracle_lib, "lwr_simplified", "arm")
    T = 25
    if case == "frame+limit":
        maps = [{"type": "EffFrame", "Type": "RPY", "EndEffector": [{"Link": "lwr_arm_7_link", "LinkOffset": [0.05, 0, 0.1, 0.9, 0.1, 0.2, 0.3]}]},
                {"type": "JointLimit", "SafePercentage": 0.2}]
        rho = np.ones((T, 2)) * [3.0, 50.0]
        rho[:, 0] = np.linspace(0.0, 5.0, T)
    elif case == "base+position":
        maps = [{"type": "EffFrame", "Type": "Quaternion", "EndEffector": [{"Link": "lwr_arm_7_link", "BaseOffset": [0.1, 0.2, 0.3, 0.5, 0.5, 0.5, 0.5]}]},
                {"type": "EffPosition", "EndEffector": [{"Link": "lwr_arm_7_link", "LinkOffset": [0, 0.1, 0]}]}]
        rho = np.ones((T, 2)) * [1.0, 7.0]
    elif case == "unused":
        maps = [{"type": "JointLimit", "SafePercentage": 0.1},
                {"type": "EffFrame", "Type": "ZYX", "used": False, "EndEffector": [{"Link": "lwr_arm_5_link"}]},
                {"type": "EffFrame", "Type": "AngleAxis", "EndEffector": [{"Link": "lwr_arm_7_link"}]}]
        rho = np.ones((T, 3)) * [10.0, 0.0, 2.0]
    else:
        T = 1
        maps = [{"type": "EffFrame", "EndEffector": [{"Link": "lwr_arm_7_link"}]}]
        rho = np.array([[4.0]])
    orc.set_problem(maps)
    tree, prob = gpu_problem_from_maps(orc, maps)
    B = 17 if T > 1 else 700
    rng = np.random.default_rng(21)
    q0 = random_states(orc, B, seed=22)
    X = np.clip(q0[:, None, :] + np.cumsum(rng.normal(0, 0.05, size=(B, T, 7)), axis=1), orc.lower - 0.1, orc.upper + 0.1)
    goal = orc.zero_goal()
    target = oracle_lib.get_frame([0.4, 0.1, 0.6, 0.3, -0.2, 0.5])
    for k, spec in enumerate(maps):
        start, length = prob.map_indexing(k)[:2]
        if spec["type"] == "EffFrame":
            goal[start:start + 3] = target[9:]
            goal[start + 3:start + length] = oracle_lib.set_rotation(target[:9], spec.get("Type", "RPY"))
        elif spec["type"] == "EffPosition":
            goal[start:start + 3] = [0.3, -0.2, 0.8]
    W = np.linspace(1.0, 2.0, 7) if T > 1 else None
    ct = 1.0 / 0.05 / T if T > 1 else 1.0
    prob.set_cost(rho, T=T, goal=goal, ct=ct, W=W)
    results = {}
    for choice, name in ((0, "exo_chain_costgrad_kernel"), (1, "exo_cost_gradient_kernel")):
        prob.set_kernel(choice)
        for layout in (_capi.LAYOUT_SOA, _capi.LAYOUT_AOS_ROWMAJOR):
            cost, grad = _gpu_cost_gradient_fused(prob, X.reshape(B * T, 7), B, T, layout=layout)
            assert prob.last_kernel() == name
            results[(choice, layout)] = (cost, grad)
            grad = grad.reshape(B, T, 7)
            for b in range(B):
                gref = orc.trajectory_cost_jacobian(X[b], rho, goal=goal, ct=ct, W=W)
                cref = orc.trajectory_cost(X[b], rho, goal=goal, ct=ct, W=W)[0]
                assert np.max(np.abs(grad[b] - gref)) <= 1e-10 * max(1.0, np.abs(gref).max()), name
                assert abs(cost[b] - cref) <= 1e-10 * max(1.0, cref), name
    # layouts agree bit for bit within a path
    for choice in (0, 1):
        a, b = results[(choice, _capi.LAYOUT_SOA)], results[(choice, _capi.LAYOUT_AOS_ROWMAJOR)]
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    # cost only (no gradient buffer): one kernel for chain programs, FK + cost kernel otherwise; same costs
    E = B * T
    dx, dc = _capi.DeviceBuffer(E * 7), _capi.DeviceBuffer(B)
    dx.upload(X.reshape(E, 7).T)
    for choice, name in ((0, "exo_chain_cost_kernel"), (1, "exo_cost_kernel" if T > 1 else "exo_cost_kernel")):
        prob.set_kernel(choice)
        prob.cost_gradient_batch_device(dx.ptr, B, None, dc.ptr)
        _capi.check(_capi.lib().exo_device_synchronize())
        assert prob.last_kernel() == name
        assert np.max(np.abs(dc.download((B,)) - results[(choice, _capi.LAYOUT_SOA)][0])) <= 1e-12 * max(1.0, np.abs(results[(choice, _capi.LAYOUT_SOA)][0]).max())
    prob.set_kernel(0)
    cost_h = prob.cost_batch_host(X if T > 1 else X.reshape(B, 7))
    assert np.max(np.abs(cost_h - results[(0, _capi.LAYOUT_SOA)][0])) <= 1e-12 * max(1.0, np.abs(cost_h).max())


@pytest.mark.parametrize("layout", [_capi.LAYOUT_SOA, _capi.LAYOUT_AOS_ROWMAJOR])
def test_aico_task_messages(oracle_lib, layout):
    # AICOSolver::GetTaskCosts: R = sum rho J^T J, r = sum rho J^T (-ydiff + J qhat), rhat, and the task cost per timestep
    orc = _oracle(oracle_lib, "lwr_simplified", "arm")
    maps = [{"type": "EffFrame", "Type": "RPY", "EndEffector": [{"Link": "lwr_arm_7_link"}]},
            {"type": "JointLimit", "SafePercentage": 0.1},
            {"type": "EffPosition", "EndEffector": [{"Link": "lwr_arm_4_link"}]}]
    orc.set_problem(maps)
    tree, prob = gpu_problem_from_maps(orc, maps)
    B, T, N = 7, 21, 7
    rng = np.random.default_rng(51)
    q0 = random_states(orc, B, seed=52)
    X = np.clip(q0[:, None, :] + np.cumsum(rng.normal(0, 0.05, size=(B, T, N)), axis=1), orc.lower - 0.1, orc.upper + 0.1)
    rho = np.ones((T, 3)) * [100.0, 1e3, 0.0]  # the third task has zero precision: skipped
    rho[:T - 1, 0] = 0.0                       # end-effector goal at the last timestep only, as AICO problems usually have it
    rho[5, 1] = -1.0                           # "if (prec > 0)": negative precisions are skipped too
    goal = orc.zero_goal()
    goal[:6] = [0.4, 0.1, 0.6, 0.3, -0.2, 0.5]
    ct = 1.0 / 0.05 / T
    E = B * T
    LP, LJ = prob.length_phi, prob.length_jacobian
    dx, dphi, djac, dR, dr, drh, dc = (_capi.DeviceBuffer(n) for n in (E * N, E * LP, E * LJ * N, E * N * N, E * N, E, E))
    for b in (dR, dr, drh, dc):
        b.fill_bytes(0xFF)
    Xe = X.reshape(E, N)
    dx.upload(Xe.T if layout == _capi.LAYOUT_SOA else Xe)
    prob.update_batch_device(dx.ptr, E, dphi.ptr, djac.ptr, T=T, layout=layout)
    prob.set_cost(rho, T=T, goal=goal, ct=ct)
    prob.aico_task_messages_device(dx.ptr, dphi.ptr, djac.ptr, B, dR.ptr, dr.ptr, drh.ptr, dc.ptr, layout=layout)
    _capi.check(_capi.lib().exo_device_synchronize())
    if layout == _capi.LAYOUT_SOA:
        R = dR.download((N, N, E)).transpose(2, 0, 1).reshape(B, T, N, N)
        r = dr.download((N, E)).T.reshape(B, T, N)
    else:
        R = dR.download((E, N, N)).reshape(B, T, N, N)
        r = dr.download((E, N)).reshape(B, T, N)
    rhat, cost = drh.download((E,)).reshape(B, T), dc.download((E,)).reshape(B, T)
    for b in range(B):
        R_ref, r_ref, rhat_ref, cost_ref = orc.aico_task_messages(X[b], rho, goal=goal, ct=ct)
        for got, ref in ((R[b], R_ref), (r[b], r_ref), (rhat[b], rhat_ref), (cost[b], cost_ref)):
            assert np.max(np.abs(got - ref)) <= 1e-10 * max(1.0, np.abs(ref).max())
        assert np.array_equal(R[b], R[b].transpose(0, 2, 1))
        assert not R[b, 5, :, :].any() or rho[5, 0] > 0  # timestep 5: JointLimit skipped (negative rho), EffFrame rho 0


def test_cost_gradient_and_aico_messages_on_a_tree_robot(oracle_lib):
    # 38 controlled joints: the general (any N) instantiations of the cost-gradient and AICO kernels; exo_problem_cost_gradient_batch
    # through the generated fused kernel of tree programs (exo_tree_costgrad_kernel: J never reaches HBM) and through the three-kernel
    # path (exo_problem_set_kernel(1)), both against the oracle and in both layouts
    orc = _oracle(oracle_lib, "valkyrie_sim", "whole_body")
    maps = [{"type": "EffFrame", "EndEffector": [{"Link": "leftPalm"}, {"Link": "rightFoot", "LinkOffset": [0.0, 0.0, -0.05]}]},
            {"type": "JointLimit", "SafePercentage": 0.1},
            {"type": "EffPosition", "EndEffector": [{"Link": "torso"}]}]
    orc.set_problem(maps)
    tree, prob = gpu_problem_from_maps(orc, maps)
    N = orc.n
    B, T = 3, 6
    rng = np.random.default_rng(71)
    lo, hi = np.maximum(orc.lower, -1.0), np.minimum(orc.upper, 1.0)
    X = np.clip(rng.uniform(lo, hi, size=(B, 1, N)) + np.cumsum(rng.normal(0, 0.03, size=(B, T, N)), axis=1), lo - 0.05, hi + 0.05)
    rho = np.ones((T, 3)) * [5.0, 100.0, 2.0]
    goal = orc.zero_goal()
    goal[:3] = [0.3, 0.4, 0.2]
    W = np.linspace(1.0, 3.0, N)
    ct = 1.0 / 0.1 / T
    prob.set_cost(rho, T=T, goal=goal, ct=ct, W=W)
    E = B * T
    prob.set_kernel(1)
    cost, grad = _gpu_cost_gradient_fused(prob, X.reshape(E, N), B, T)
    assert prob.last_kernel() == "exo_cost_gradient_kernel"
    prob.set_kernel(0)
    for layout in (_capi.LAYOUT_SOA, _capi.LAYOUT_AOS_ROWMAJOR):
        cost_f, grad_f = _gpu_cost_gradient_fused(prob, X.reshape(E, N), B, T, layout=layout)
        assert prob.last_kernel() == "exo_tree_costgrad_kernel"
        assert record("tree_costgrad_vs_unfused", np.max(np.abs(grad_f - grad)) / max(1.0, np.abs(grad).max())) <= 1e-12
        assert np.max(np.abs(cost_f - cost)) <= 1e-12 * max(1.0, np.abs(cost).max())
        for b in range(B):
            gref = orc.trajectory_cost_jacobian(X[b], rho, goal=goal, ct=ct, W=W)
            assert np.max(np.abs(grad_f.reshape(B, T, N)[b] - gref)) <= 1e-10 * max(1.0, np.abs(gref).max())
    src = prob.generated_kernel_source(cost=True)
    assert "exo_tree_costgrad_kernel" in src and "get_angle_axis(M, vx, vy, vz);" in src
    # cost only (NULL gradient) through the same generated kernel
    dxc, dcc = _capi.DeviceBuffer(E * N), _capi.DeviceBuffer(B)
    dxc.upload(X.reshape(E, N).T)
    prob.cost_gradient_batch_device(dxc.ptr, B, None, dcc.ptr)
    _capi.check(_capi.lib().exo_device_synchronize())
    assert prob.last_kernel() == "exo_tree_costgrad_kernel" and np.max(np.abs(dcc.download((B,)) - cost)) <= 1e-12 * max(1.0, np.abs(cost).max())
    grad = grad.reshape(B, T, N)
    LP, LJ = prob.length_phi, prob.length_jacobian
    dx, dphi, djac, dR, dr, drh, dc = (_capi.DeviceBuffer(n) for n in (E * N, E * LP, E * LJ * N, E * N * N, E * N, E, E))
    dx.upload(X.reshape(E, N).T)
    prob.update_batch_device(dx.ptr, E, dphi.ptr, djac.ptr, T=T)
    prob.aico_task_messages_device(dx.ptr, dphi.ptr, djac.ptr, B, dR.ptr, dr.ptr, drh.ptr, dc.ptr)
    _capi.check(_capi.lib().exo_device_synchronize())
    # Update + GetCost in one call on the tree robot (two-pass generated kernel with the cost epilogue) against the separate calls
    dct, dce = _capi.DeviceBuffer(B), _capi.DeviceBuffer(E)
    prob.cost_batch_device(dx.ptr, dphi.ptr, B, dct.ptr, dce.ptr)
    _capi.check(_capi.lib().exo_device_synchronize())
    assert np.max(np.abs(dct.download((B,)) - cost)) <= 1e-12 * max(1.0, np.abs(cost).max())
    _check_update_cost(prob, dx, B, T, _capi.LAYOUT_SOA, dphi, djac, dct, dce)
    R = dR.download((N, N, E)).transpose(2, 0, 1).reshape(B, T, N, N)
    r = dr.download((N, E)).T.reshape(B, T, N)
    rhat = drh.download((E,)).reshape(B, T)
    for b in range(B):
        gref = orc.trajectory_cost_jacobian(X[b], rho, goal=goal, ct=ct, W=W)
        cref = orc.trajectory_cost(X[b], rho, goal=goal, ct=ct, W=W)[0]
        assert np.max(np.abs(grad[b] - gref)) <= 1e-10 * max(1.0, np.abs(gref).max())
        assert abs(cost[b] - cref) <= 1e-10 * max(1.0, cref)
        R_ref, r_ref, rhat_ref, _ = orc.aico_task_messages(X[b], rho, goal=goal, ct=ct)
        for got, ref in ((R[b], R_ref), (r[b], r_ref), (rhat[b], rhat_ref)):
            assert np.max(np.abs(got - ref)) <= 1e-10 * max(1.0, np.abs(ref).max())


def test_end_pose_cost_jacobian_and_finite_differences(oracle_lib):
    # UnconstrainedEndPoseProblem::GetScalarJacobian = 2 J^T S ydiff; for position-type maps it is the exact gradient of the cost
    orc = _oracle(oracle_lib, "lwr_simplified", "arm")
    maps = [{"type": "EffPosition", "EndEffector": [{"Link": "lwr_arm_7_link"}]}, {"type": "Distance", "EndEffector": [{"Link": "lwr_arm_5_link", "BaseOffset": [0.5, 0, 0.3]}]}]
    orc.set_problem(maps)
    tree, prob = gpu_problem_from_maps(orc, maps)
    X = random_states(orc, 300, seed=19)
    rho = [2.0, 0.5]
    goal = np.array([0.4, -0.2, 0.7, 0.25])
    grad = _gpu_cost_gradient(prob, X, 300, 1, rho, goal=goal)
    ref = np.array([orc.trajectory_cost_jacobian(X[e:e + 1], rho, goal=goal)[0] for e in range(300)])
    assert np.max(np.abs(grad - ref)) <= 1e-10 * max(1.0, np.abs(ref).max())
    h = 1e-6
    for e in range(0, 300, 37):
        for n in range(7):
            xp, xm = X[e:e + 1].copy(), X[e:e + 1].copy()
            xp[0, n] += h
            xm[0, n] -= h
            fd = (orc.trajectory_cost(xp, rho, goal=goal)[0] - orc.trajectory_cost(xm, rho, goal=goal)[0]) / (2 * h)
            assert abs(fd - grad[e, n]) <= 1e-6 * max(1.0, abs(fd))


def test_min_sphere_distance_and_trajectory_min(oracle_lib):
    orc = _oracle(oracle_lib, "lwr_simplified", "arm")
    maps = _lwr_sphere_maps()
    # obstacle spheres fixed in the world instead of on a trajectory-driven link
    for f in maps[0]["EndEffector"]:
        if f["Group"] == "Obstacle":
            f["Link"] = ""
            f["LinkOffset"] = [0.45, f["LinkOffset"][1], 0.55]
    orc.set_problem(maps)
    tree, prob = gpu_problem_from_maps(orc, maps)
    B, T = 9, 33
    X = random_states(orc, B * T, seed=15)
    E, N = X.shape
    dx, dphi, djac, dmin, dtraj = (_capi.DeviceBuffer(n) for n in (E * N, E, E * N, E, B))
    dx.upload(X.T)
    prob.set_min_distance_output(dmin.ptr)
    prob.update_batch_device(dx.ptr, E, dphi.ptr, djac.ptr, T=T)
    _capi.check(_capi.lib().exo_trajectory_reduce(dmin.ptr, B, T, 0, 0, dtraj.ptr, None))
    _capi.check(_capi.lib().exo_device_synchronize())
    got, got_traj = dmin.download((E,)), dtraj.download((B,))
    # reference: centre distances from the oracle's frames
    frames = maps[0]["EndEffector"]
    orc.request_frames([{k: v for k, v in f.items() if k in ("Link", "LinkOffset")} for f in frames])
    ref = np.zeros(E)
    rob = [i for i, f in enumerate(frames) if f["Group"] == "Robot"]
    obs = [i for i, f in enumerate(frames) if f["Group"] == "Obstacle"]
    for e in range(E):
        orc.update(X[e])
        p = orc.phi()[:, 9:]
        ref[e] = min(np.linalg.norm(p[i] - p[j]) - frames[i]["Radius"] - frames[j]["Radius"] for i in rob for j in obs)
    assert np.max(np.abs(got - ref)) <= 1e-10
    assert np.max(np.abs(got_traj - ref.reshape(B, T).min(axis=1))) <= 1e-10
    orc.set_problem(maps)


# ---- sibling task maps that reuse the same frames ------------------------------------------------
@pytest.mark.parametrize("rot", ["RPY", "Quaternion", "AngleAxis", "Matrix"])
def test_sibling_maps_on_a_chain(oracle_lib, rot):
    # EffOrientation / EffPositionXY / Distance / JointPose at the chain tip: the DH chain kernel's generic row emitter
    orc = _oracle(oracle_lib, "lwr_simplified", "arm")
    maps = [{"type": "EffOrientation", "Type": rot, "EndEffector": [{"Link": "lwr_arm_7_link", "LinkOffset": [0, 0, 0.05, 0.2, 0.1, -0.3]}]},
            {"type": "EffPositionXY", "EndEffector": [{"Link": "lwr_arm_7_link"}]},
            {"type": "Distance", "EndEffector": [{"Link": "lwr_arm_7_link", "BaseOffset": [0.5, 0.2, 0.4]}]},
            {"type": "EffPosition", "EndEffector": [{"Link": "lwr_arm_7_link"}]}]
    prob, phi, jac = _check(orc, maps, random_states(orc, 2048, seed=101))
    assert "exo_chain_kernel" in prob.kernels_used


def test_sibling_maps_on_trees_and_joint_pose(oracle_lib):
    orc = _oracle(oracle_lib, "lwr_simplified", "arm")
    maps = [{"type": "JointPose", "JointMap": [6, 2, 0], "JointRef": [0.1, -0.2, 0.3]},
            {"type": "EffOrientation", "EndEffector": [{"Link": "lwr_arm_5_link"}, {"Link": "lwr_arm_7_link", "Base": "lwr_arm_2_link"}]},
            {"type": "Distance", "EndEffector": [{"Link": "lwr_arm_7_link", "Base": "lwr_arm_3_link"}, {"Link": "lwr_arm_4_link"}]},
            {"type": "JointPose"},
            {"type": "EffPositionXY", "EndEffector": [{"Link": "lwr_arm_6_link", "LinkOffset": [0.1, 0, 0]}]}]
    _check(orc, maps, random_states(orc, 1024, seed=102))
    orc = _oracle(oracle_lib, "valkyrie_sim", "whole_body")
    maps = [{"type": "EffOrientation", "Type": "Matrix", "EndEffector": [{"Link": "leftPalm"}, {"Link": "rightFoot"}]},
            {"type": "EffPositionXY", "EndEffector": [{"Link": "head"}, {"Link": "leftFoot"}]},
            {"type": "JointPose", "JointMap": list(range(18, 38))},
            {"type": "EffPosition", "EndEffector": [{"Link": "rightPalm"}]}]
    X = random_states(orc, 512, seed=103)
    _check(orc, maps, X)
    _check(orc, maps + [{"type": "Distance", "EndEffector": [{"Link": "leftPalm", "Base": "rightPalm"}]}], X)


def test_eff_axis_alignment(oracle_lib):
    # eff_axis_alignment.cpp: one row per frame, direction . (R axis) - 1; offsets with rotation and position, a relative base,
    # non-unit Axis / Direction (normalised by Initialize), chain and tree robots, all interpreter program shapes
    orc = _oracle(oracle_lib, "lwr_simplified", "arm")
    maps = [{"type": "EffAxisAlignment", "EndEffector": [
                {"Link": "lwr_arm_7_link", "Axis": [1, 0, 0], "Direction": [0, 0, 1]},
                {"Link": "lwr_arm_5_link", "LinkOffset": [0.1, -0.05, 0.2, 0.1, 0.2, 0.3, 0.9], "Axis": [0, 2, 1], "Direction": [1, 1, 0]},
                {"Link": "lwr_arm_6_link", "Base": "lwr_arm_2_link", "BaseOffset": [0.0, 0.1, 0.0, 0.5, 0.5, 0.5, 0.5], "Axis": [0, 0, 1], "Direction": [0, -1, 0]}]},
            {"type": "EffPosition", "EndEffector": [{"Link": "lwr_arm_7_link"}]}]
    X = random_states(orc, 777, seed=31)
    prob, phi, jac = _check(orc, maps, X)
    assert phi[:, :3].min() >= -2.0 - 1e-12 and phi[:, :3].max() <= 1e-12  # cos(angle) - 1 for unit vectors without position offsets...
    orc = _oracle(oracle_lib, "valkyrie_sim", "whole_body")
    maps = [{"type": "EffAxisAlignment", "EndEffector": [{"Link": "leftPalm", "Axis": [0, 0, 1], "Direction": [0, 0, -1]},
                                                          {"Link": "rightFoot", "Axis": [1, 0, 0], "Direction": [1, 0, 0]}]}]
    _check(orc, maps, random_states(orc, 300, seed=32))


@pytest.mark.parametrize("enable_z", [True, False])
def test_center_of_mass_map(oracle_lib, enable_z):
    # center_of_mass.cpp with explicit frames: mass-weighted mean of the frames' positions and of their position Jacobians
    orc = _oracle(oracle_lib, "valkyrie_sim", "whole_body")
    frames = []
    for e in orc.tree.elements:
        if e.is_robot_link and e.mass > 0:
            frames.append({"Link": e.link_name, "LinkOffset": list(e.cog), "Mass": e.mass})
    assert len(frames) > 20
    maps = [{"type": "CenterOfMass", "EnableZ": enable_z, "EndEffector": frames},
            {"type": "EffPosition", "EndEffector": [{"Link": "leftPalm"}]}]
    X = random_states(orc, 400, seed=33)
    prob, phi, jac = _check(orc, maps, X)
    assert phi.shape[1] == (3 if enable_z else 2) + 3
    # a chain robot too (one-pass program), with an unused CoM map next to it
    orc = _oracle(oracle_lib, "lwr_simplified", "arm")
    frames = [{"Link": "lwr_arm_%d_link" % k, "LinkOffset": [0.0, 0.01 * k, 0.05], "Mass": 1.0 + 0.5 * k} for k in range(1, 8)]
    maps = [{"type": "CenterOfMass", "EnableZ": enable_z, "EndEffector": frames},
            {"type": "CenterOfMass", "used": False, "EndEffector": frames[:2]},
            {"type": "JointLimit"}]
    _check(orc, maps, random_states(orc, 333, seed=34), layout=_capi.LAYOUT_AOS_ROWMAJOR)


def test_sibling_maps_cost_and_unused_fill(oracle_lib):
    orc = _oracle(oracle_lib, "lwr_simplified", "arm")
    maps = [{"type": "EffOrientation", "Type": "Quaternion", "EndEffector": [{"Link": "lwr_arm_7_link"}]},
            {"type": "EffOrientation", "Type": "Matrix", "used": False, "EndEffector": [{"Link": "lwr_arm_4_link"}]},
            {"type": "JointPose", "JointRef": list(np.linspace(-0.3, 0.3, 7))},
            {"type": "Distance", "EndEffector": [{"Link": "lwr_arm_7_link"}]}]
    orc.set_problem(maps)
    X = random_states(orc, 300, seed=104)
    prob, phi, jac = _check(orc, maps, X)
    assert np.all(phi[:, 4:13] == np.array([1.0, 0, 0, 0, 1, 0, 0, 0, 1]))  # unused map: identity matrix fill
    tree, prob = gpu_problem_from_maps(orc, maps)
    goal = orc.zero_goal()
    goal[0:4] = oracle_lib.set_rotation(oracle_lib.get_frame([0, 0, 0, 0.4, -0.3, 0.2])[:9], "Quaternion")
    cost, _ = _gpu_costs(prob, X, 300, 1, [2.0, 0.0, 0.5, 3.0], goal=goal)
    ref = np.array([orc.trajectory_cost(X[e:e + 1], [2.0, 0.0, 0.5, 3.0], goal=goal)[0] for e in range(300)])
    assert record("sibling_cost_rel", np.max(np.abs(cost - ref)) / max(1.0, ref.max())) <= TOL


# ---- BASELINE.json configurations 3, 4, 5 at full size: size-independent properties + sampled oracle parity ---------------------
def _smooth_trajectories(orc, B, T, seed):
    rng = np.random.default_rng(seed)
    q0 = random_states(orc, B, seed=seed + 1)
    return np.clip(q0[:, None, :] + np.cumsum(rng.normal(0, 0.05, size=(B, T, orc.n)), axis=1), orc.lower - 0.1, orc.upper + 0.1)


def test_full_size_cfg3_time_indexed(oracle_lib):
    orc = _oracle(oracle_lib, "lwr_simplified", "arm")
    maps = [{"type": "EffFrame", "EndEffector": [{"Link": "lwr_arm_7_link"}]}, {"type": "JointLimit"}]
    orc.set_problem(maps)
    tree, prob = gpu_problem_from_maps(orc, maps)
    B, T = 4096, 100
    X = _smooth_trajectories(orc, B, T, 1237).reshape(B * T, 7)
    phi, jac = run_gpu(prob, X, T=T)
    # JointLimit rows: phi is the signed violation, J is diag(violated) and nothing else
    viol = np.where(X < orc.lower, X - orc.lower, 0.0) + np.where(X > orc.upper, X - orc.upper, 0.0)
    assert np.array_equal(phi[:, 6:], viol)
    Jl = jac[:, 6:, :]
    assert np.array_equal(Jl, np.eye(7)[None] * ((X < orc.lower) | (X > orc.upper))[:, None, :].astype(float))
    # the last joint rotates about the tool axis: it cannot move the tool centre point
    assert np.abs(jac[:, 0:3, 6]).max() < 1e-12
    # timestep-major batches: evaluating one trajectory alone gives the same bits
    phi1, jac1 = run_gpu(prob, X[17 * T:18 * T], T=T)
    assert np.array_equal(phi1, phi[17 * T:18 * T]) and np.array_equal(jac1, jac[17 * T:18 * T])
    sel = np.random.default_rng(3).integers(0, B * T, 4096)
    phi_ref, jac_ref = orc.problem_update_batch(X[sel])
    assert max_abs_err(phi[sel], phi_ref, angle_entries(maps, orc)) <= TOL and max_abs_err(jac[sel], jac_ref) <= TOL
    # Update + GetCost of all 4096 trajectories in one call (automatic choice at this size: the generated kernel with the cost
    # epilogue) -- Phi / J against the run above, costs against the oracle
    E, N, LP, LJ = B * T, 7, prob.length_phi, prob.length_jacobian
    dx, dphi, djac, dct = (_capi.DeviceBuffer(n) for n in (E * N, E * LP, E * LJ * N, B))
    dx.upload(X.T)
    goal = orc.zero_goal()
    goal[:6] = [0.4, 0.1, 0.6, 0.2, -0.1, 0.3]
    prob.set_cost(np.ones((T, 2)) * [10.0, 1.0], T=T, goal=goal, ct=1.0 / 0.05 / T, W=np.ones(7))
    prob.update_cost_batch_device(dx.ptr, B, dphi.ptr, djac.ptr, dct.ptr)
    _capi.check(_capi.lib().exo_device_synchronize())
    assert prob.last_kernel() == "exo_tree_update_cost_kernel"
    assert np.max(np.abs(dphi.download((LP, E)).T - phi)) <= 1e-12 and np.max(np.abs(djac.download((LJ, N, E)).transpose(2, 0, 1) - jac)) <= 1e-12
    cost = dct.download((B,))
    for b in range(0, B, 128):
        ref = orc.trajectory_cost(X[b * T:(b + 1) * T], np.ones((T, 2)) * [10.0, 1.0], goal=goal, ct=1.0 / 0.05 / T, W=np.ones(7))[0]
        assert record("cfg3_update_cost_rel", abs(cost[b] - ref) / max(1.0, ref)) <= TOL


def test_full_size_cfg4_humanoid(oracle_lib):
    orc = _oracle(oracle_lib, "valkyrie_sim", "whole_body")
    maps = [{"type": "EffFrame", "Type": "Matrix", "EndEffector": [{"Link": "leftFoot"}, {"Link": "rightFoot"}, {"Link": "leftPalm"}, {"Link": "rightPalm"}]}]
    orc.set_problem(maps)
    tree, prob = gpu_problem_from_maps(orc, maps)
    assert prob.program_info()["joint_slots"] > 8  # large tree: the two-pass program is chosen automatically
    E = 1 << 18
    X = random_states(orc, E, seed=1238)
    X[:, 3:6] = np.random.default_rng(1239).uniform(-np.pi, np.pi, size=(E, 3))
    phi, jac = run_gpu(prob, X)
    J = jac.reshape(E, 4, 6, 38)
    # floating base: the three translations move every end-effector rigidly (identity linear block, zero angular block)
    assert np.array_equal(J[:, :, 0:3, 0:3], np.broadcast_to(np.eye(3), (E, 4, 3, 3))) and not J[:, :, 3:6, 0:3].any()
    # structural zeros: leg joints (6..17) never move the palms, arm / neck joints never move the feet
    assert not J[:, 2:4, :, 6:18].any() and not J[:, 0:2, :, 18:38].any()
    # left-leg joints (6..11) do not move the right foot and vice versa
    assert not J[:, 1, :, 6:12].any() and not J[:, 0, :, 12:18].any()
    R = phi.reshape(E, 4, 12)[:, :, 3:12].reshape(E, 4, 3, 3)
    assert np.abs(np.einsum("efij,efkj->efik", R, R) - np.eye(3)).max() < 1e-11
    sel = np.random.default_rng(4).integers(0, E, 2048)
    phi_ref, jac_ref = orc.problem_update_batch(X[sel])
    assert max_abs_err(phi[sel], phi_ref) <= TOL and max_abs_err(jac[sel], jac_ref) <= TOL


def test_full_size_cfg5_sphere_collision(oracle_lib):
    orc = _oracle(oracle_lib, "lwr_simplified", "arm")
    rng = np.random.default_rng(5)
    eff = [{"Link": "lwr_arm_%d_link" % (1 + i % 7), "LinkOffset": [0.0, 0.0, 0.02 * (i // 7)], "Radius": 0.05, "Group": "Robot"} for i in range(32)]
    eff += [{"Link": "", "LinkOffset": list(rng.uniform([-0.8, -0.8, 0.0], [0.8, 0.8, 1.2])), "Radius": 0.05, "Group": "Obstacle"} for _ in range(32)]
    maps = [{"type": "SphereCollision", "Precision": 0.1, "EndEffector": eff}]
    orc.set_problem(maps)
    tree, prob = gpu_problem_from_maps(orc, maps)
    B, T = 4096, 100
    X = _smooth_trajectories(orc, B, T, 1240).reshape(B * T, 7)
    phi, jac = run_gpu(prob, X, T=T)
    # 1024 sigmoids per evaluation, each in (0, 1)
    assert phi.min() > 0.0 and phi.max() < 1024.0 and np.isfinite(jac).all()
    # permutation equivariance, bit for bit
    perm = np.random.default_rng(6).permutation(B * T)[: 1 << 14]
    phi_p, jac_p = run_gpu(prob, X[perm])
    assert np.array_equal(phi_p, phi[perm]) and np.array_equal(jac_p, jac[perm])
    sel = np.random.default_rng(7).integers(0, B * T, 1024)
    phi_ref, jac_ref = orc.problem_update_batch(X[sel])
    # sums of 1024 sigmoids (values up to a few hundred): the measured difference is recorded, the bar stays 1e-10 absolute
    assert record("cfg5_full_phi", max_abs_err(phi[sel], phi_ref)) <= TOL and record("cfg5_full_jac", max_abs_err(jac[sel], jac_ref)) <= TOL


# ---- task-map Hessians ---------------------------------------------------------------------------
def _gpu_hessian(prob, X, layout=_capi.LAYOUT_SOA, T=1):
    E, N = X.shape
    LP, LJ = prob.length_phi, prob.length_jacobian
    dx, dphi, djac, dh = (_capi.DeviceBuffer(max(n, 1)) for n in (E * N, E * LP, E * LJ * N, E * LJ * N * N))
    dh.fill_bytes(0xFF)
    dx.upload(X.T if layout == _capi.LAYOUT_SOA else X)
    prob.update_batch_device(dx.ptr, E, dphi.ptr, djac.ptr, T=T, layout=layout)
    prob.hessian_batch_device(dx.ptr, djac.ptr, E, dh.ptr, T=T, layout=layout)
    _capi.check(_capi.lib().exo_device_synchronize())
    if layout == _capi.LAYOUT_SOA:
        return np.ascontiguousarray(dh.download((LJ, N, N, E)).transpose(3, 0, 1, 2))
    return dh.download((E, LJ, N, N))


@pytest.mark.parametrize("layout", [_capi.LAYOUT_SOA, _capi.LAYOUT_AOS_ROWMAJOR])
def test_task_map_hessians(oracle_lib, layout):
    orc = _oracle(oracle_lib, "lwr_simplified", "arm")
    maps = [{"type": "EffFrame", "EndEffector": [{"Link": "lwr_arm_7_link", "LinkOffset": [0.1, 0, 0.05]}, {"Link": "lwr_arm_5_link", "Base": "lwr_arm_2_link"}]},
            {"type": "JointLimit", "SafePercentage": 0.2},
            {"type": "EffOrientation", "Type": "Quaternion", "EndEffector": [{"Link": "lwr_arm_6_link"}]},
            {"type": "Distance", "EndEffector": [{"Link": "lwr_arm_7_link", "BaseOffset": [0.3, 0.2, 0.4]}]},
            {"type": "EffPositionXY", "EndEffector": [{"Link": "lwr_arm_4_link"}]},
            {"type": "JointPose", "JointMap": [0, 3]},
            {"type": "EffPosition", "used": False, "EndEffector": [{"Link": "lwr_arm_3_link"}]},
            {"type": "EffPosition", "EndEffector": [{"Link": "lwr_arm_7_link"}]}]
    orc.set_problem(maps)
    tree, prob = gpu_problem_from_maps(orc, maps)
    X = random_states(orc, 300, seed=201, spread=0.1)
    H = _gpu_hessian(prob, X, layout=layout)
    for e in range(0, 300, 7):
        ref = orc.problem_hessian(X[e])
        assert np.max(np.abs(H[e] - ref)) <= 1e-10
    # ComputeH quirk: angular rows of EffFrame are strictly lower triangular (kinematic_tree.cpp:969-974)
    assert not np.triu(H[:, 3:6]).any() and np.abs(np.tril(H[:, 3:6], -1)).max() > 0.1


def test_hessian_matches_finite_differences_of_the_jacobian(oracle_lib):
    # the linear rows of ComputeH are the true second derivative of the position: check against central differences of J
    orc = _oracle(oracle_lib, "lwr_simplified", "arm")
    maps = [{"type": "EffPosition", "EndEffector": [{"Link": "lwr_arm_7_link", "LinkOffset": [0.05, 0.02, 0.1]}]}]
    orc.set_problem(maps)
    tree, prob = gpu_problem_from_maps(orc, maps)
    x = random_states(orc, 1, seed=202)[0]
    h = 1e-6
    X = np.array([x] + [x + h * np.eye(7)[i] for i in range(7)] + [x - h * np.eye(7)[i] for i in range(7)])
    phi, jac = run_gpu(prob, X)
    H = _gpu_hessian(prob, X[:1])[0]
    fd = np.stack([(jac[1 + i] - jac[8 + i]) / (2 * h) for i in range(7)], axis=1)  # [row][i][col j] = d J[row][j] / d q_i
    assert np.max(np.abs(H - fd)) < 1e-8


# ---- batched IK -------------------------------------------------------------------------------------
def test_batched_ik_matches_the_reference_solver(oracle_lib, monkeypatch):
    orc = _oracle(oracle_lib, "lwr_simplified", "arm")
    maps = [{"type": "EffFrame", "EndEffector": [{"Link": "lwr_arm_7_link"}]}, {"type": "JointLimit"}]
    orc.set_problem(maps)
    tree, prob = gpu_problem_from_maps(orc, maps)
    rng = np.random.default_rng(301)
    q_goal = rng.uniform(orc.lower, orc.upper) * 0.6
    goal = orc.zero_goal()
    goal[:6] = orc.problem_update(q_goal)[0][:6]
    rho = [1.0, 0.0]  # the JointLimit rows are listed but carry no weight: they must not change the solve
    W = np.linspace(1.0, 2.0, 7)
    prob.set_cost(rho, T=1, goal=goal, W=W)
    E = 96
    Q0 = np.clip(q_goal + rng.normal(0, 0.5, size=(E, 7)), orc.lower, orc.upper)
    # three ways through the same algorithm: (1) one launch per stage with trial / update / accept launches for the line search,
    # (default) the same with the ten trials fused in registers (exo_chain_ik_linesearch_kernel), and (EXO_IK_ONE_LAUNCH=1) the
    # whole solve of every start state in ONE launch (exo_chain_ik_solve_kernel).  All must walk the reference's path.
    prob.set_kernel(1)
    Q1, cost1, iters1, term1 = prob.ik_solve_batch(Q0, max_iterations=60)
    prob.set_kernel(0)
    monkeypatch.setenv("EXO_IK_ONE_LAUNCH", "1")
    n0 = prob.launch_count()
    Q2, cost2, iters2, term2 = prob.ik_solve_batch(Q0, max_iterations=60)
    assert prob.launch_count() - n0 <= 4  # the solve kernel + the final FK and cost launches
    monkeypatch.delenv("EXO_IK_ONE_LAUNCH")
    Q, cost, iters, term = prob.ik_solve_batch(Q0, max_iterations=60)
    for Qx, itx, tx in ((Q1, iters1, term1), (Q2, iters2, term2)):
        same = (iters == itx) & (term == tx)
        conv = same & (term == 3)  # non-converged solves amplify rounding differences over their 60 iterations
        assert same.mean() >= 0.95 and conv.sum() >= E // 2 and np.max(np.abs(Q[conv] - Qx[conv])) <= 1e-6
    n_same = 0
    for e in range(E):
        q_ref, c_ref, it_ref, t_ref = orc.ik_solve(Q0[e], rho, goal=goal, W=W, max_iterations=60)
        # the returned cost is the cost of the returned state
        assert record("ik_cost_rel", abs(cost[e] - orc.trajectory_cost(Q[e:e + 1], rho, goal=goal)[0]) / max(1.0, cost[e])) <= TOL
        if t_ref == 3:  # the reference converged (FunctionTolerance): same path, same answer
            assert term[e] == 3 and iters[e] == it_ref
            assert np.max(np.abs(Q[e] - q_ref)) <= 1e-6 and abs(cost[e] - c_ref) <= 1e-9
            n_same += 1
        else:
            assert term[e] == t_ref or abs(cost[e] - c_ref) <= 1e-3 * max(1.0, c_ref)
    assert n_same >= E // 2 and (cost[term == 3] < 1e-5).all()


def test_batched_ik_whole_body(oracle_lib):
    """IKSolver::Solve on the 38-DoF humanoid with four end-effector frames (24 task rows: the step kernel that gives every active
    configuration a warp and keeps its 24 x 24 system in shared memory) and a JointLimit map without weight, from perturbed start states: the reference's path per start state where
    it converges -- iteration counts, termination criteria, solutions.  With the JointLimit rows the system is 62 x 62: more rows than a
    warp has lanes (the factorisation's row loop wraps), close to the 64-row limit."""
    orc = _oracle(oracle_lib, "valkyrie_sim", "whole_body")
    maps = [{"type": "EffFrame", "EndEffector": [{"Link": l} for l in ("leftFoot", "rightFoot", "leftPalm", "rightPalm")]}, {"type": "JointLimit"}]
    orc.set_problem(maps)
    tree, prob = gpu_problem_from_maps(orc, maps)
    rng = np.random.default_rng(311)
    lo, hi = np.maximum(orc.lower, -0.5), np.minimum(orc.upper, 0.5)
    q_goal = rng.uniform(lo, hi) * 0.5
    goal = orc.zero_goal()
    goal[:24] = orc.problem_update(q_goal)[0][:24]
    rho = [1.0, 0.0]
    prob.set_cost(rho, T=1, goal=goal)
    E = 64
    Q0 = q_goal + rng.normal(0, 0.05, size=(E, orc.n))
    Q, cost, iters, term = prob.ik_solve_batch(Q0, max_iterations=40)
    n_same = 0
    for e in range(0, E, 4):
        q_ref, c_ref, it_ref, t_ref = orc.ik_solve(Q0[e], rho, goal=goal, max_iterations=40)
        assert record("ik_wb_cost_rel", abs(cost[e] - orc.trajectory_cost(Q[e:e + 1], rho, goal=goal)[0]) / max(1.0, cost[e])) <= TOL
        if t_ref == 3:
            assert term[e] == 3 and iters[e] == it_ref
            assert np.max(np.abs(Q[e] - q_ref)) <= 1e-6 and abs(cost[e] - c_ref) <= 1e-9
            n_same += 1
        else:
            assert term[e] == t_ref or abs(cost[e] - c_ref) <= 1e-3 * max(1.0, c_ref)
    assert n_same >= 8


def test_batched_ik_position_only_on_the_three_dof_robot(oracle_lib):
    orc = _oracle(oracle_lib, "test_robot", "arm")
    maps = [{"type": "EffPosition", "EndEffector": [{"Link": "endeff"}]}]
    orc.set_problem(maps)
    tree, prob = gpu_problem_from_maps(orc, maps)
    goal = orc.problem_update(np.array([0.2, -0.1, 0.3]))[0]
    prob.set_cost([10.0], T=1, goal=goal)
    Q0 = random_states(orc, 257, seed=302)
    Q, cost, iters, term = prob.ik_solve_batch(Q0, tolerance=1e-10)
    for e in range(0, 257, 8):
        q_ref, c_ref, it_ref, t_ref = orc.ik_solve(Q0[e], [10.0], goal=goal, tolerance=1e-10)
        assert term[e] == t_ref and iters[e] == it_ref and np.max(np.abs(Q[e] - q_ref)) <= 1e-7
    assert np.isin(term, (0, 2, 3)).all() and np.median(cost) < 1e-6  # step / function tolerance (or the iteration limit in a local minimum)


def test_full_size_cfg2a_rpy_1m(oracle_lib):
    """The benchmark's own workload (bench.py: EffFrame(RPY), 2^20 configurations, seed 1234 + 2): sampled oracle parity at 1e-10 with
    the Euler angles compared modulo 2 pi, every kernel that can run it."""
    orc = _oracle(oracle_lib, "lwr_simplified", "arm")
    maps = [{"type": "EffFrame", "Type": "RPY", "EndEffector": [{"Link": "lwr_arm_7_link"}]}]
    orc.set_problem(maps)
    tree, prob = gpu_problem_from_maps(orc, maps)
    E = 1 << 20
    X = np.random.default_rng(1234 + 2).uniform(orc.lower, orc.upper, size=(E, 7))
    sel = np.random.default_rng(8).choice(E, size=8192, replace=False)
    phi_ref, jac_ref = orc.problem_update_batch(X[sel])
    for choice in (0, 3, 1):
        prob.set_kernel(choice)
        phi, jac = run_gpu(prob, X)
        e_phi, e_jac = max_abs_err(phi[sel], phi_ref, np.arange(3, 6)), max_abs_err(jac[sel], jac_ref)
        assert record("cfg2a_rpy_1m_phi_%s" % prob.last_kernel(), e_phi) <= TOL and record("cfg2a_rpy_1m_jac_%s" % prob.last_kernel(), e_jac) <= TOL
        # pitch stays inside [-pi/2, pi/2], roll / yaw inside (-pi, pi]: KDL's GetRPY ranges
        assert np.abs(phi[:, 4]).max() <= np.pi / 2 and np.abs(phi[:, 3]).max() <= np.pi and np.abs(phi[:, 5]).max() <= np.pi


@pytest.mark.parametrize("kernel", [0, 1])
def test_batched_ik_single_iteration_modes(oracle_lib, kernel, monkeypatch):
    """MaxIterations == 1, the reference's interactive configuration (example_ik.xml:5): with MaxStep the step is clamped per joint
    and taken without a line search (ik_solver.cpp:139-143), without it lambda adapts to the accepted step length and reaching
    MaximumRegularization ends the solve with Divergence (:187-199).  example_ik.xml's own shape: link 6 + quaternion offset, W."""
    orc = _oracle(oracle_lib, "lwr_simplified", "arm")
    off = [0, 0, 0, 0.7071067811865476, -4.3297802811774664e-17, 0.7071067811865475, 4.3297802811774664e-17]
    maps = [{"type": "EffFrame", "EndEffector": [{"Link": "lwr_arm_6_link", "LinkOffset": off}]}]
    orc.set_problem(maps)
    tree, prob = gpu_problem_from_maps(orc, maps)
    prob.set_kernel(kernel)
    rng = np.random.default_rng(401)
    goal = np.array([0.5, 0.0, 0.5, 0.0, 0.0, 0.0])
    W = np.array([7.0, 6, 5, 4, 3, 2, 1])
    prob.set_cost([1.0], T=1, goal=goal, W=W)
    E = 64
    Q0 = random_states(orc, E, seed=402) * 0.5
    cases = [dict(max_step=0.1), dict(max_step=1e-3), dict(), dict(regularization_rate=1e8), dict(maximum_regularization=1e-8),
             dict(threshold_increase=1.0), dict(max_step=0.05, tolerance=1e3),
             dict(maximum_regularization=5e-9, threshold_increase=1.0, threshold_decrease=2.0)]  # the last one ends in Divergence
    for one_launch in (False, True):
        if one_launch:
            if kernel != 0:
                continue
            monkeypatch.setenv("EXO_IK_ONE_LAUNCH", "1")
        for kw in cases:
            Q, cost, iters, term = prob.ik_solve_batch(Q0, max_iterations=1, **kw)
            for e in range(E):
                q_ref, c_ref, it_ref, t_ref = orc.ik_solve(Q0[e], [1.0], goal=goal, W=W, max_iterations=1, **kw)
                assert term[e] == t_ref and iters[e] == it_ref, (kw, e, term[e], t_ref, iters[e], it_ref)
                # lambda = 1e-9 leaves the 6 x 6 system ill-conditioned near singular arms (cond ~ |A| / lambda): such steps are hundreds
                # of radians long and amplify the last bits of J.  1e-10 absolute where the step is sane (<= 1 rad per joint), 1e-8 of the
                # step length otherwise.
                step = np.abs(q_ref - Q0[e]).max()
                err = np.max(np.abs(Q[e] - q_ref))
                if step <= 1.0:
                    assert record("ik_single_iteration_q", err) <= TOL, kw
                else:
                    assert record("ik_single_iteration_q_rel_long_steps", err / step) <= 1e-8, kw
                if step <= 1.0:  # (the cost at a state thousands of radians away magnifies the last bits of q)
                    assert abs(cost[e] - c_ref) <= TOL * max(1.0, c_ref)
            if kw.get("max_step"):
                assert np.max(np.abs(Q - Q0)) <= kw["max_step"] * (1 + 1e-12)  # no joint moves further than MaxStep
    # a solver that keeps calling Solve with MaxIterations = 1 (interactive IK) converges like the multi-iteration solver does
    monkeypatch.delenv("EXO_IK_ONE_LAUNCH", raising=False)
    Q = Q0.copy()
    for _ in range(30):
        Q, cost, iters, term = prob.ik_solve_batch(Q, max_iterations=1)
    assert np.median(cost) < 1e-4

"""CPU-only checks of the C-ABI library: it loads, exports every symbol include/*.h declares, and its host-side logic
(problem compilation, TaskMap indexing, Denavit-Hartenberg lowering, error reporting) behaves like the reference's
PlanningProblem::InstantiateBase (exotica_core/src/planning_problem.cpp:148-244).  No kernel is launched here."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, robot_paths
from exotica_b200 import _capi
from helpers import gpu_problem_from_maps


def _declared(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(exo_[a-z0-9_]+)\s*\(", text)))


@pytest.mark.parametrize("header", ["exotica_b200.h", "exotica_b200_host.h"])
def test_library_exports_every_declared_symbol(header):
    L = ctypes.CDLL(_capi.LIB_PATH)
    names = _declared(header)
    assert len(names) > 20
    for name in names:
        assert hasattr(L, name), "%s declared in include/%s but not exported" % (name, header)


def test_binding_table_matches_header():
    assert sorted(n for n, _, _ in _capi.SYMBOLS) == _declared("exotica_b200.h")


def _oracle(oracle_lib, robot, group):
    return oracle_lib.Oracle(oracle_lib.load_robot(*robot_paths(robot), joint_group=group))


def test_indexing_follows_map_order(oracle_lib):
    orc = _oracle(oracle_lib, "lwr_simplified", "arm")
    maps = [{"type": "EffFrame", "Type": "Quaternion", "EndEffector": [{"Link": "lwr_arm_7_link"}, {"Link": "lwr_arm_3_link"}]},
            {"type": "JointLimit"},
            {"type": "EffPosition", "EndEffector": [{"Link": "lwr_arm_5_link"}]}]
    orc.set_problem(maps)
    tree, prob = gpu_problem_from_maps(orc, maps)
    assert (prob.length_phi, prob.length_jacobian) == (14 + 7 + 3, 12 + 7 + 3) == (orc.length_phi, orc.length_jacobian)
    assert [prob.map_indexing(i) for i in range(3)] == [(0, 14, 0, 12), (14, 7, 12, 7), (21, 3, 19, 3)]


@pytest.mark.parametrize("robot,group,link,joints", [("lwr_simplified", "arm", "lwr_arm_7_link", 7), ("lwr_simplified", "arm", "lwr_arm_6_link", 6),
                                                     ("test_robot", "arm", "endeff", 3)])
def test_serial_chains_lower_to_dh(oracle_lib, robot, group, link, joints):
    orc = _oracle(oracle_lib, robot, group)
    maps = [{"type": "EffFrame", "EndEffector": [{"Link": link, "LinkOffset": [0.01, 0.02, 0.03, 0.1, 0.2, 0.3]}]}, {"type": "JointLimit"}]
    orc.set_problem(maps)
    tree, prob = gpu_problem_from_maps(orc, maps)
    info = prob.chain_info()
    # the LWR's consecutive axes are perpendicular and intersect (alpha = +-90 degrees, a = 0 exactly): the signed-swap link form
    # applies; the test robot has two parallel axes in a row
    assert info == {"lowered": True, "joints": joints, "all_revolute": True, "orthogonal_links": robot == "lwr_simplified", "frames": 1}


def test_trees_and_relative_frames_stay_on_the_interpreter(oracle_lib):
    orc = _oracle(oracle_lib, "valkyrie_sim", "whole_body")
    maps = [{"type": "EffFrame", "EndEffector": [{"Link": "leftFoot"}, {"Link": "rightPalm"}]}]
    orc.set_problem(maps)
    tree, prob = gpu_problem_from_maps(orc, maps)
    assert not prob.chain_info()["lowered"]
    orc = _oracle(oracle_lib, "lwr_simplified", "arm")
    maps = [{"type": "EffFrame", "EndEffector": [{"Link": "lwr_arm_7_link", "Base": "lwr_arm_2_link"}]}]
    orc.set_problem(maps)
    tree, prob = gpu_problem_from_maps(orc, maps)
    assert not prob.chain_info()["lowered"]


def test_errors_use_the_status_and_message_channel(oracle_lib):
    orc = _oracle(oracle_lib, "lwr_simplified", "arm")
    tree = _capi.Tree(orc.tree.flat(), orc.tree.n_controlled, orc.lower, orc.upper, orc.tree.tree_state)
    prob = _capi.Problem(tree)
    with pytest.raises(_capi.ExoticaError, match="Unknown rotation representation"):
        prob.add_task_map(_capi.MAP_EFF_FRAME, elem_a=[3], rotation_type=17)
    with pytest.raises(_capi.ExoticaError, match="No frame_A link"):
        prob.add_task_map(_capi.MAP_EFF_POSITION, elem_a=[999])
    with pytest.raises(_capi.ExoticaError, match="radius"):
        prob.add_task_map(_capi.MAP_SPHERE_COLLISION, elem_a=[3], precision=0.1)
    prob.add_task_map(_capi.MAP_EFF_POSITION, elem_a=[3])
    with pytest.raises(_capi.ExoticaError, match="not finalized"):
        prob.map_indexing(0)
    prob.finalize()
    with pytest.raises(_capi.ExoticaError, match="already finalized"):
        prob.finalize()
    x = np.zeros(7)
    rc = _capi.lib().exo_problem_update_batch(prob.h, x.ctypes.data, -1, 1, 0, x.ctypes.data, None, None)
    assert rc != 0 and b"negative" in _capi.lib().exo_last_error()
    # the solver-facing entry points need a cost first (all of this is checked before any CUDA call)
    p = x.ctypes.data
    L = _capi.lib()
    for rc in (L.exo_problem_cost_batch(prob.h, p, p, 1, 0, p, None, None),
               L.exo_problem_cost_jacobian_batch(prob.h, p, p, p, 1, 0, p, None),
               L.exo_problem_cost_gradient_batch(prob.h, p, 1, 0, None, p, None),
               L.exo_problem_update_cost_batch(prob.h, p, 1, 0, p, p, p, None, None),
               L.exo_problem_aico_task_messages_batch(prob.h, p, p, p, 1, 0, p, p, p, p, None)):
        assert rc != 0 and b"exo_problem_set_cost has not been called" in L.exo_last_error()
    with pytest.raises(_capi.ExoticaError, match="CenterOfMass needs mass"):
        prob2 = _capi.Problem(tree)
        prob2.add_task_map(_capi.MAP_CENTER_OF_MASS, elem_a=[3])
    with pytest.raises(_capi.ExoticaError, match="kernel choice"):
        prob.set_kernel(10)


def test_trajectory_matches_the_independent_restatement():
    """exotica::Trajectory (host C++, exotica_b200/csrc/host/trajectory.cpp) against oracle/trajectory.py, plus the properties KDL's
    line path / single-axis rotation / cubic profile guarantee: knots are hit exactly, stationary segments hold, the end frame
    persists after the duration, the start frame before t = 0, and the rotation stays about one axis."""
    import exotica_b200 as exo
    from oracle.trajectory import Trajectory as RefTrajectory
    from oracle.oracle import get_frame

    text = "0.5\n5 7\n" + "\n".join([
        "0.0  0.5 0.1 0.6   0 0 0",
        "1.0  0.5 0.1 0.6   0 0 0",          # stationary
        "2.5  0.3 0.4 0.7   0.2 -0.1 0.4",   # translation-limited (radius * angle < distance)
        "3.0  0.3 0.4 0.7   1.2 0.3 -0.5",   # pure rotation (rotation-limited)
        "4.0  0.31 0.4 0.7  -1.0 0.8 0.9"])  # rotation-limited with a small translation
    traj = exo.Trajectory(text)
    ref = RefTrajectory(text)
    assert abs(traj.get_duration() - 4.0) < 1e-15 and abs(ref.get_duration() - 4.0) < 1e-15
    times = np.concatenate([np.linspace(-0.5, 4.5, 201), [0.0, 1.0, 2.5, 3.0, 4.0]])
    got = traj.get_position(times)
    want = np.array([ref.get_position(t) for t in times])
    assert np.max(np.abs(got - want)) <= 1e-12
    knots = {0.0: [0.5, 0.1, 0.6, 0, 0, 0], 1.0: [0.5, 0.1, 0.6, 0, 0, 0], 2.5: [0.3, 0.4, 0.7, 0.2, -0.1, 0.4], 3.0: [0.3, 0.4, 0.7, 1.2, 0.3, -0.5]}
    for t, v in knots.items():
        assert np.max(np.abs(traj.get_position(t) - np.asarray(get_frame(v)))) <= 1e-12
    assert np.max(np.abs(traj.get_position(9.0) - np.asarray(get_frame([0.31, 0.4, 0.7, -1.0, 0.8, 0.9])))) <= 1e-12
    assert np.array_equal(traj.get_position(-3.0), traj.get_position(0.0)) and np.array_equal(traj.get_position(0.4), traj.get_position(0.9))
    R = got[:, :9].reshape(-1, 3, 3)
    assert np.max(np.abs(np.einsum("eij,ekj->eik", R, R) - np.eye(3))) <= 1e-12
    # matrix form (data, radius) and the reference's error messages
    data = np.array([[0, 0, 0, 0], [2, 1, 0, 0]], dtype=np.float64)
    mid = exo.Trajectory(data, radius=1.0).get_position(1.0)
    assert np.max(np.abs(mid[9:] - [0.5, 0, 0])) <= 1e-15  # the cubic profile is symmetric
    with pytest.raises(exo.ExoticaError, match="monotonically increasing"):
        exo.Trajectory("1\n2 4\n0 0 0 0\n0 1 0 0").get_position(0.0)
    with pytest.raises(exo.ExoticaError, match="Invalid trajectory data size"):
        exo.Trajectory("1\n2 5\n0 0 0 0 0\n1 1 0 0 0").get_position(0.0)


def test_generated_tree_kernel_source_compiles_without_a_gpu():
    """tree_codegen.cpp unrolls the 38-DoF humanoid's two-pass program into CUDA C++ and NVRTC compiles it for sm_100a (neither
    needs a device); serial chains of SphereCollision maps are outside the generator's scope and say so."""
    from oracle import oracle
    urdf, srdf = robot_paths("valkyrie_sim")
    orc = oracle.Oracle(oracle.load_robot(urdf, srdf, "whole_body"))
    maps = [{"type": "EffFrame", "EndEffector": [{"Link": l} for l in ("leftFoot", "rightFoot", "leftPalm", "rightPalm")]}, {"type": "JointLimit"}]
    orc.set_problem(maps)
    gtree, prob = gpu_problem_from_maps(orc, maps)
    src = prob.generated_kernel_source(compile=True)
    # 6 + 6 + 6 + 3 + 7 + 7 joints feed the four frames, three of them the sliding joints of the floating base; two passes
    assert src.count("SINCOS(q, sn, cs);") == 2 * 32 and src.count(": JointLimit\n") == 1
    assert "cp.async.bulk.tensor.3d.global.shared::cta" in src
    urdf, srdf = robot_paths("lwr_simplified")
    orc = oracle.Oracle(oracle.load_robot(urdf, srdf, "arm"))
    maps = [{"type": "SphereCollision", "Precision": 0.1, "EndEffector": [{"Link": "lwr_arm_7_link", "Radius": 0.1, "Group": "a"},
                                                                       {"Link": "", "LinkOffset": [0.5, 0, 0.5], "Radius": 0.1, "Group": "b"}]}]
    orc.set_problem(maps)
    gtree, prob = gpu_problem_from_maps(orc, maps)
    with pytest.raises(_capi.ExoticaError, match="generated tree kernel not applicable"):
        prob.generated_kernel_source()


def build_c_consumer(tmp_path):
    """gcc -std=c99 on tests/c_consumer/consumer.c against include/*.h and the in-tree shared library; returns the binary's path."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib_dir = os.path.join(root, "exotica_b200", "lib")
    exe = str(tmp_path / "consumer")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-O2", "-I" + os.path.join(root, "include"),
                           os.path.join(root, "tests", "c_consumer", "consumer.c"), "-o", exe, "-L" + lib_dir, "-lexotica_b200", "-Wl,-rpath," + lib_dir])
    return exe


def test_plain_c_consumer_of_the_boundary(tmp_path):
    """The drop-in boundary is a C ABI: a C99 program with nothing but the two headers and the .so loads the reference's example_ik.xml
    (Setup::LoadProblem), reads dimensions, task indexing and joint limits -- the host logic, no CUDA call (n_evals = 0)."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    examples = os.path.join(root, "tests", "fixtures", "exotica_examples")
    exe = build_c_consumer(tmp_path)
    out = subprocess.check_output([exe, "exotica_examples", examples, os.path.join(examples, "resources", "configs", "example_ik.xml"), "0"], text=True).splitlines()
    assert out[0] == "type exotica/UnconstrainedEndPoseProblem N 7 length_Phi 6 length_jacobian 6 T 1"
    assert out[1] == "task Position id 0 start 0 length 6 startJ 0 lengthJ 6 used 1"
    limits = np.array(out[2].split()[1:], dtype=float).reshape(7, 2)
    assert np.all(limits[:, 0] < 0) and np.allclose(limits[:, 0], -limits[:, 1])
    # errors reach the C caller through the return code and exo_last_error
    r = subprocess.run([exe, "exotica_examples", examples, os.path.join(examples, "resources", "configs", "example_ik.xml").replace("example_ik", "nope"), "0"],
                       capture_output=True, text=True)
    assert r.returncode == 2 and "cannot open" in r.stderr

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
    assert info == {"lowered": True, "joints": joints, "all_revolute": True, "frames": 1}


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

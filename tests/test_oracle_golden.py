"""Pins the CPU oracle against every known-answer vector the reference's own tests hold for this path
(SURVEY.md section 8c):

  * EffPosition goldens      exotations/exotica_core_task_maps/test/test_maps.cpp:340-347 (eps 1e-5)
  * SphereCollision goldens  exotations/exotica_core_task_maps/test/test_maps.cpp:651-673 (eps 1e-5)
  * Valkyrie centre of mass  exotica_examples/test/test_valkyrie_com:13-120 (Drake, assert_allclose rtol 1e-7)

The golden numbers below are the reference's test data (not code); robots are the kinematics-only fixtures
written by tests/golden/make_robot_fixtures.py.
"""
import numpy as np
import pytest

from conftest import robot_paths


@pytest.fixture(scope="module")
def test_robot(oracle_lib):
    tree = oracle_lib.load_robot(*robot_paths("test_robot"), joint_group="arm")
    return oracle_lib.Oracle(tree)


def test_test_robot_structure(test_robot):
    t = test_robot.tree
    assert t.model_link_names == ["world_frame", "base", "link1", "link2", "link3", "endeff"]
    assert [e.id for e in t.elements] == [-1, 1, 2, 3, 4, 5]
    assert t.controlled_joint_names == ["joint1", "joint2", "joint3"]
    assert [e.control_id for e in t.elements] == [-1, -1, 0, 1, 2, -1]
    np.testing.assert_allclose(t.lower, -0.4)
    np.testing.assert_allclose(t.upper, 0.4)


def test_eff_position_goldens(test_robot):
    # test_maps.cpp:340-347
    X = np.array([0.680375, -0.211234, 0.566198, 0.59688, 0.823295, -0.604897, -0.329554, 0.536459, -0.444451, 0.10794,
                  -0.0452059, 0.257742, -0.270431, 0.0268018, 0.904459]).reshape(5, 3)
    Y = np.array([0.0645323, 0.0522249, 1.21417, 0.292945, 0.199075, 1.12724, 0.208378, -0.0712708, 1.19893, 0.0786457,
                  0.00852213, 1.23952, 0.356984, -0.0989639, 1.06844]).reshape(5, 3)
    J = np.array([-0.0522249, 0.594015, 0.327994, 0.0645323, 0.480726, 0.26544, 0, -0.0830172, -0.156401, -0.199075,
                  0.560144, 0.363351, 0.292945, 0.380655, 0.246921, 0, -0.354186, -0.0974994, 0.0712708, 0.708627,
                  0.423983, 0.208378, -0.24237, -0.145014, 0, -0.220229, -0.0413455, -0.00852213, 0.784922, 0.437315,
                  0.0786457, 0.085055, 0.0473879, 0, -0.0791061, -0.0949228, 0.0989639, 0.595968, 0.258809, 0.356984,
                  -0.165215, -0.0717477, 0, -0.370448, -0.361068]).reshape(5, 3, 3)
    test_robot.set_problem([{"type": "EffPosition", "EndEffector": [{"Link": "endeff"}]}])
    for i in range(5):
        phi, jac = test_robot.problem_update(X[i])
        # the reference test uses ||.|| < 1e-5 on 6-significant-figure goldens
        assert np.linalg.norm(phi - Y[i]) < 1e-5
        assert np.linalg.norm(jac - J[i]) < 1e-5


def test_sphere_collision_goldens(test_robot):
    # test_maps.cpp:651-673
    X = np.array([-0.590167, 1.2309, 1.67611, -1.72098, 1.79731, 0.103981, -1.65578, -1.23114, 0.652908, 1.56093,
                  -0.604428, -1.74331, -1.91991, -0.169193, -1.74762]).reshape(5, 3)
    Y = np.array([1, 5.71023e-44, 1.83279e-110, 6.87352e-16, 7.26371e-45])
    J = np.array([0, 0.431392, 0.449344, 0, 0.431735, 0.26014, 0, -0.234475, -0.0135658, 0, -0.349195, -0.447214, 0,
                  -0.270171, -0.430172]).reshape(5, 1, 3)
    test_robot.set_problem([{"type": "SphereCollision", "Precision": 1e-2,
                             "EndEffector": [{"Link": "base", "Radius": 0.3, "Group": "base"},
                                             {"Link": "endeff", "Radius": 0.3, "Group": "eff"}]}])
    assert test_robot.length_phi == 1 and test_robot.length_jacobian == 1
    for i in range(5):
        phi, jac = test_robot.problem_update(X[i])
        assert abs(phi[0] - Y[i]) < 1e-5
        if Y[i] > 1e-300:
            # x is printed to 6 digits and phi = 1/(1+exp(500*(d-0.6))): a 1e-6 error in d is a 5e-4 relative
            # error in phi, so the tail values can only be pinned to ~1e-3 relative
            assert abs(phi[0] - Y[i]) / Y[i] < 2e-3
        assert np.linalg.norm(jac - J[i]) < 1e-5


def _com_frames(orc, links=None):
    """CenterOfMass::Initialize (center_of_mass.cpp:150-170): the frame offset is the link's centre of gravity, its weight the
    link's mass; without EndEffector entries the map walks every robot link with mass (:88-117).  The masses are the tree's:
    for a fixed base the first child of the root link carries the ROOT's inertial (kinematic_tree.cpp:264-268), so link1 of the
    test robot is massless in the reference -- the goldens below only come out with that quirk in place."""
    out = []
    for e in orc.tree.elements:
        if e.is_robot_link and e.mass > 0 and (links is None or e.link_name in links):
            out.append({"Link": e.link_name, "LinkOffset": list(e.cog), "Mass": e.mass})
    return out


def test_center_of_mass_goldens(test_robot):
    # test_maps.cpp:789-892: (a) all links, (b) EndEffector = link2, endeff, (c) all links projected on the XY plane
    cases = [
        (None, True,
         [-0.281809, 0.10497, 0.15886, -0.0948483, 0.374775, -0.80072, 0.061616, 0.514588, -0.39141, 0.984457, 0.153942, 0.755228, 0.495619,
          0.25782, -0.929158],
         [0.081112, -0.0234831, 0.924368, 0.0080578, -0.000766568, 0.895465, 0.157569, 0.00972105, 0.897157, 0.117213, 0.176455, 0.846633,
          -0.0587457, -0.0317596, 0.877501],
         [0.0234831, 0.455657, 0.200919, 0.081112, -0.131919, -0.0581688, 0, -0.0844429, -0.0565023, 0.000766568, 0.443462, 0.19642, 0.0080578,
          -0.0421882, -0.0186862, 0, -0.00809418, 0.0895227, -0.00972105, 0.446309, 0.214617, 0.157569, 0.0275346, 0.0132406, 0, -0.157868,
          -0.0266211, -0.176455, 0.219463, 0.0736575, 0.117213, 0.330384, 0.110885, 0, -0.211838, -0.170949, 0.0317596, 0.376061, 0.149235,
          -0.0587457, 0.203309, 0.0806805, 0, 0.0667812, 0.134774]),
        (("link2", "endeff"), True,
         [0.299414, -0.503912, 0.258959, -0.541726, 0.40124, -0.366266, -0.342446, -0.537144, -0.851678, 0.266144, -0.552687, 0.302264,
          0.0213719, 0.942931, -0.439916],
         [-0.167532, -0.0517162, 0.913823, 0.083533, -0.0502684, 0.931962, -0.3632, 0.129477, 0.693081, -0.17971, -0.048991, 0.907923, 0.314586,
          0.00672432, 0.823106],
         [0.0517162, 0.443188, 0.254921, -0.167532, 0.13681, 0.0786927, 0, 0.175333, 0.0666905, 0.0502684, 0.412954, 0.235481, 0.083533,
          -0.248507, -0.141708, 0, -0.0974919, -0.00961589, -0.129477, 0.228967, 0.0468775, -0.3632, -0.0816248, -0.0167114, 0, 0.385588,
          0.270459, 0.048991, 0.441801, 0.257042, -0.17971, 0.12044, 0.0700725, 0, 0.186268, 0.0681488, -0.00672432, 0.373021, 0.240882,
          0.314586, 0.00797337, 0.00514888, 0, -0.314658, -0.132569]),
        (None, False,
         [0.350952, -0.0340994, -0.0361284, -0.390089, 0.424175, -0.634888, 0.243646, -0.918271, -0.172033, 0.391968, 0.347873, 0.27528,
          -0.305768, -0.630755, 0.218212],
         [-0.0228141, -0.0083524, 0.0595937, -0.0245025, -0.392081, -0.0974653, 0.200868, 0.0830305, -0.232814, 0.0734919],
         [0.0083524, 0.453225, 0.202958, -0.0228141, 0.165929, 0.0743046, 0.0245025, 0.420734, 0.195957, 0.0595937, -0.172988, -0.0805696,
          0.0974653, 0.254325, 0.0971889, -0.392081, 0.0632213, 0.0241597, -0.0830305, 0.394279, 0.162599, 0.200868, 0.162978, 0.0672114,
          -0.0734919, 0.394649, 0.189283, -0.232814, -0.124578, -0.0597504]),
    ]
    for links, enable_z, X, Y, J in cases:
        dim = 3 if enable_z else 2
        X, Y, J = np.array(X).reshape(5, 3), np.array(Y).reshape(5, dim), np.array(J).reshape(5, dim, 3)
        test_robot.set_problem([{"type": "CenterOfMass", "EnableZ": enable_z, "EndEffector": _com_frames(test_robot, links)}])
        assert test_robot.length_phi == dim and test_robot.length_jacobian == dim
        for i in range(5):
            phi, jac = test_robot.problem_update(X[i])
            assert np.linalg.norm(phi - Y[i]) < 1e-5
            assert np.linalg.norm(jac - J[i]) < 1e-5


VALKYRIE_COM = [
    # (state builder, Drake centre of mass) -- exotica_examples/test/test_valkyrie_com:13-120
    (lambda v: v, [-0.007956278619357, -0.000791781655122, -0.023887289533801]),
    (lambda v: _set(v, {0: 1, 1: 2, 2: 3}), [0.992043721380643, 1.999208218344877, 2.976112710466198]),
    (lambda v: v + 1.0, [1.011827488179183, 0.984257082034692, 0.909060824252459]),
    (lambda v: _set(v + 1.0, {3: 0, 4: 0, 5: 0}), [1.072817938570169, 0.943826381822213, 0.985867234115696]),
    (lambda v: _set(v, {0: 1, 1: 1, 2: 1, 3: 1.57}), [0.992043721380643, 1.023886651443021, 0.999189196509224]),
    (lambda v: _set(v, {0: 1, 1: 1, 2: 1, 3: 1.57 / 3}), [0.992043721380643, 1.011252345052587, 0.978914122017840]),
    (lambda v: _set(v, {0: 1, 1: 1, 2: 1, 4: 1.57}), [0.976106382242915, 0.999208218344877, 1.007937254009971]),
    (lambda v: _set(v, {0: 1, 1: 1, 2: 1, 5: 1.57}), [1.000785445606890, 0.992043093386445, 0.976112710466199]),
    (lambda v: _set(v, {0: 1, 1: 1, 2: 1, 3: 1.57, 4: 1.57}), [0.999182860969121, 1.023886651443021, 1.007955630432197]),
    (lambda v: _set(v, {3: 1.57 / 4, 4: -1.57 / 4, 5: 1.57 / 4}), [-0.002100122545919, 0.008227677204985, -0.023715537144087]),
]


def _set(v, d):
    v = v.copy()
    for k, val in d.items():
        v[k] = val
    return v


@pytest.fixture(scope="module")
def valkyrie(oracle_lib):
    tree = oracle_lib.load_robot(*robot_paths("valkyrie_sim"), joint_group="whole_body")
    return oracle_lib.Oracle(tree)


def test_valkyrie_structure(valkyrie):
    t = valkyrie.tree
    assert t.n_controlled == 38
    assert t.base_type == "FLOATING" and t.controlled_base_type == "FLOATING"
    assert t.controlled_joint_names[:6] == ["world_joint/trans_x", "world_joint/trans_y", "world_joint/trans_z",
                                            "world_joint/rot_x", "world_joint/rot_y", "world_joint/rot_z"]
    # state order pinned by the SRDF whole_body listing (valkyrie_sim.srdf:37-71) == MoveIt DFS order
    assert t.controlled_joint_names[6:12] == ["leftHipYaw", "leftHipRoll", "leftHipPitch", "leftKneePitch",
                                              "leftAnklePitch", "leftAnkleRoll"]
    assert t.controlled_joint_names[12] == "rightHipYaw"
    assert t.controlled_joint_names[18:21] == ["torsoYaw", "torsoPitch", "torsoRoll"]
    assert t.controlled_joint_names[21] == "leftShoulderPitch"
    assert t.controlled_joint_names[28:31] == ["lowerNeckPitch", "neckYaw", "upperNeckPitch"]
    assert t.controlled_joint_names[31] == "rightShoulderPitch"
    # virtual chain: TransX,TransY,TransZ,RotZ,RotY,RotX; roll (state 3) drives the deepest element
    assert [e.id for e in t.elements[:8]] == [-1, 0, 1, 2, 3, 4, 5, 7]
    assert [e.control_id for e in t.elements[1:7]] == [0, 1, 2, 5, 4, 3]
    assert t.elements[7].link_name == "pelvis"
    assert t.elements[6].link_name == "world_frame/rot_x"
    np.testing.assert_allclose(t.lower[3:6], -np.pi)
    assert t.lower[0] == -1e20 and t.upper[2] == 1e20


@pytest.mark.parametrize("case", range(len(VALKYRIE_COM)))
def test_valkyrie_com_goldens(valkyrie, case):
    """CenterOfMass::Update (center_of_mass.cpp:37-54): sum_i mass_i * FK(link_i, cog_i).p / M over all
    links with mass (center_of_mass.cpp:186-205 initialises one frame per massive robot link)."""
    t = valkyrie.tree
    frames, masses = [], []
    for e in t.elements:
        if e.is_robot_link and e.mass > 0:
            frames.append({"Link": e.link_name, "LinkOffset": e.cog})
            masses.append(e.mass)
    masses = np.array(masses)
    valkyrie.request_frames(frames)
    build, golden = VALKYRIE_COM[case]
    x = build(np.zeros(38))
    valkyrie.update(x)
    p = valkyrie.phi()[:, 9:12]
    com = (p * masses[:, None]).sum(axis=0) / masses.sum()
    np.testing.assert_allclose(com, np.array(golden), rtol=1e-7)
    # Jacobian of the CoM vs finite differences (pins the floating-base column convention)
    J = valkyrie.jacobian()[:, :3, :]
    Jcom = (J * masses[:, None, None]).sum(axis=0) / masses.sum()
    h = 1e-6
    for c in (0, 3, 4, 5, 8, 20, 25, 33):
        xp, xm = x.copy(), x.copy()
        xp[c] += h
        xm[c] -= h
        valkyrie.update(xp)
        cp = (valkyrie.phi()[:, 9:12] * masses[:, None]).sum(axis=0) / masses.sum()
        valkyrie.update(xm)
        cm = (valkyrie.phi()[:, 9:12] * masses[:, None]).sum(axis=0) / masses.sum()
        np.testing.assert_allclose(Jcom[:, c], (cp - cm) / (2 * h), atol=1e-7)


def test_lwr_structure_and_zero_pose(oracle_lib):
    tree = oracle_lib.load_robot(*robot_paths("lwr_simplified"), joint_group="arm")
    assert tree.model_link_names == ["world_frame", "base"] + ["lwr_arm_%d_link" % i for i in range(8)]
    assert [e.id for e in tree.elements] == [-1] + list(range(1, 10))
    assert tree.controlled_joint_names == ["lwr_arm_%d_joint" % i for i in range(7)]
    assert tree.controlled_link_names == ["lwr_arm_%d_link" % i for i in range(1, 8)]
    assert [e.joint_name for e in tree.elements[:3]] == ["world_joint", "NoName", "base_joint"]
    # MoveIt prefers the safety_controller soft limits (unpinned by the reference)
    np.testing.assert_allclose(tree.upper, [2.93215314335, 2.05948851735] * 3 + [2.93215314335])
    o = oracle_lib.Oracle(tree)
    o.request_frames([{"Link": "lwr_arm_7_link"}])
    o.update(np.zeros(7))
    phi = o.phi()[0]
    # translation (0,0,1.1985), rotation Rz(3.14159265359)
    np.testing.assert_allclose(phi[9:], [0, 0, 1.1985], atol=1e-15)
    rpy = oracle_lib.set_rotation(phi[:9], "RPY")
    assert abs(rpy[2] - (-3.1415926535895866)) < 1e-15  # App. C #11: yaw sits on the -pi side of the cut


def test_lwr_collision_elements(oracle_lib):
    tree = oracle_lib.load_robot(*robot_paths("lwr_simplified"), joint_group="arm", add_collision_elements=True)
    # 10 model elements + 16 collision elements
    assert len(tree.elements) == 26

"""URDF/SRDF -> EXOTica element tree, restated for the CPU oracle.  TEST INFRASTRUCTURE ONLY.

Restates, in plain Python, what the reference does at initialisation time
(exotica_core/src/kinematic_tree.cpp:101-162 Instantiate, :164-421 BuildTree, :580-598
AddElementFromSegmentMapIterator / IsControlled, :1217-1297 joint limits; scene.cpp:795-821 collision
elements) together with the behaviour of the external libraries it calls:

* urdfdom: links_/joints_ are std::map (sorted by name); child_links of a link are appended while
  iterating joints_ => siblings ordered by ascending JOINT name; rpy -> quaternion via setFromRPY.
* kdl_parser::treeFromUrdfModel: revolute/continuous -> Joint(name, F.p, F.M*axis, RotAxis), prismatic ->
  TransAxis, fixed -> None; Segment(child, joint, F) so f_tip = joint.pose(0)^-1 * F.
* MoveIt RobotModel: joints in DFS pre-order over child_links; variable names; group semantics
  (chain / joint / link / subgroup, sorted by joint index); bounds prefer safety_controller soft limits.

This module is deliberately independent of the product's C++ loader (exotica_b200/csrc/host) so that
tests can cross-check element order, ids and control ids between the two implementations.
"""
import math
import xml.etree.ElementTree as ET

import numpy as np

# KDL::Joint::JointType order (matches exo_oracle.h)
ROT_AXIS, ROT_X, ROT_Y, ROT_Z, TRANS_AXIS, TRANS_X, TRANS_Y, TRANS_Z, NONE = range(9)

INF = 1e20  # exotica "inf" (kinematic_tree.h:70)


def _floats(s, n, default):
    if s is None:
        return list(default)
    v = [float(t) for t in s.split()]
    if len(v) != n:
        raise ValueError("expected %d numbers in %r" % (n, s))
    return v


def quaternion_from_rpy(roll, pitch, yaw):
    """urdf::Rotation::setFromRPY followed by normalize()."""
    phi, the, psi = roll / 2.0, pitch / 2.0, yaw / 2.0
    x = math.sin(phi) * math.cos(the) * math.cos(psi) - math.cos(phi) * math.sin(the) * math.sin(psi)
    y = math.cos(phi) * math.sin(the) * math.cos(psi) + math.sin(phi) * math.cos(the) * math.sin(psi)
    z = math.cos(phi) * math.cos(the) * math.sin(psi) - math.sin(phi) * math.sin(the) * math.cos(psi)
    w = math.cos(phi) * math.cos(the) * math.cos(psi) + math.sin(phi) * math.sin(the) * math.sin(psi)
    s = math.sqrt(x * x + y * y + z * z + w * w)
    if s == 0.0:
        return 0.0, 0.0, 0.0, 1.0
    return x / s, y / s, z / s, w / s


def rotation_from_quaternion(x, y, z, w):
    """KDL::Rotation::Quaternion, row-major 9."""
    x2, y2, z2, w2 = x * x, y * y, z * z, w * w
    return [w2 + x2 - y2 - z2, 2 * x * y - 2 * w * z, 2 * x * z + 2 * w * y,
            2 * x * y + 2 * w * z, w2 - x2 + y2 - z2, 2 * y * z - 2 * w * x,
            2 * x * z - 2 * w * y, 2 * y * z + 2 * w * x, w2 - x2 - y2 + z2]


def frame_from_origin(xyz, rpy):
    """kdl_parser toKdl(urdf::Pose): 12 doubles (row-major R, p)."""
    q = quaternion_from_rpy(*rpy)
    return rotation_from_quaternion(*q) + list(xyz)


IDENTITY = [1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0]


class UrdfJoint(object):
    def __init__(self, e):
        self.name = e.attrib["name"]
        self.type = e.attrib["type"]
        origin = e.find("origin")
        self.xyz = _floats(origin.attrib.get("xyz") if origin is not None else None, 3, (0, 0, 0))
        self.rpy = _floats(origin.attrib.get("rpy") if origin is not None else None, 3, (0, 0, 0))
        axis = e.find("axis")
        self.axis = _floats(axis.attrib.get("xyz") if axis is not None else None, 3, (1, 0, 0))
        self.parent = e.find("parent").attrib["link"]
        self.child = e.find("child").attrib["link"]
        lim = e.find("limit")
        self.limits = None
        if lim is not None:
            self.limits = dict(lower=float(lim.attrib.get("lower", 0.0)), upper=float(lim.attrib.get("upper", 0.0)),
                               velocity=float(lim.attrib.get("velocity", 0.0)))
        saf = e.find("safety_controller")
        self.safety = None
        if saf is not None:
            self.safety = dict(lower=float(saf.attrib.get("soft_lower_limit", 0.0)),
                               upper=float(saf.attrib.get("soft_upper_limit", 0.0)))
        mim = e.find("mimic")
        self.mimic = None
        if mim is not None:
            self.mimic = dict(joint=mim.attrib["joint"], multiplier=float(mim.attrib.get("multiplier", 1.0)),
                              offset=float(mim.attrib.get("offset", 0.0)))


class UrdfLink(object):
    def __init__(self, e):
        self.name = e.attrib["name"]
        self.mass = 0.0
        self.cog = [0.0, 0.0, 0.0]
        inertial = e.find("inertial")
        if inertial is not None:
            m = inertial.find("mass")
            if m is not None:
                self.mass = float(m.attrib["value"])
            o = inertial.find("origin")
            if o is not None:
                self.cog = _floats(o.attrib.get("xyz"), 3, (0, 0, 0))
        self.collisions = []
        for c in e.findall("collision"):
            o = c.find("origin")
            xyz = _floats(o.attrib.get("xyz") if o is not None else None, 3, (0, 0, 0))
            rpy = _floats(o.attrib.get("rpy") if o is not None else None, 3, (0, 0, 0))
            self.collisions.append(frame_from_origin(xyz, rpy))
        self.parent_joint = None
        self.child_links = []


class Element(object):
    """One exotica::KinematicElement (kinematic_element.h:54-143), flattened."""

    def __init__(self, id_, parent, link_name, joint_name, joint_type, origin=(0, 0, 0), axis=(0, 0, 0), f_tip=None):
        self.id = id_
        self.parent = parent  # element index or -1
        self.link_name = link_name
        self.joint_name = joint_name
        self.joint_type = joint_type
        self.origin = list(origin)
        self.axis = list(axis)
        self.f_tip = list(f_tip) if f_tip is not None else list(IDENTITY)
        self.control_id = -1
        self.mimic_elem = -1
        self.mimic_mult = 1.0
        self.mimic_off = 0.0
        self.generated = -1
        self.mass = 0.0
        self.cog = [0.0, 0.0, 0.0]
        self.is_robot_link = False


class RobotTree(object):
    """Flat element tree in model_tree_ order + name lists + limits (the oracle's view of a Scene)."""

    def __init__(self, urdf_text, srdf_text, joint_group="", add_collision_elements=False):
        self._parse_urdf(urdf_text)
        self._parse_srdf(srdf_text)
        self._build_moveit_model()
        self._instantiate(joint_group)
        self._build_tree()
        if add_collision_elements:
            self._add_collision_elements()
        self.tree_state = np.zeros(len(self.elements))

    # -- urdfdom ----------------------------------------------------------------------------
    def _parse_urdf(self, text):
        root = ET.fromstring(text)
        self.links = {}
        for e in root.findall("link"):
            link = UrdfLink(e)
            self.links[link.name] = link
        self.joints = {}
        for e in root.findall("joint"):
            j = UrdfJoint(e)
            self.joints[j.name] = j
        # initTree: iterate joints_ (std::map => sorted by name, byte-wise)
        for name in sorted(self.joints):
            j = self.joints[name]
            child, parent = self.links[j.child], self.links[j.parent]
            child.parent_joint = j
            parent.child_links.append(child)
        roots = [l for l in self.links.values() if l.parent_joint is None]
        if len(roots) != 1:
            raise ValueError("URDF must have exactly one root link, found %d" % len(roots))
        self.root_link = roots[0]

    def _parse_srdf(self, text):
        self.virtual_joints = []
        self.groups = {}
        self.group_order = []
        if not text:
            return
        root = ET.fromstring(text)
        for vj in root.findall("virtual_joint"):
            self.virtual_joints.append(dict(name=vj.attrib["name"], type=vj.attrib["type"],
                                            parent_frame=vj.attrib["parent_frame"], child_link=vj.attrib["child_link"]))
        for g in root.findall("group"):
            spec = dict(chains=[], joints=[], links=[], subgroups=[])
            for e in g:
                if e.tag == "chain":
                    spec["chains"].append((e.attrib["base_link"], e.attrib["tip_link"]))
                elif e.tag == "joint":
                    spec["joints"].append(e.attrib["name"])
                elif e.tag == "link":
                    spec["links"].append(e.attrib["name"])
                elif e.tag == "group":
                    spec["subgroups"].append(e.attrib["name"])
            self.groups[g.attrib["name"]] = spec
            self.group_order.append(g.attrib["name"])

    # -- MoveIt RobotModel ------------------------------------------------------------------
    def _build_moveit_model(self):
        # root joint: SRDF virtual joint whose child is the root link, else an assumed fixed joint
        self.root_joint_name = "ASSUMED_FIXED_ROOT_JOINT"
        self.root_joint_type = "fixed"
        self.world_frame_name = ""
        for vj in self.virtual_joints:
            if vj["child_link"] == self.root_link.name:
                self.root_joint_name = vj["name"]
                self.root_joint_type = vj["type"]
        # exotica: world frame = parent_frame of the virtual joint named like the root joint (kinematic_tree.cpp:178-191)
        for vj in self.virtual_joints:
            if vj["name"] == self.root_joint_name:
                self.world_frame_name = vj["parent_frame"]
        if not self.world_frame_name:
            self.world_frame_name = "world_frame"
        # DFS pre-order joint list: (joint name, type, variable names, parent link, child link)
        self.moveit_joints = []

        def rec(link):
            if link.parent_joint is None:
                name, typ = self.root_joint_name, self.root_joint_type
                if typ == "floating":
                    var = [name + s for s in ("/trans_x", "/trans_y", "/trans_z", "/rot_x", "/rot_y", "/rot_z", "/rot_w")]
                elif typ == "planar":
                    var = [name + s for s in ("/x", "/y", "/theta")]
                else:
                    var = []
                self.moveit_joints.append(dict(name=name, type=typ, vars=var, child=link.name, parent=None, mimic=False))
            else:
                j = link.parent_joint
                typ = j.type if j.type in ("revolute", "continuous", "prismatic", "planar", "floating") else "fixed"
                var = [j.name] if typ in ("revolute", "continuous", "prismatic") else []
                self.moveit_joints.append(dict(name=j.name, type=typ, vars=var, child=link.name, parent=j.parent,
                                               mimic=j.mimic is not None))
            for c in link.child_links:
                rec(c)

        rec(self.root_link)
        self.moveit_joint_index = {j["name"]: i for i, j in enumerate(self.moveit_joints)}
        self.moveit_variable_names = [v for j in self.moveit_joints for v in j["vars"]]
        self.link_parent_joint = {j["child"]: j["name"] for j in self.moveit_joints}

    def _group_joint_names(self, name, seen=None):
        """RobotModel::addJointModelGroup: joints of a group (unordered set of joint names)."""
        seen = seen or set()
        if name in seen:
            return set()
        seen.add(name)
        spec = self.groups[name]
        out = set()
        for base, tip in spec["chains"]:
            if base not in self.links or tip not in self.links:
                continue
            # from the tip up until base (or the root)
            cj, lm, found = [], tip, False
            while lm is not None:
                if lm == base:
                    found = True
                    break
                pj = self.link_parent_joint[lm]
                cj.append(pj)
                lm = self.moveit_joints[self.moveit_joint_index[pj]]["parent"]
            if not found:
                # chain through a common ancestor: walk from base up to the first link already on the tip path
                tip_path = []
                lm = tip
                while lm is not None:
                    tip_path.append(lm)
                    lm = self.moveit_joints[self.moveit_joint_index[self.link_parent_joint[lm]]]["parent"]
                cj2, lm = [], base
                while lm is not None and lm not in tip_path:
                    cj2.append(self.link_parent_joint[lm])
                    lm = self.moveit_joints[self.moveit_joint_index[self.link_parent_joint[lm]]]["parent"]
                if lm is not None:
                    cj = []
                    l2 = tip
                    while l2 != lm:
                        cj.append(self.link_parent_joint[l2])
                        l2 = self.moveit_joints[self.moveit_joint_index[self.link_parent_joint[l2]]]["parent"]
                    cj += cj2
                    found = True
            if found:
                out.update(cj)
        for j in spec["joints"]:
            if j in self.moveit_joint_index:
                out.add(j)
        for l in spec["links"]:
            if l in self.link_parent_joint:
                out.add(self.link_parent_joint[l])
        for sg in spec["subgroups"]:
            if sg in self.groups:
                out.update(self._group_joint_names(sg, seen))
        return out

    def _variable_bounds(self, var):
        """MoveIt jointBoundsFromURDF + exotica ResetJointLimits (kinematic_tree.cpp:1217-1256)."""
        j = self.joints.get(var)
        if j is None or j.type == "continuous":
            return -INF, INF, INF
        lo = hi = None
        if j.safety is not None:
            lo, hi = j.safety["lower"], j.safety["upper"]
            if j.limits is not None:
                if j.limits["lower"] > lo:
                    lo = j.limits["lower"]
                if j.limits["upper"] < hi:
                    hi = j.limits["upper"]
        elif j.limits is not None:
            lo, hi = j.limits["lower"], j.limits["upper"]
        vel = INF
        if j.limits is not None and abs(j.limits["velocity"]) > 2.220446049250313e-16:
            vel = abs(j.limits["velocity"])
        if lo is None:
            return -INF, INF, vel
        return lo, hi, vel

    # -- exotica KinematicTree::Instantiate ---------------------------------------------------
    def _instantiate(self, joint_group):
        self.model_joint_names = list(self.moveit_variable_names)
        if joint_group and joint_group not in self.groups:
            raise ValueError("Joint group '%s' not defined in the robot model" % joint_group)
        if not joint_group:
            # all active joints (no fixed, no mimic): joint NAMES (kinematic_tree.cpp:130-131)
            self.controlled_joint_names = [j["name"] for j in self.moveit_joints if j["vars"] and not j["mimic"]]
        else:
            names = self._group_joint_names(joint_group)
            ordered = sorted(names, key=lambda n: self.moveit_joint_index[n])
            self.controlled_joint_names = [v for n in ordered for v in self.moveit_joints[self.moveit_joint_index[n]]["vars"]]

    # -- exotica KinematicTree::BuildTree ------------------------------------------------------
    def _build_tree(self):
        el = []
        el.append(Element(-1, -1, self.world_frame_name, self.root_joint_name, NONE))
        rj = self.root_joint_name
        if self.root_joint_type == "fixed":
            self.base_type = "FIXED"
        elif self.root_joint_type == "floating":
            self.base_type = "FLOATING"
            types = [TRANS_X, TRANS_Y, TRANS_Z, ROT_Z, ROT_Y, ROT_X]
            suffix = ["/trans_x", "/trans_y", "/trans_z", "/rot_z", "/rot_y", "/rot_x"]
            for i in range(6):
                el.append(Element(i, i, self.world_frame_name + suffix[i], rj + suffix[i], types[i]))
            rot_w = rj + "/rot_w"
            if rot_w in self.controlled_joint_names:
                self.controlled_joint_names.remove(rot_w)
            if rot_w in self.model_joint_names:
                self.model_joint_names.remove(rot_w)
        elif self.root_joint_type == "planar":
            self.base_type = "PLANAR"
            types = [TRANS_X, TRANS_Y, ROT_Z]
            names = [rj + "/x", rj + "/y", rj + "/theta"]
            for i in range(3):
                el.append(Element(i, i, names[i], names[i], types[i]))
        else:
            raise ValueError("Unsupported root joint type: %s" % self.root_joint_type)

        # AddElementFromSegmentMapIterator: recursive pre-order, id = model_tree_.size() at insertion
        def add(link, parent_index):
            j = link.parent_joint
            if j is None:
                # KDL tree root segment: Segment(root_name) => Joint("NoName", None), identity tip
                e = Element(len(el), parent_index, link.name, "NoName", NONE)
            else:
                F = frame_from_origin(j.xyz, j.rpy)
                if j.type in ("revolute", "continuous", "prismatic"):
                    M = F[:9]
                    a = [M[0] * j.axis[0] + M[1] * j.axis[1] + M[2] * j.axis[2],
                         M[3] * j.axis[0] + M[4] * j.axis[1] + M[5] * j.axis[2],
                         M[6] * j.axis[0] + M[7] * j.axis[1] + M[8] * j.axis[2]]
                    n = _kdl_norm(a)
                    a = [a[0] / n, a[1] / n, a[2] / n]  # KDL::Joint ctor: axis / axis.Norm()
                    jt = TRANS_AXIS if j.type == "prismatic" else ROT_AXIS
                    # f_tip = joint.pose(0)^-1 * F = (I, -origin) * F = (F.M, F.p - origin)
                    tip = F[:9] + [F[9] - F[9], F[10] - F[10], F[11] - F[11]]
                    e = Element(len(el), parent_index, link.name, j.name, jt, origin=F[9:], axis=a, f_tip=tip)
                else:
                    e = Element(len(el), parent_index, link.name, j.name, NONE, f_tip=F)
            e.mass, e.cog = link.mass, list(link.cog)
            el.append(e)
            idx = len(el) - 1
            for c in link.child_links:
                add(c, idx)

        add(self.root_link, len(el) - 1)
        # "Set root inertial" (kinematic_tree.cpp:264-276): the URDF root link's inertial is written to model_tree_[2] (fixed
        # base), [7] (floating) or [4] (planar).  [7] and [4] are the root link itself; for a FIXED base the root link is
        # model_tree_[1], so the FIRST CHILD of the root loses its own inertia to the root's (a reference quirk the
        # CenterOfMass goldens of test_maps.cpp:789-827 depend on: link1 of the test robot counts as massless).
        k = {"FIXED": 2, "FLOATING": 7, "PLANAR": 4}[self.base_type]
        if k < len(el):
            el[k].mass, el[k].cog = self.root_link.mass, list(self.root_link.cog)
        self.elements = el
        self.n_controlled = len(self.controlled_joint_names)
        if self.n_controlled < 1:
            raise ValueError("No update joints specified!")
        self.controlled_elements = [-1] * self.n_controlled
        self.model_link_names = []
        self.controlled_link_names = []
        self.controlled_base_type = "FIXED"
        for i, e in enumerate(el):
            e.is_robot_link = True
            cid = -1
            for k, n in enumerate(self.controlled_joint_names):  # IsControlled
                if n == e.joint_name:
                    cid = k
                    break
            e.control_id = cid
            self.model_link_names.append(e.link_name)
            if cid >= 0:
                self.controlled_elements[cid] = i
                self.controlled_link_names.append(e.link_name)
                if self.root_joint_name in e.joint_name:
                    self.controlled_base_type = self.base_type
        el[0].is_robot_link = False
        self.link_index = {}
        for i, e in enumerate(el):
            self.link_index[e.link_name] = i  # tree_map_: later duplicates overwrite, like std::map operator[]
        self.joint_element = {e.joint_name: i for i, e in enumerate(el)}
        # mimic joints (kinematic_tree.cpp:342-356)
        for link in self.links.values():
            j = link.parent_joint
            if j is not None and j.mimic is not None:
                e = el[self.link_index[link.name]]
                e.mimic_elem = self.joint_element[j.mimic["joint"]]
                e.mimic_mult = j.mimic["multiplier"]
                e.mimic_off = j.mimic["offset"]
        # limits (ResetJointLimits)
        self.lower = np.zeros(self.n_controlled)
        self.upper = np.zeros(self.n_controlled)
        self.velocity_limits = np.zeros(self.n_controlled)
        for var in self.moveit_variable_names:
            if var in self.controlled_joint_names:
                k = self.controlled_joint_names.index(var)
                self.lower[k], self.upper[k], self.velocity_limits[k] = self._variable_bounds(var)
        if self.controlled_base_type == "FLOATING":
            self.lower[:3], self.upper[:3] = -INF, INF
            self.lower[3:6], self.upper[3:6] = -math.pi, math.pi
        elif self.controlled_base_type == "PLANAR":
            self.lower[:2], self.upper[:2] = -INF, INF
            self.lower[2], self.upper[2] = -math.pi, math.pi

    def _add_collision_elements(self):
        """scene.cpp:795-821: one fixed element per collision shape, appended after the model elements."""
        n_model = len(self.elements)
        for i in range(n_model):
            e = self.elements[i]
            link = self.links.get(e.link_name)
            if link is None:
                continue
            for k, frame in enumerate(link.collisions):
                ce = Element(len(self.elements), i, "%s_collision_%d" % (e.link_name, k), "NoName", NONE, f_tip=frame)
                self.elements.append(ce)
                self.link_index[ce.link_name] = len(self.elements) - 1

    # -- helpers used by tests ----------------------------------------------------------------------
    def add_environment_element(self, name, frame12=None, parent="", generated_slot=-1):
        """Fixed element attached to `parent` (default: root), optionally trajectory-driven
        (KinematicTree::AddEnvironmentElement / Scene trajectories, scene.cpp:329-335,917-935)."""
        p = self.link_index[parent] if parent else 0
        e = Element(len(self.elements), p, name, "NoName", NONE, f_tip=frame12 if frame12 is not None else IDENTITY)
        e.generated = generated_slot
        self.elements.append(e)
        self.link_index[name] = len(self.elements) - 1
        self.tree_state = np.zeros(len(self.elements))
        return len(self.elements) - 1

    def flat(self):
        el = self.elements
        return dict(
            id=np.array([e.id for e in el], dtype=np.int32),
            parent=np.array([e.parent for e in el], dtype=np.int32),
            joint_type=np.array([e.joint_type for e in el], dtype=np.int32),
            origin=np.array([e.origin for e in el], dtype=np.float64).reshape(-1, 3),
            axis=np.array([e.axis for e in el], dtype=np.float64).reshape(-1, 3),
            f_tip=np.array([e.f_tip for e in el], dtype=np.float64).reshape(-1, 12),
            control_id=np.array([e.control_id for e in el], dtype=np.int32),
            mimic_elem=np.array([e.mimic_elem for e in el], dtype=np.int32),
            mimic_mult=np.array([e.mimic_mult for e in el], dtype=np.float64),
            mimic_off=np.array([e.mimic_off for e in el], dtype=np.float64),
            generated=np.array([e.generated for e in el], dtype=np.int32),
        )


def _kdl_norm(d):
    t1, t2 = abs(d[0]), abs(d[1])
    if t1 >= t2:
        t2 = abs(d[2])
        if t1 >= t2:
            if t1 == 0:
                return 0.0
            return t1 * math.sqrt(1 + (d[1] / d[0]) ** 2 + (d[2] / d[0]) ** 2)
        return t2 * math.sqrt(1 + (d[0] / d[2]) ** 2 + (d[1] / d[2]) ** 2)
    t1 = abs(d[2])
    if t2 > t1:
        return t2 * math.sqrt(1 + (d[0] / d[1]) ** 2 + (d[2] / d[1]) ** 2)
    return t1 * math.sqrt(1 + (d[0] / d[2]) ** 2 + (d[1] / d[2]) ** 2)

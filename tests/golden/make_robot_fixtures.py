#!/usr/bin/env python3
"""Generate the kinematics-only robot fixtures under tests/fixtures/robots/.

Run in the build container only (needs /root/reference, which does NOT exist on the GPU box):

    python tests/golden/make_robot_fixtures.py

The reference ships full URDF/SRDF robot descriptions (visual meshes, gazebo plugins,
transmissions, collision matrices ...).  The hot path only consumes the kinematic skeleton:
link names, inertial mass + centre of gravity (needed by the Valkyrie centre-of-mass
known-answer test), the number/origin of collision shapes (the reference adds one fixed
tree element per collision shape, scene.cpp:795-821, which matters for the CPU baseline's
per-update cost), joints (type, origin, axis, parent, child, limits, soft limits, mimic) and
from the SRDF the planning groups and the virtual (root) joint.  This script extracts exactly
that, keeping every numeric literal verbatim (e.g. rpy="0 0 3.14159265359" on the LWR base
joint), and writes compact fixture files.

Sources (read-only):
  exotica_examples/resources/robots/lwr_simplified.{urdf,srdf}
  exotica_examples/resources/robots/valkyrie_sim.{urdf,srdf}
  exotations/exotica_core_task_maps/test/test_maps.cpp:43-44  (inline 3-DoF test robot)
"""
import os
import re
import sys
import xml.etree.ElementTree as ET

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "fixtures", "robots")


def attrs(e, names):
    return " ".join('%s="%s"' % (n, e.attrib[n]) for n in names if n in e.attrib)


def strip_urdf(text):
    root = ET.fromstring(text)
    out = ['<robot name="%s">' % root.attrib.get("name", "robot")]
    for link in root.findall("link"):
        body = []
        inertial = link.find("inertial")
        if inertial is not None:
            parts = []
            mass = inertial.find("mass")
            if mass is not None:
                parts.append("<mass %s/>" % attrs(mass, ["value"]))
            origin = inertial.find("origin")
            if origin is not None:
                parts.append("<origin %s/>" % attrs(origin, ["xyz", "rpy"]))
            body.append("<inertial>%s</inertial>" % "".join(parts))
        for col in link.findall("collision"):
            origin = col.find("origin")
            o = "<origin %s/>" % attrs(origin, ["xyz", "rpy"]) if origin is not None else ""
            body.append("<collision>%s</collision>" % o)
        if body:
            out.append('  <link name="%s">%s</link>' % (link.attrib["name"], "".join(body)))
        else:
            out.append('  <link name="%s"/>' % link.attrib["name"])
    for joint in root.findall("joint"):
        # sensors inside <gazebo> are not direct children of <robot>, so only real joints land here
        body = []
        for tag, names in (("origin", ["xyz", "rpy"]), ("axis", ["xyz"]), ("parent", ["link"]),
                           ("child", ["link"]), ("limit", ["lower", "upper", "velocity", "effort"]),
                           ("safety_controller", ["soft_lower_limit", "soft_upper_limit", "k_position", "k_velocity"]),
                           ("mimic", ["joint", "multiplier", "offset"])):
            e = joint.find(tag)
            if e is not None:
                body.append("<%s %s/>" % (tag, attrs(e, names)))
        out.append('  <joint name="%s" type="%s">%s</joint>' % (joint.attrib["name"], joint.attrib["type"], "".join(body)))
    out.append("</robot>")
    return "\n".join(out) + "\n"


def strip_srdf(text):
    root = ET.fromstring(text)
    out = ['<robot name="%s">' % root.attrib.get("name", "robot")]
    for group in root.findall("group"):
        out.append('  <group name="%s">' % group.attrib["name"])
        for e in group:
            if e.tag == "chain":
                out.append("    <chain %s/>" % attrs(e, ["base_link", "tip_link"]))
            elif e.tag in ("joint", "link", "group"):
                out.append('    <%s name="%s"/>' % (e.tag, e.attrib["name"]))
        out.append("  </group>")
    for vj in root.findall("virtual_joint"):
        out.append("  <virtual_joint %s/>" % attrs(vj, ["name", "type", "parent_frame", "child_link"]))
    out.append("</robot>")
    return "\n".join(out) + "\n"


def cxx_string_literal(src, var):
    m = re.search(r'%s\s*=\s*"((?:[^"\\]|\\.)*)";' % re.escape(var), src)
    if not m:
        raise RuntimeError("literal %s not found" % var)
    return m.group(1).replace('\\"', '"')


def main():
    if not os.path.isdir(REF):
        sys.exit("needs %s (build container only)" % REF)
    os.makedirs(OUT, exist_ok=True)
    robots = os.path.join(REF, "exotica_examples", "resources", "robots")
    for name in ("lwr_simplified", "valkyrie_sim"):
        with open(os.path.join(robots, name + ".urdf")) as f:
            urdf = strip_urdf(f.read())
        with open(os.path.join(robots, name + ".srdf")) as f:
            srdf = strip_srdf(f.read())
        with open(os.path.join(OUT, name + ".urdf"), "w") as f:
            f.write(urdf)
        with open(os.path.join(OUT, name + ".srdf"), "w") as f:
            f.write(srdf)
    with open(os.path.join(REF, "exotations", "exotica_core_task_maps", "test", "test_maps.cpp")) as f:
        src = f.read()
    with open(os.path.join(OUT, "test_robot.urdf"), "w") as f:
        f.write(strip_urdf(cxx_string_literal(src, "urdf_string_")))
    with open(os.path.join(OUT, "test_robot.srdf"), "w") as f:
        f.write(strip_srdf(cxx_string_literal(src, "srdf_string_")))
    for fn in sorted(os.listdir(OUT)):
        print(fn, os.path.getsize(os.path.join(OUT, fn)))


if __name__ == "__main__":
    main()

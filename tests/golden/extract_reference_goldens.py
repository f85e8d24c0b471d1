#!/usr/bin/env python3
"""Extracts the golden vectors the reference's own tests hold for the hot path and writes tests/golden/reference_goldens.json.

    python tests/golden/extract_reference_goldens.py [/root/reference]

Run in the build container (the reference tree does not travel to the GPU box; the JSON does).  Sources:
  exotations/exotica_core_task_maps/test/test_maps.cpp -- `X << ...; Y << ...; jacobian << ...;` blocks of testEffPosition,
  testDistance, testSphereCollision, testJointPose (3 blocks) and testCoM (first 3 blocks; the fourth attaches a payload object
  through Scene::AddObject, which is outside the hot path), evaluated on the 3-DoF test robot of that file (tests/fixtures/robots/
  test_robot.*).  Each case records the task-map initializer the test builds (transcribed next to the block it belongs to), the
  five joint states X[5][N], the expected Phi Y[5][M] and Jacobians J[5][M][N] (the test stacks them as a (5 M) x N matrix), and
  the tolerance the reference's test_values() applies (||.|| < 1e-5 on 6-significant-figure numbers).
"""
import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))

# (test name, index of the X/Y/jacobian block inside the test, map spec as tests/helpers.py and oracle.set_problem take it,
#  the reference's initializer for exotica_b200.Setup.create_problem)
CASES = [
    ("testEffPosition", 0, {"type": "EffPosition", "EndEffector": [{"Link": "endeff"}]},
     ["exotica/EffPosition", {"Name": "MyTask", "EndEffector": [["exotica/Frame", {"Link": "endeff"}]]}]),
    ("testDistance", 0, {"type": "Distance", "EndEffector": [{"Link": "endeff"}]},
     ["exotica/Distance", {"Name": "MyTask", "EndEffector": [["exotica/Frame", {"Link": "endeff"}]]}]),
    ("testSphereCollision", 0,
     {"type": "SphereCollision", "Precision": 1e-2, "EndEffector": [{"Link": "base", "Radius": 0.3, "Group": "base"}, {"Link": "endeff", "Radius": 0.3, "Group": "eff"}]},
     ["exotica/SphereCollision", {"Name": "MyTask", "Precision": "1e-2", "ReferenceFrame": "base",
                                  "EndEffector": [["exotica/Sphere", {"Link": "base", "Radius": "0.3", "Group": "base"}],
                                                  ["exotica/Sphere", {"Link": "endeff", "Radius": "0.3", "Group": "eff"}]]}]),
    ("testJointPose", 0, {"type": "JointPose"}, ["exotica/JointPose", {"Name": "MyTask"}]),
    ("testJointPose", 1, {"type": "JointPose", "JointRef": [0.5, 0.5, 0.5]}, ["exotica/JointPose", {"Name": "MyTask", "JointRef": "0.5 0.5 0.5"}]),
    ("testJointPose", 2, {"type": "JointPose", "JointMap": [0]}, ["exotica/JointPose", {"Name": "MyTask", "JointMap": "0"}]),
    ("testCoM", 0, {"type": "CenterOfMass", "EnableZ": True, "links": None}, ["exotica/CenterOfMass", {"Name": "MyTask", "EnableZ": "1"}]),
    ("testCoM", 1, {"type": "CenterOfMass", "EnableZ": True, "links": ["link2", "endeff"]},
     ["exotica/CenterOfMass", {"Name": "MyTask", "EnableZ": "1", "EndEffector": [["exotica/Frame", {"Link": "link2"}], ["exotica/Frame", {"Link": "endeff"}]]}]),
    ("testCoM", 2, {"type": "CenterOfMass", "EnableZ": False, "links": None}, ["exotica/CenterOfMass", {"Name": "MyTask", "EnableZ": "0"}]),
]


def numbers(text):
    return [float(v) for v in re.findall(r"[-+]?(?:\d+\.?\d*|\.\d+)(?:[eE][-+]?\d+)?", text)]


def blocks_of(test_body):
    out = []
    for m in re.finditer(r"\bX\s*<<([^;]*);\s*Y\s*<<([^;]*);\s*jacobian\s*<<([^;]*);", test_body, re.S):
        out.append((numbers(m.group(1)), numbers(m.group(2)), numbers(m.group(3)), test_body[:m.start()].count("\n")))
    return out


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    rel = "exotations/exotica_core_task_maps/test/test_maps.cpp"
    src = open(os.path.join(ref, rel)).read()
    tests = {}
    starts = [(m.start(), m.group(1)) for m in re.finditer(r"^TEST\(ExoticaTaskMaps,\s*(\w+)\)", src, re.M)]
    for i, (pos, name) in enumerate(starts):
        end = starts[i + 1][0] if i + 1 < len(starts) else len(src)
        tests[name] = (src[pos:end], src[:pos].count("\n") + 1)
    cases = []
    for name, k, spec, init in CASES:
        body, line0 = tests[name]
        X, Y, J, rel_line = blocks_of(body)[k]
        n = 3  # the test robot's controlled joints
        rows = len(X) // n
        m = len(Y) // rows
        assert len(X) == rows * n and len(Y) == rows * m and len(J) == rows * m * n, (name, k, len(X), len(Y), len(J))
        cases.append({"name": "%s[%d]" % (name, k), "source": "%s:%d" % (rel, line0 + rel_line), "robot": "test_robot", "joint_group": "arm",
                      "map": spec, "initializer": init, "tolerance": 1e-5,
                      "X": [X[r * n:(r + 1) * n] for r in range(rows)],
                      "Y": [Y[r * m:(r + 1) * m] for r in range(rows)],
                      "J": [[J[(r * m + i) * n:(r * m + i + 1) * n] for i in range(m)] for r in range(rows)]})
    out = os.path.join(HERE, "reference_goldens.json")
    with open(out, "w") as f:
        json.dump({"generated_by": "tests/golden/extract_reference_goldens.py", "reference": "ipab-slmc/exotica", "cases": cases}, f, indent=1)
    print("wrote %s: %d cases" % (out, len(cases)))


if __name__ == "__main__":
    main()

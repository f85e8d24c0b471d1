#!/usr/bin/env python3
"""Copy the reference's own problem configurations and scene files that the hot path is tested on into
tests/fixtures/exotica_examples/ (laid out like the reference's `exotica_examples` package, so that
Setup.set_package_path("exotica_examples", <that directory>) resolves their {exotica_examples}/... paths unmodified).

Run in the build container only (needs /root/reference, which does NOT exist on the GPU box):

    python tests/golden/make_config_fixtures.py

Files (verbatim copies of small data files; the robots are the kinematics-only extracts of make_robot_fixtures.py):
  resources/configs/example_ik.xml                  BASELINE.json configs[0]
  resources/configs/example_sphere_collision.xml    BASELINE.json configs[4]'s fixture (LoadScene + Trajectory File)
  resources/scenes/example_moving_obstacle.scene    MoveIt scene geometry with one box, "Obstacle"
  resources/scenes/example_moving_obstacle.traj     its Cartesian trajectory
  resources/robots/lwr_simplified.{urdf,srdf}       copied from tests/fixtures/robots/
"""
import os
import shutil
import sys

REF = "/root/reference/exotica_examples"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "fixtures", "exotica_examples")
FILES = ["resources/configs/example_ik.xml", "resources/configs/example_sphere_collision.xml",
         "resources/scenes/example_moving_obstacle.scene", "resources/scenes/example_moving_obstacle.traj"]


def main():
    if not os.path.isdir(REF):
        sys.exit("needs %s (build container only)" % REF)
    for rel in FILES:
        dst = os.path.join(OUT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(REF, rel), dst)
    robots = os.path.join(OUT, "resources", "robots")
    os.makedirs(robots, exist_ok=True)
    for name in ("lwr_simplified.urdf", "lwr_simplified.srdf"):
        shutil.copyfile(os.path.join(HERE, "..", "fixtures", "robots", name), os.path.join(robots, name))
    for root, _, files in os.walk(OUT):
        for fn in sorted(files):
            p = os.path.join(root, fn)
            print(os.path.relpath(p, OUT), os.path.getsize(p))


if __name__ == "__main__":
    main()

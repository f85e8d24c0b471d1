"""exotica_b200 -- B200-native batched kinematics behind EXOTica's Scene / TaskMap / PlanningProblem surface.

The compute path is the CUDA library exotica_b200/lib/libexotica_b200.so (C-ABI: include/exotica_b200.h,
include/exotica_b200_host.h).  This package only forwards to it; importing it does not need a GPU, evaluating does.
"""
from ._capi import (ExoticaError, LAYOUT_AOS_COLMAJOR, LAYOUT_AOS_ROWMAJOR, LAYOUT_SOA, DeviceBuffer, PinnedArray,  # noqa: F401
                    device_count)
from .pyexotica import IKSolver, KinematicTree, PlanningProblem, Scene, Setup, Trajectory, get_frame, initializer_to_xml  # noqa: F401

__all__ = ["Setup", "IKSolver", "Trajectory", "Scene", "KinematicTree", "PlanningProblem", "ExoticaError", "get_frame", "initializer_to_xml", "DeviceBuffer",
           "PinnedArray", "device_count", "LAYOUT_SOA", "LAYOUT_AOS_ROWMAJOR", "LAYOUT_AOS_COLMAJOR"]

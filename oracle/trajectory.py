"""TEST INFRASTRUCTURE ONLY -- CPU restatement of exotica::Trajectory (exotica_core/src/trajectory.cpp:47-141).

The reference delegates to orocos_kdl 1.4 (not vendored): Trajectory_Composite of Trajectory_Stationary /
Trajectory_Segment(Path_Line(RotationalInterpolation_SingleAxis, eqradius), VelocityProfile_Spline).  This file restates those
published algorithms with numpy, written independently of exotica_b200/csrc/host/trajectory.cpp (the product's C++ version)
so that the two can be compared.  Parity unpinned: the reference holds no golden values for trajectories.
"""
import numpy as np

from .oracle import get_frame


def _rot_angle_axis(R):
    """KDL::Rotation::GetRotAngle(axis, eps=1e-6) on a 3x3 matrix."""
    eps, eps2 = 1e-6, 1e-5
    if abs(R[0, 1] - R[1, 0]) < eps and abs(R[0, 2] - R[2, 0]) < eps and abs(R[1, 2] - R[2, 1]) < eps:
        if abs(R[0, 1] + R[1, 0]) < eps2 and abs(R[0, 2] + R[2, 0]) < eps2 and abs(R[1, 2] + R[2, 1]) < eps2 and abs(np.trace(R) - 3) < eps2:
            return 0.0, np.array([0.0, 0.0, 1.0])
        xx, yy, zz = (R[0, 0] + 1) / 2, (R[1, 1] + 1) / 2, (R[2, 2] + 1) / 2
        xy, xz, yz = (R[0, 1] + R[1, 0]) / 4, (R[0, 2] + R[2, 0]) / 4, (R[1, 2] + R[2, 1]) / 4
        if xx > yy and xx > zz:
            x = np.sqrt(xx)
            v = np.array([x, xy / x, xz / x])
        elif yy > zz:
            y = np.sqrt(yy)
            v = np.array([xy / y, y, yz / y])
        else:
            z = np.sqrt(zz)
            v = np.array([xz / z, yz / z, z])
        return np.pi, v
    f = min(1.0, max(-1.0, (np.trace(R) - 1) / 2))
    v = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    return float(np.arccos(f)), v / np.linalg.norm(v)


def _rot2(v, angle):
    """KDL::Rotation::Rot2: Rodrigues' formula for a unit axis."""
    K = np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])
    return np.eye(3) + np.sin(angle) * K + (1 - np.cos(angle)) * (K @ K)


class Trajectory(object):
    def __init__(self, data, radius=1.0):
        if isinstance(data, str):
            tok = data.split()
            radius, n, m = float(tok[0]), int(tok[1]), int(tok[2])
            data = np.array([float(v) for v in tok[3:3 + n * m]]).reshape(n, m)
        data = np.asarray(data, dtype=np.float64)
        if data.shape[1] not in (4, 7, 8) or data.shape[0] < 2:
            raise ValueError("Invalid trajectory data size!")
        self.radius = radius
        self.segments = []
        for i in range(data.shape[0] - 1):
            f1, f2 = get_frame(data[i, 1:]), get_frame(data[i + 1, 1:])
            dt = data[i + 1, 0] - data[i, 0]
            if dt <= 0:
                raise ValueError("Time indices must be monotonically increasing!")
            self.segments.append((dt, np.asarray(f1), np.asarray(f2)))

    def get_duration(self):
        return sum(s[0] for s in self.segments)

    @staticmethod
    def _segment_pos(seg, time, radius):
        dt, f1, f2 = seg
        if np.max(np.abs(f1 - f2)) <= 1e-6:  # KDL::Equal(f1, f2, 1e-6) -> Trajectory_Stationary
            return f1.copy()
        R1, R2 = f1[:9].reshape(3, 3), f2[:9].reshape(3, 3)
        p1, p2 = f1[9:], f2[9:]
        d = p2 - p1
        dist = np.linalg.norm(d)
        u = d / dist if dist > 0 else d
        alpha, axis = _rot_angle_axis(R1.T @ R2)
        if alpha != 0 and alpha * radius > dist:
            length, scale_rot, scale_lin = alpha * radius, 1 / radius, dist / (alpha * radius)
        elif dist != 0:
            length, scale_rot, scale_lin = dist, alpha / dist, 1.0
        else:
            length, scale_rot, scale_lin = 0.0, 1.0, 1.0
        tau = time / dt
        s = length * (3 * tau ** 2 - 2 * tau ** 3)  # VelocityProfile_Spline, cubic with zero end velocities
        out = np.empty(12)
        out[:9] = (R1 @ _rot2(axis, s * scale_rot)).reshape(9)
        out[9:] = p1 + u * s * scale_lin
        return out

    def get_position(self, t):
        if t < 0:
            return self._segment_pos(self.segments[0], 0.0, self.radius)
        previous = 0.0
        for seg in self.segments:
            if t < previous + seg[0]:
                return self._segment_pos(seg, t - previous, self.radius)
            previous += seg[0]
        return self._segment_pos(self.segments[-1], self.segments[-1][0], self.radius)

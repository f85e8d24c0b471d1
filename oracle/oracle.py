"""ctypes front-end of the CPU oracle (oracle/libexo_oracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes
import os
import subprocess

import numpy as np

from . import robot_model

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

ROTATION_TYPES = {"Quaternion": 0, "RPY": 1, "ZYX": 2, "ZYZ": 3, "AngleAxis": 4, "Matrix": 5}
ROTATION_LENGTH = [4, 3, 3, 3, 3, 9]
MAP_KINDS = {"EffFrame": 0, "EffPosition": 1, "JointLimit": 2, "SphereCollision": 3, "EffOrientation": 5, "EffPositionXY": 6, "Distance": 7,
             "JointPose": 8, "EffAxisAlignment": 9, "CenterOfMass": 10}

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int)


def build(force=False):
    so = os.path.join(_HERE, "libexo_oracle.so")
    src = os.path.join(_HERE, "exo_oracle.c")
    if force or not os.path.exists(so) or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(so)):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


BUILD_FLAGS = "gcc -O3 -ffp-contract=off -std=c99 (generic x86-64)"


def use_native_build():
    """bench.py's CPU-baseline legs: compile the oracle on THIS host with -march=native (oracle/Makefile `native`) and load that
    copy instead of the generic one.  Must be called before the first lib(); falls back to the generic build if gcc fails."""
    global BUILD_FLAGS, _NATIVE
    try:
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "native"])
        _NATIVE = os.path.join(_HERE, "libexo_oracle_native.so")
        BUILD_FLAGS = "gcc -O3 -march=native -ffp-contract=off -std=c99, built on the measuring host"
    except Exception:
        _NATIVE = None
    return BUILD_FLAGS


_NATIVE = None


def lib():
    global _LIB
    if _LIB is None:
        so = _NATIVE or os.path.join(_HERE, "libexo_oracle.so")
        if not os.path.exists(so):
            build()
        L = ctypes.CDLL(so)
        L.exo_oracle_create.restype = ctypes.c_void_p
        L.exo_oracle_create.argtypes = [ctypes.c_int, _ip, _ip, _ip, _dp, _dp, _dp, _ip, _ip, _dp, _dp, _ip, ctypes.c_int]
        L.exo_oracle_destroy.argtypes = [ctypes.c_void_p]
        L.exo_oracle_set_tree_state.argtypes = [ctypes.c_void_p, _dp]
        L.exo_oracle_set_limits.argtypes = [ctypes.c_void_p, _dp, _dp]
        L.exo_oracle_set_generated.argtypes = [ctypes.c_void_p, ctypes.c_int, _dp]
        L.exo_oracle_request_frames.argtypes = [ctypes.c_void_p, ctypes.c_int, _ip, _dp, _ip, _dp]
        L.exo_oracle_update.argtypes = [ctypes.c_void_p, _dp]
        L.exo_oracle_get_phi.argtypes = [ctypes.c_void_p, _dp]
        L.exo_oracle_get_jacobian.argtypes = [ctypes.c_void_p, _dp]
        L.exo_oracle_get_element_frame.argtypes = [ctypes.c_void_p, ctypes.c_int, _dp]
        L.exo_oracle_set_rotation.argtypes = [_dp, ctypes.c_int, _dp]
        L.exo_oracle_get_frame.argtypes = [_dp, ctypes.c_int, _dp]
        L.exo_oracle_problem_clear.argtypes = [ctypes.c_void_p]
        L.exo_oracle_problem_add_map.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                 ctypes.c_double, _ip, _dp, ctypes.c_int]
        L.exo_oracle_problem_length_phi.argtypes = [ctypes.c_void_p]
        L.exo_oracle_problem_length_jacobian.argtypes = [ctypes.c_void_p]
        L.exo_oracle_problem_update.argtypes = [ctypes.c_void_p, _dp, _dp, _dp]
        L.exo_oracle_problem_update_batch.argtypes = [ctypes.c_void_p, _dp, ctypes.c_long, ctypes.c_int, _dp, ctypes.c_int,
                                                      _dp, _dp, ctypes.c_int]
        L.exo_oracle_problem_hessian.argtypes = [ctypes.c_void_p, _dp, _dp]
        L.exo_oracle_ik_solve.argtypes = [ctypes.c_void_p, _dp, _dp, _dp, _dp, ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_double,
                                          ctypes.c_double, ctypes.c_double, _dp, _ip]
        L.exo_oracle_ik_solve_ex.argtypes = [ctypes.c_void_p, _dp, _dp, _dp, _dp, ctypes.c_int] + [ctypes.c_double] * 9 + [_dp, _ip, _dp]
        L.exo_oracle_get_rotation.argtypes = [_dp, ctypes.c_int, _dp]
        L.exo_oracle_problem_ydiff.argtypes = [ctypes.c_void_p, _dp, _dp, _dp]
        L.exo_oracle_trajectory_cost.restype = ctypes.c_double
        L.exo_oracle_trajectory_cost.argtypes = [ctypes.c_void_p, _dp, ctypes.c_int, ctypes.c_int, _dp, _dp, ctypes.c_int, ctypes.c_double,
                                                 _dp, _dp, ctypes.c_int, _dp]
        L.exo_oracle_trajectory_cost_jacobian.restype = None
        L.exo_oracle_trajectory_cost_jacobian.argtypes = [ctypes.c_void_p, _dp, ctypes.c_int, ctypes.c_int, _dp, _dp, ctypes.c_int, ctypes.c_double,
                                                          _dp, _dp, ctypes.c_int, _dp]
        L.exo_oracle_aico_task_messages.restype = None
        L.exo_oracle_aico_task_messages.argtypes = [ctypes.c_void_p, _dp, ctypes.c_int, _dp, _dp, ctypes.c_int, ctypes.c_double, _dp, ctypes.c_int,
                                                    _dp, _dp, _dp, _dp]
        _LIB = L
    return _LIB


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def get_frame(v):
    """exotica::GetFrame(vector) -> 12 doubles (row-major R, p)."""
    v = np.ascontiguousarray(v, dtype=np.float64)
    out = np.zeros(12)
    if lib().exo_oracle_get_frame(_d(v), int(v.size), _d(out)) != 0:
        raise ValueError("invalid frame vector of length %d" % v.size)
    return out


def set_rotation(rot9, rot_type):
    rot9 = np.ascontiguousarray(rot9, dtype=np.float64)
    out = np.zeros(9)
    n = lib().exo_oracle_set_rotation(_d(rot9), ROTATION_TYPES[rot_type] if isinstance(rot_type, str) else rot_type, _d(out))
    return out[:n]


def get_rotation(data, rot_type):
    """exotica::GetRotation (tools/conversions.cpp:114-148): encoding -> row-major rotation matrix."""
    data = np.ascontiguousarray(data, dtype=np.float64)
    out = np.zeros(9)
    lib().exo_oracle_get_rotation(_d(data), ROTATION_TYPES[rot_type] if isinstance(rot_type, str) else rot_type, _d(out))
    return out


IDENTITY_FRAME = np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0], dtype=np.float64)


class Oracle(object):
    """KinematicTree + task maps + problem Update of the reference, one configuration at a time."""

    def __init__(self, tree):
        self.tree = tree
        f = tree.flat()
        self._keep = f
        self.n = tree.n_controlled
        self.h = lib().exo_oracle_create(len(tree.elements), _i(f["id"]), _i(f["parent"]), _i(f["joint_type"]),
                                         _d(f["origin"]), _d(f["axis"]), _d(f["f_tip"]), _i(f["control_id"]),
                                         _i(f["mimic_elem"]), _d(f["mimic_mult"]), _d(f["mimic_off"]), _i(f["generated"]),
                                         self.n)
        self.set_limits(tree.lower, tree.upper)
        lib().exo_oracle_set_tree_state(self.h, _d(np.ascontiguousarray(tree.tree_state, dtype=np.float64)))
        self.n_frames = 0
        self.maps = []

    def __del__(self):
        try:
            if self.h:
                lib().exo_oracle_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def set_limits(self, lower, upper):
        self.lower = np.ascontiguousarray(lower, dtype=np.float64).copy()
        self.upper = np.ascontiguousarray(upper, dtype=np.float64).copy()
        lib().exo_oracle_set_limits(self.h, _d(self.lower), _d(self.upper))

    def set_generated(self, slot, frame12):
        lib().exo_oracle_set_generated(self.h, slot, _d(np.ascontiguousarray(frame12, dtype=np.float64)))

    def resolve_frames(self, frames):
        """frames: list of dicts with the reference's Frame initializer fields
        (exotica_core/init/frame.in: Link, LinkOffset, Base, BaseOffset)."""
        ea, eb, oa, ob = [], [], [], []
        for fr in frames:
            link, base = fr.get("Link", ""), fr.get("Base", "")
            for name, what in ((link, "frame_A"), (base, "frame_B")):
                if name and name not in self.tree.link_index:
                    raise KeyError("No %s link exists named '%s'" % (what, name))
            ea.append(self.tree.link_index[link] if link else 0)
            eb.append(self.tree.link_index[base] if base else 0)
            oa.append(get_frame(fr["LinkOffset"]) if "LinkOffset" in fr and fr["LinkOffset"] is not None else IDENTITY_FRAME)
            ob.append(get_frame(fr["BaseOffset"]) if "BaseOffset" in fr and fr["BaseOffset"] is not None else IDENTITY_FRAME)
        return (np.array(ea, dtype=np.int32), np.array(oa, dtype=np.float64).reshape(-1, 12),
                np.array(eb, dtype=np.int32), np.array(ob, dtype=np.float64).reshape(-1, 12))

    def request_frames(self, frames):
        ea, oa, eb, ob = self.resolve_frames(frames)
        self.n_frames = len(frames)
        self.frame_request = (ea, oa, eb, ob)
        if lib().exo_oracle_request_frames(self.h, self.n_frames, _i(ea), _d(oa), _i(eb), _d(ob)) != 0:
            raise ValueError("bad frame request")

    def update(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        if x.size != self.n:
            raise ValueError("Wrong state vector size! Got %d expected %d" % (x.size, self.n))
        lib().exo_oracle_update(self.h, _d(x))

    def phi(self):
        out = np.zeros((self.n_frames, 12))
        lib().exo_oracle_get_phi(self.h, _d(out))
        return out

    def jacobian(self):
        """[F][6][N] (the per-frame KDL::Jacobian, rows vx vy vz wx wy wz)."""
        out = np.zeros((self.n_frames, self.n, 6))
        lib().exo_oracle_get_jacobian(self.h, _d(out))
        return np.ascontiguousarray(out.transpose(0, 2, 1))

    def element_frame(self, elem):
        out = np.zeros(12)
        lib().exo_oracle_get_element_frame(self.h, elem, _d(out))
        return out

    # ---- problem level ----------------------------------------------------------------------
    def set_problem(self, maps):
        """maps: list of dicts {"type": "EffFrame"|"EffPosition"|"JointLimit"|"SphereCollision",
        "EndEffector": [frame dicts (+ Radius, Group for spheres)], "Type": rotation type,
        "SafePercentage": .., "Precision": .., "used": bool}.  Mirrors PlanningProblem::InstantiateBase
        (planning_problem.cpp:206-238): frames of all maps are concatenated in map order."""
        frames = []
        specs = []
        for m in maps:
            fr = list(m.get("EndEffector", []))
            if m["type"] == "EffAxisAlignment":
                # EffAxisAlignment::Initialize (eff_axis_alignment.cpp:39-58): frames_[i + n] = frames_[i] with
                # frame_A_offset.p = Axis.normalized()
                extra = []
                for f in fr:
                    axis = np.asarray(f.get("Axis", [1.0, 0.0, 0.0]), dtype=np.float64)
                    axis = axis / np.linalg.norm(axis)
                    g = dict(f)
                    off = list(f.get("LinkOffset", [0.0, 0.0, 0.0]))
                    g["LinkOffset"] = list(axis) + off[3:]
                    extra.append(g)
                fr = fr + extra
            specs.append((m, len(frames), len(fr)))
            frames.extend(fr)
        self.request_frames(frames)
        lib().exo_oracle_problem_clear(self.h)
        self.maps = []
        for m, start, n in specs:
            kind = MAP_KINDS[m["type"]]
            rot = ROTATION_TYPES[m.get("Type", "RPY")]
            param = 0.0
            grp = rad = None
            if kind == 2:
                param = float(m.get("SafePercentage", 0.0))
            if kind == 8:
                # JointPose::Initialize (joint_pose.cpp:66-97): JointMap defaults to all controlled joints, JointRef to zeros
                jm = list(m.get("JointMap", range(self.n)))
                grp = np.array(jm, dtype=np.int32)
                rad = np.array(m.get("JointRef", [0.0] * len(jm)), dtype=np.float64)
                param = float(len(jm))
            if kind == 9:
                dirs = [np.asarray(f.get("Direction", [1.0, 0.0, 0.0]), dtype=np.float64) for f in m["EndEffector"]]
                rad = np.concatenate([d / np.linalg.norm(d) for d in dirs]) if dirs else np.zeros(0)
            if kind == 10:
                param = 3.0 if m.get("EnableZ", True) else 2.0
                rad = np.array([float(f["Mass"]) for f in m["EndEffector"]], dtype=np.float64)
            if kind == 3:
                param = float(m["Precision"])
                names = sorted(set(f.get("Group", "default") for f in m["EndEffector"]))  # std::map order
                grp = np.array([names.index(f.get("Group", "default")) for f in m["EndEffector"]], dtype=np.int32)
                rad = np.array([float(f["Radius"]) for f in m["EndEffector"]], dtype=np.float64)
            r = lib().exo_oracle_problem_add_map(self.h, kind, start, n, rot, param,
                                                 _i(grp) if grp is not None else None, _d(rad) if rad is not None else None,
                                                 1 if m.get("used", True) else 0)
            if r < 0:
                raise ValueError("bad task map %r" % (m,))
            self.maps.append(m)
        self.length_phi = lib().exo_oracle_problem_length_phi(self.h)
        self.length_jacobian = lib().exo_oracle_problem_length_jacobian(self.h)

    def problem_update(self, x):
        """Returns (Phi[length_Phi], J[length_jacobian][N])."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        phi = np.zeros(self.length_phi)
        jac = np.zeros((self.n, self.length_jacobian))
        lib().exo_oracle_problem_update(self.h, _d(x), _d(phi), _d(jac))
        return phi, np.ascontiguousarray(jac.T)

    def problem_update_batch(self, X, T=1, generated=None, n_threads=1, discard=False):
        """X[E][N] -> Phi[E][length_Phi], J[E][length_jacobian][N] (row-major per evaluation)."""
        X = np.ascontiguousarray(X, dtype=np.float64).reshape(-1, self.n)
        E = X.shape[0]
        n_gen = 0
        gp = None
        if generated is not None:
            generated = np.ascontiguousarray(generated, dtype=np.float64).reshape(T, -1, 12)
            n_gen = generated.shape[1]
            gp = _d(generated)
        if discard:
            lib().exo_oracle_problem_update_batch(self.h, _d(X), E, T, gp, n_gen, None, None, n_threads)
            return None, None
        phi = np.zeros((E, self.length_phi))
        jac = np.zeros((E, self.n, self.length_jacobian))
        lib().exo_oracle_problem_update_batch(self.h, _d(X), E, T, gp, n_gen, _d(phi), _d(jac), n_threads)
        return phi, np.ascontiguousarray(jac.transpose(0, 2, 1))


    def ik_solve(self, q0, rho, goal=None, W=None, max_iterations=100, tolerance=1e-5, regularization_rate=1e-9, maximum_regularization=1e9,
                 gradient_tolerance=1e-9, step_tolerance=1e-5, max_step=0.0, threshold_increase=0.1, threshold_decrease=0.5,
                 minimum_regularization=1e-12):
        """IKSolver::Solve from q0: returns (q, cost, iterations, termination)."""
        q = np.ascontiguousarray(q0, dtype=np.float64).copy()
        rho = np.ascontiguousarray(rho, dtype=np.float64)
        goal = np.ascontiguousarray(self.zero_goal() if goal is None else goal, dtype=np.float64)
        Wp = _d(np.ascontiguousarray(W, dtype=np.float64)) if W is not None else None
        cost = ctypes.c_double(0)
        iters = ctypes.c_int(0)
        lam = ctypes.c_double(0)
        term = lib().exo_oracle_ik_solve_ex(self.h, _d(q), _d(rho), _d(goal), Wp, int(max_iterations), float(tolerance), float(regularization_rate),
                                            float(maximum_regularization), float(gradient_tolerance), float(step_tolerance), float(max_step),
                                            float(threshold_increase), float(threshold_decrease), float(minimum_regularization),
                                            ctypes.byref(cost), ctypes.byref(iters), ctypes.byref(lam))
        self.last_ik_lambda = lam.value
        return q, cost.value, iters.value, term

    def problem_hessian(self, x):
        """exotica::Hessian of the problem's task maps at x: [length_jacobian][N][N]."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = np.zeros((self.length_jacobian, self.n, self.n))
        lib().exo_oracle_problem_hessian(self.h, _d(x), _d(out))
        return out

    # ---- task cost ---------------------------------------------------------
    def zero_goal(self):
        """TaskSpaceVector::SetZero (task_space_vector.cpp:54-62): zeros with identity rotation encodings."""
        y = np.zeros(self.length_phi)
        start = 0
        ident = np.array([1.0, 0, 0, 0, 1, 0, 0, 0, 1])
        for m in self.maps:
            kind = MAP_KINDS[m["type"]]
            nf = len(m.get("EndEffector", []))
            if kind in (0, 5):
                rl = ROTATION_LENGTH[ROTATION_TYPES[m.get("Type", "RPY")]]
                lead = 3 if kind == 0 else 0
                for f in range(nf):
                    y[start + f * (lead + rl) + lead:start + (f + 1) * (lead + rl)] = set_rotation(ident, m.get("Type", "RPY"))
                start += nf * (lead + rl)
            elif kind == 6:
                start += 2 * nf
            elif kind == 7:
                start += nf
            elif kind == 8:
                start += len(m.get("JointMap", range(self.n)))
            elif kind == 1:
                start += 3 * nf
            elif kind == 2:
                start += self.n
            else:
                g = len(set(f.get("Group", "default") for f in m["EndEffector"]))
                start += g * (g - 1) // 2
        return y

    def ydiff(self, phi, goal):
        out = np.zeros(self.length_jacobian)
        lib().exo_oracle_problem_ydiff(self.h, _d(np.ascontiguousarray(phi, dtype=np.float64)), _d(np.ascontiguousarray(goal, dtype=np.float64)),
                                       _d(out))
        return out

    def trajectory_cost(self, X, rho, goal=None, ct=1.0, W=None, t_first=None, generated=None):
        """Scalar cost of one trajectory X[T][N] and its per-timestep terms (AbstractTimeIndexedProblem::GetCost).
        rho[T][n_maps]; goal[1 or T][length_Phi] (None: SetZero goal); end-pose problems: T = 1, t_first = 0."""
        X = np.ascontiguousarray(X, dtype=np.float64).reshape(-1, self.n)
        T = X.shape[0]
        rho = np.ascontiguousarray(np.broadcast_to(np.asarray(rho, dtype=np.float64), (T, len(self.maps))))
        goal = self.zero_goal() if goal is None else goal
        goal = np.ascontiguousarray(goal, dtype=np.float64).reshape(-1, self.length_phi)
        if t_first is None:
            t_first = 1 if T > 1 else 0
        n_gen, gp = 0, None
        if generated is not None:
            generated = np.ascontiguousarray(generated, dtype=np.float64).reshape(T, -1, 12)
            n_gen, gp = generated.shape[1], _d(generated)
        cost_t = np.zeros(T)
        Wp = _d(np.ascontiguousarray(W, dtype=np.float64)) if W is not None else None
        total = lib().exo_oracle_trajectory_cost(self.h, _d(X), T, int(t_first), _d(rho), _d(goal), goal.shape[0], float(ct), Wp, gp, n_gen,
                                                 _d(cost_t))
        return total, cost_t

    def aico_task_messages(self, Q, rho, goal=None, ct=1.0, generated=None):
        """AICOSolver::GetTaskCosts for every timestep of one trajectory Q[T][N] (= qhat): (R[T][N][N], r[T][N], rhat[T], cost[T])."""
        Q = np.ascontiguousarray(Q, dtype=np.float64).reshape(-1, self.n)
        T = Q.shape[0]
        rho = np.ascontiguousarray(np.broadcast_to(np.asarray(rho, dtype=np.float64), (T, len(self.maps))))
        goal = self.zero_goal() if goal is None else goal
        goal = np.ascontiguousarray(goal, dtype=np.float64).reshape(-1, self.length_phi)
        n_gen, gp = 0, None
        if generated is not None:
            generated = np.ascontiguousarray(generated, dtype=np.float64).reshape(T, -1, 12)
            n_gen, gp = generated.shape[1], _d(generated)
        R, r, rhat, cost = np.zeros((T, self.n, self.n)), np.zeros((T, self.n)), np.zeros(T), np.zeros(T)
        lib().exo_oracle_aico_task_messages(self.h, _d(Q), T, _d(rho), _d(goal), goal.shape[0], float(ct), gp, n_gen, _d(R), _d(r), _d(rhat), _d(cost))
        return R, r, rhat, cost

    def trajectory_cost_jacobian(self, X, rho, goal=None, ct=1.0, W=None, t_first=None, generated=None):
        """Gradient of trajectory_cost with respect to X as grad[T][N] (AbstractTimeIndexedProblem::GetCostJacobian; row 0 is zero for
        time-indexed problems; end-pose problems: UnconstrainedEndPoseProblem::GetScalarJacobian)."""
        X = np.ascontiguousarray(X, dtype=np.float64).reshape(-1, self.n)
        T = X.shape[0]
        rho = np.ascontiguousarray(np.broadcast_to(np.asarray(rho, dtype=np.float64), (T, len(self.maps))))
        goal = self.zero_goal() if goal is None else goal
        goal = np.ascontiguousarray(goal, dtype=np.float64).reshape(-1, self.length_phi)
        if t_first is None:
            t_first = 1 if T > 1 else 0
        n_gen, gp = 0, None
        if generated is not None:
            generated = np.ascontiguousarray(generated, dtype=np.float64).reshape(T, -1, 12)
            n_gen, gp = generated.shape[1], _d(generated)
        grad = np.zeros((T, self.n))
        Wp = _d(np.ascontiguousarray(W, dtype=np.float64)) if W is not None else None
        lib().exo_oracle_trajectory_cost_jacobian(self.h, _d(X), T, int(t_first), _d(rho), _d(goal), goal.shape[0], float(ct), Wp, gp, n_gen,
                                                  _d(grad))
        return grad


def load_robot(urdf_path, srdf_path, joint_group="", add_collision_elements=False):
    with open(urdf_path) as f:
        urdf = f.read()
    with open(srdf_path) as f:
        srdf = f.read()
    return robot_model.RobotTree(urdf, srdf, joint_group, add_collision_elements)

"""Python surface mirroring the forwarded subset of `pyexotica` (exotica_python/src/pyexotica.cpp).

Every class here only forwards to the C++ host library through the C-ABI (include/exotica_b200_host.h):
no torch, no numpy math, no CPU fallback.  Method names follow pyexotica:

    Setup.load_problem / create_problem / get_maps / get_problems      pyexotica.cpp:683-706
    problem.update(x) / update(x, t), .Phi, .jacobian, .N, .T, .get_scene(), .get_tasks()   :935-1060
    scene.update / fk / jacobian / get_kinematic_tree / get_controlled_joint_names ...       :1221-1300
    tree.get_joint_limits / get_controlled_joint_names / get_random_controlled_state / set_seed   :1376-1455
    task.id / start / length / startJ / lengthJ                         :835-842

plus the batched entry points the reference does not have: problem.update_batch(X), scene.update_batch(X).
"""
import ctypes
import os
import xml.sax.saxutils as _sx

import numpy as np

from . import _capi, _host
from ._capi import ExoticaError, LAYOUT_AOS_COLMAJOR, LAYOUT_AOS_ROWMAJOR, LAYOUT_SOA

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int32)


def _check(rc):
    if rc != 0:
        raise ExoticaError(_capi.lib().exo_last_error().decode("utf-8", "replace"))


def _d(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def _s(text):
    return text.encode("utf-8") if text is not None else None


def _split(b):
    t = b.decode("utf-8") if b else ""
    return t.split("\n") if t else []


def get_frame(v):
    """exotica::GetFrame: [x y z], [x y z r p y] or [x y z qx qy qz qw] -> 12 doubles (row-major R, p)."""
    v = np.ascontiguousarray(v, dtype=np.float64)
    out = np.zeros(12)
    _check(_host.lib().exo_host_get_frame(_d(v), int(v.size), _d(out)))
    return out


def _offset(o):
    if o is None:
        return None
    o = np.ascontiguousarray(o, dtype=np.float64)
    return o if o.size == 12 else get_frame(o)


# ---- initializers -----------------------------------------------------------------------------------------
def _value_to_text(v):
    if isinstance(v, str):
        return v
    if isinstance(v, bool):
        return "1" if v else "0"
    if isinstance(v, (int, np.integer)):
        return str(int(v))
    if isinstance(v, (float, np.floating)):
        return repr(float(v))
    arr = np.asarray(v, dtype=np.float64).ravel()
    return " ".join(repr(float(x)) for x in arr)


def _is_initializer(v):
    return isinstance(v, (tuple, list)) and len(v) == 2 and isinstance(v[0], str) and isinstance(v[1], dict)


def initializer_to_xml(init):
    """("exotica/Type", {property: value}) -> XML element text, the format XMLLoader reads
    (pyexotica passes such tuples straight to C++; here they travel as XML text through the C-ABI)."""
    name, props = init
    tag = name.split("/")[-1]
    attrs, children = [], []
    for key, val in props.items():
        if _is_initializer(val):
            children.append("<%s>%s</%s>" % (key, initializer_to_xml(val), key))
        elif isinstance(val, (list, tuple)) and len(val) > 0 and all(_is_initializer(x) for x in val):
            children.append("<%s>%s</%s>" % (key, "".join(initializer_to_xml(x) for x in val), key))
        elif isinstance(val, (list, tuple)) and len(val) == 0:
            continue
        else:
            text = _value_to_text(val)
            if text.lstrip().startswith("<"):
                children.append("<%s>%s</%s>" % (key, _sx.escape(text), key))
            else:
                attrs.append("%s=%s" % (key, _sx.quoteattr(text)))
    return "<%s %s>%s</%s>" % (tag, " ".join(attrs), "".join(children), tag)


class KinematicTree(object):
    """pyexotica.KinematicTree (pyexotica.cpp:1376-1455)."""

    def __init__(self, scene):
        self._scene = scene

    def get_num_controlled_joints(self):
        return _host.lib().exo_host_scene_num_controlled(self._scene._h)

    def get_num_model_joints(self):
        return _host.lib().exo_host_scene_num_model_joints(self._scene._h)

    def get_controlled_joint_names(self):
        return _split(_host.lib().exo_host_scene_names(self._scene._h, 0))

    def get_controlled_link_names(self):
        return _split(_host.lib().exo_host_scene_names(self._scene._h, 1))

    def get_model_joint_names(self):
        return _split(_host.lib().exo_host_scene_names(self._scene._h, 2))

    def get_model_link_names(self):
        return _split(_host.lib().exo_host_scene_names(self._scene._h, 6))

    def get_root_frame_name(self):
        return _split(_host.lib().exo_host_scene_names(self._scene._h, 4))[0]

    def get_root_joint_name(self):
        return _split(_host.lib().exo_host_scene_names(self._scene._h, 5))[0]

    def get_model_base_type(self):
        out = np.zeros(2, dtype=np.int32)
        _check(_host.lib().exo_host_scene_base_types(self._scene._h, out.ctypes.data_as(_ip)))
        return {0: "FIXED", 10: "FLOATING", 20: "PLANAR"}[int(out[0])]

    def get_controlled_base_type(self):
        out = np.zeros(2, dtype=np.int32)
        _check(_host.lib().exo_host_scene_base_types(self._scene._h, out.ctypes.data_as(_ip)))
        return {0: "FIXED", 10: "FLOATING", 20: "PLANAR"}[int(out[1])]

    def get_joint_limits(self):
        n = self.get_num_controlled_joints()
        lo, hi = np.zeros(n), np.zeros(n)
        _check(_host.lib().exo_host_scene_joint_limits(self._scene._h, _d(lo), _d(hi)))
        return np.stack([lo, hi], axis=1)

    def set_joint_limits_lower(self, lower):
        lo = np.ascontiguousarray(lower, dtype=np.float64)
        if lo.size != self.get_num_controlled_joints():
            raise ExoticaError("Got %d but %d expected." % (lo.size, self.get_num_controlled_joints()))
        _check(_host.lib().exo_host_scene_set_joint_limits(self._scene._h, _d(lo), None))

    def set_joint_limits_upper(self, upper):
        hi = np.ascontiguousarray(upper, dtype=np.float64)
        if hi.size != self.get_num_controlled_joints():
            raise ExoticaError("Got %d but %d expected." % (hi.size, self.get_num_controlled_joints()))
        _check(_host.lib().exo_host_scene_set_joint_limits(self._scene._h, None, _d(hi)))

    def get_random_controlled_state(self):
        x = np.zeros(self.get_num_controlled_joints())
        _check(_host.lib().exo_host_scene_random_state(self._scene._h, _d(x)))
        return x

    def set_seed(self, seed):
        _check(_host.lib().exo_host_scene_set_seed(self._scene._h, int(seed)))

    def get_tree(self):
        """Flattened element tree: dict of arrays in tree order (index 0 = world frame)."""
        n = _host.lib().exo_host_scene_num_elements(self._scene._h)
        out = dict(id=np.zeros(n, np.int32), parent=np.zeros(n, np.int32), joint_type=np.zeros(n, np.int32),
                   control_id=np.zeros(n, np.int32), origin=np.zeros((n, 3)), axis=np.zeros((n, 3)), f_tip=np.zeros((n, 12)))
        _check(_host.lib().exo_host_scene_tree(self._scene._h, out["id"].ctypes.data_as(_ip), out["parent"].ctypes.data_as(_ip),
                                               out["joint_type"].ctypes.data_as(_ip), out["control_id"].ctypes.data_as(_ip),
                                               _d(out["origin"]), _d(out["axis"]), _d(out["f_tip"])))
        out["link_name"] = _split(_host.lib().exo_host_scene_names(self._scene._h, 3))
        out["joint_name"] = _split(_host.lib().exo_host_scene_names(self._scene._h, 7))
        return out


class Scene(object):
    """pyexotica.Scene (pyexotica.cpp:1221-1300), plus update_batch."""

    def __init__(self, handle, owner=None):
        self._h = handle
        self._owner = owner  # a problem that owns the scene, or None if we own it
        self._n_frames = 0
        self._with_jacobian = True

    @classmethod
    def create(cls, urdf, srdf="", joint_group=""):
        """urdf / srdf: XML text or file paths."""
        def load(v):
            if v and not v.lstrip().startswith("<"):
                with open(v) as f:
                    return f.read()
            return v
        h = ctypes.c_void_p()
        _check(_host.lib().exo_host_scene_create(_s(load(urdf)), _s(load(srdf)), _s(joint_group), ctypes.byref(h)))
        return cls(h.value)

    def __del__(self):
        try:
            if self._owner is None and self._h:
                _host.lib().exo_host_scene_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def get_kinematic_tree(self):
        return KinematicTree(self)

    def get_controlled_joint_names(self):
        return self.get_kinematic_tree().get_controlled_joint_names()

    def get_controlled_link_names(self):
        return self.get_kinematic_tree().get_controlled_link_names()

    def get_model_joint_names(self):
        return self.get_kinematic_tree().get_model_joint_names()

    def get_model_link_names(self):
        return self.get_kinematic_tree().get_model_link_names()

    def get_root_frame_name(self):
        return self.get_kinematic_tree().get_root_frame_name()

    def get_root_joint_name(self):
        return self.get_kinematic_tree().get_root_joint_name()

    @property
    def num_positions(self):
        return self.get_kinematic_tree().get_num_controlled_joints()

    def get_model_state(self):
        x = np.zeros(_host.lib().exo_host_scene_num_model_joints(self._h))
        _check(_host.lib().exo_host_scene_get_model_state(self._h, _d(x)))
        return x

    def set_model_state(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        _check(_host.lib().exo_host_scene_set_model_state(self._h, _d(x), int(x.size)))

    def add_object(self, name, transform=None, parent=""):
        _check(_host.lib().exo_host_scene_add_object(self._h, _s(name), _d(_offset(transform)), _s(parent)))

    def load_scene(self, scene_text, offset=None):
        """Scene::LoadScene: world objects from MoveIt ".scene" geometry text (each becomes a named frame)."""
        _check(_host.lib().exo_host_scene_load_scene(self._h, _s(scene_text), _d(_offset(offset))))

    def load_scene_file(self, file_name, offset=None):
        """Scene::LoadSceneFile ({package} paths resolve through Setup.set_package_path)."""
        _check(_host.lib().exo_host_scene_load_scene_file(self._h, _s(file_name), _d(_offset(offset))))

    def clean_scene(self):
        _check(_host.lib().exo_host_scene_clean(self._h))

    def add_trajectory(self, link, frames):
        """Scene::AddTrajectory: `frames` is either the reference's trajectory text ("radius\\n rows cols\\n time x y z [rot]...",
        sampled at t * tau by time-indexed problems) or an explicit table of [T][12] frames / [T][3|6|7] GetFrame vectors."""
        if isinstance(frames, str):
            _check(_host.lib().exo_host_scene_add_trajectory_string(self._h, _s(link), _s(frames)))
            return
        fr = np.array([_offset(f) for f in frames], dtype=np.float64)
        _check(_host.lib().exo_host_scene_add_trajectory(self._h, _s(link), fr.shape[0], _d(fr)))

    def get_trajectory(self, link):
        """Scene::GetTrajectory: the Trajectory attached to `link` (a copy built from its text form)."""
        buf = ctypes.create_string_buffer(1 << 20)
        _check(_host.lib().exo_host_scene_get_trajectory(self._h, _s(link), buf, len(buf)))
        return Trajectory(buf.value.decode())

    def request_frames(self, frames, with_jacobian=True):
        """frames: list of dicts with the Frame initializer fields (Link, LinkOffset, Base, BaseOffset)."""
        n = len(frames)
        la = "\n".join(f.get("Link", "") for f in frames)
        lb = "\n".join(f.get("Base", "") for f in frames)
        ident = np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0], dtype=np.float64)
        oa = np.array([_offset(f.get("LinkOffset")) if f.get("LinkOffset") is not None else ident for f in frames]).reshape(n, 12)
        ob = np.array([_offset(f.get("BaseOffset")) if f.get("BaseOffset") is not None else ident for f in frames]).reshape(n, 12)
        _check(_host.lib().exo_host_scene_request_frames(self._h, n, _s(la), _d(oa), _s(lb), _d(ob), 1 if with_jacobian else 0))
        self._n_frames, self._with_jacobian = n, with_jacobian

    def update(self, x, t=0.0):
        """Scene::Update(x, t): returns (Phi[F][12], J[F][6][N]) of the requested frames."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        n = self.num_positions
        phi = np.zeros((self._n_frames, 12))
        jac = np.zeros((self._n_frames, n, 6))
        _check(_host.lib().exo_host_scene_update(self._h, _d(x), int(x.size), float(t), _d(phi), _d(jac) if self._with_jacobian else None))
        return phi, (np.ascontiguousarray(jac.transpose(0, 2, 1)) if self._with_jacobian else None)

    def update_batch(self, X, layout=LAYOUT_AOS_ROWMAJOR):
        """Batched KinematicTree::Update: X[E][N] -> Phi[E][F*12], J[E][F*6][N] (AoS row-major) or the SoA forms."""
        X = np.ascontiguousarray(X, dtype=np.float64)
        n, F = self.num_positions, self._n_frames
        if layout == LAYOUT_SOA:
            if X.ndim != 2 or X.shape[0] != n:
                raise ExoticaError("Wrong state vector size! Got %s expected (%d, E)" % (X.shape, n))
            E = X.shape[1]
            phi, jac = np.empty((12 * F, E)), np.empty((6 * F, n, E))
        else:
            if X.ndim != 2 or X.shape[1] != n:
                raise ExoticaError("Wrong state vector size! Got %s expected (E, %d)" % (X.shape, n))
            E = X.shape[0]
            phi = np.empty((E, 12 * F))
            jac = np.empty((E, 6 * F, n)) if layout == LAYOUT_AOS_ROWMAJOR else np.empty((E, n, 6 * F))
        if not self._with_jacobian:
            jac = None
        _check(_host.lib().exo_host_scene_update_batch(self._h, X.ctypes.data, E, int(layout), phi.ctypes.data,
                                                       jac.ctypes.data if jac is not None else None))
        return phi, jac

    def fk(self, link_a, offset_a=None, link_b="", offset_b=None):
        out = np.zeros(12)
        _check(_host.lib().exo_host_scene_fk(self._h, _s(link_a), _d(_offset(offset_a)), _s(link_b), _d(_offset(offset_b)), _d(out)))
        return out

    def jacobian(self, link_a, offset_a=None, link_b="", offset_b=None):
        n = self.num_positions
        out = np.zeros((n, 6))
        _check(_host.lib().exo_host_scene_jacobian(self._h, _s(link_a), _d(_offset(offset_a)), _s(link_b), _d(_offset(offset_b)), _d(out)))
        return np.ascontiguousarray(out.T)


class TaskMapView(object):
    """pyexotica.TaskMap indexing fields (pyexotica.cpp:835-842)."""

    def __init__(self, name, idx):
        self.name = name
        self.id, self.start, self.length, self.startJ, self.lengthJ = (int(v) for v in idx[:5])
        self.is_used = bool(idx[5])


class PlanningProblem(object):
    """pyexotica.PlanningProblem subset + update_batch."""

    def __init__(self, handle):
        self._h = handle
        d = np.zeros(4, dtype=np.int32)
        _check(_host.lib().exo_host_problem_dims(handle, d.ctypes.data_as(_ip)))
        self.N, self.length_Phi, self.length_jacobian, self.T = (int(v) for v in d)
        self.type = _host.lib().exo_host_problem_type(handle).decode()
        self._scene = Scene(_host.lib().exo_host_problem_scene(handle), owner=self)
        self.time_indexed = "TimeIndexed" in self.type
        if self.time_indexed:
            self.Phi = [np.zeros(self.length_Phi) for _ in range(self.T)]
            self.jacobian = [np.zeros((self.length_jacobian, self.N)) for _ in range(self.T)]
        else:
            self.Phi = np.zeros(self.length_Phi)
            self.jacobian = np.zeros((self.length_jacobian, self.N))

    def __del__(self):
        try:
            if self._h:
                _host.lib().exo_host_problem_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def get_scene(self):
        return self._scene

    def get_tasks(self):
        out = []
        for i in range(_host.lib().exo_host_problem_num_tasks(self._h)):
            idx = np.zeros(6, dtype=np.int32)
            _check(_host.lib().exo_host_problem_task_indexing(self._h, i, idx.ctypes.data_as(_ip)))
            out.append(TaskMapView(_host.lib().exo_host_problem_task_name(self._h, i).decode(), idx))
        return out

    def get_task_maps(self):
        return {t.name: t for t in self.get_tasks()}

    @property
    def start_state(self):
        x = np.zeros(self.N)
        _check(_host.lib().exo_host_problem_start_state(self._h, _d(x)))
        return x

    @property
    def number_of_problem_updates(self):
        return int(_host.lib().exo_host_problem_update_count(self._h))

    def update(self, x, t=None):
        """UnconstrainedEndPoseProblem.update(x) / TimeIndexedProblem.update(x, t): one evaluation on the GPU."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        if self.time_indexed and t is None:
            # AbstractTimeIndexedProblem::Update(x_trajectory) (abstract_time_indexed_problem.cpp:209-218): (T-1)*N values,
            # timesteps 1..T-1 are evaluated, t = 0 keeps the start state's values
            if x.size != (self.T - 1) * self.N:
                raise ExoticaError("To update using the trajectory Update method, please use a trajectory of size N x (T-1) (%d), given: %d"
                                   % (self.N * (self.T - 1), x.size))
            phi = np.zeros((self.T, self.length_Phi))
            jac = np.zeros((self.T, self.N, self.length_jacobian))
            _check(_host.lib().exo_host_problem_update_trajectory(self._h, _d(x), int(x.size), _d(phi), _d(jac)))
            self.Phi = [phi[k] for k in range(self.T)]
            self.jacobian = [np.ascontiguousarray(jac[k].T) for k in range(self.T)]
            return
        phi = np.zeros(self.length_Phi)
        jac = np.zeros((self.N, self.length_jacobian))
        _check(_host.lib().exo_host_problem_update(self._h, _d(x), int(x.size), -1 if t is None else int(t), _d(phi), _d(jac)))
        if self.time_indexed:
            tt = int(t) if t >= 0 else self.T - 1
            self.Phi[tt] = phi
            self.jacobian[tt] = np.ascontiguousarray(jac.T)
        else:
            self.Phi = phi
            self.jacobian = np.ascontiguousarray(jac.T)

    def _shapes(self, X, layout):
        N, LP, LJ = self.N, self.length_Phi, self.length_jacobian
        if layout == LAYOUT_SOA:
            if X.ndim != 2 or X.shape[0] != N:
                raise ExoticaError("Wrong state vector size! Got %s expected (%d, E)" % (X.shape, N))
            return X.shape[1], (LP, X.shape[1]), (LJ, N, X.shape[1])
        if X.ndim == 3:  # [B][T][N]
            X = X.reshape(-1, X.shape[-1])
        if X.ndim != 2 or X.shape[1] != N:
            raise ExoticaError("Wrong state vector size! Got %s expected (E, %d)" % (X.shape, N))
        E = X.shape[0]
        return E, (E, LP), ((E, LJ, N) if layout == LAYOUT_AOS_ROWMAJOR else (E, N, LJ))

    def update_batch(self, X, layout=LAYOUT_AOS_ROWMAJOR, out_phi=None, out_jacobian=None):
        """Fill Phi and J for all evaluations in one call.  X[E][N] (or [B][T][N] for time-indexed problems,
        E = B*T, timestep innermost) -> Phi[E][length_Phi], J[E][length_jacobian][N]."""
        X = np.ascontiguousarray(X, dtype=np.float64)
        E, pshape, jshape = self._shapes(X, layout)
        phi = _capi.checked_output(out_phi, pshape, "out_phi") if out_phi is not None else np.empty(pshape)
        jac = _capi.checked_output(out_jacobian, jshape, "out_jacobian") if out_jacobian is not None else np.empty(jshape)
        _check(_host.lib().exo_host_problem_update_batch(self._h, X.ctypes.data, E, int(layout), phi.ctypes.data, jac.ctypes.data))
        return phi, jac

    def update_batch_device(self, d_x, n_evals, d_phi, d_jacobian, layout=LAYOUT_SOA, stream=None):
        """Same, on device pointers already resident in HBM (asynchronous on `stream`)."""
        _check(_host.lib().exo_host_problem_update_batch_device(self._h, d_x, int(n_evals), int(layout), d_phi, d_jacobian, stream))

    def cost_batch(self, X):
        """Scalar cost of every trajectory: X[B][T][N] (end-pose problems: X[B][N]) -> cost[B].  Time-indexed problems:
        AbstractTimeIndexedProblem.get_cost() of each trajectory (sum over t = 1..T-1 of task + transition cost); end-pose
        problems: get_scalar_cost().  Only X crosses PCIe in, 8 bytes per trajectory come back."""
        X = np.ascontiguousarray(X, dtype=np.float64)
        want = (self.T, self.N) if self.time_indexed else (self.N,)
        if X.ndim != len(want) + 1 or tuple(X.shape[1:]) != want:
            raise ExoticaError("Wrong state vector size! Got %s expected (B, %s)" % (X.shape, ", ".join(str(v) for v in want)))
        cost = np.empty(X.shape[0])
        _check(_host.lib().exo_host_problem_cost_batch(self._h, X.ctypes.data, X.shape[0], cost.ctypes.data))
        return cost

    def cost_jacobian_batch(self, X, out_cost=None, out_jacobian=None):
        """(cost[B], jacobian like X): the gradient of cost_batch with respect to every state.  Time-indexed problems:
        AbstractTimeIndexedProblem.get_cost_jacobian() of each trajectory reshaped to [T][N] with a zero row for t = 0; end-pose
        problems: get_scalar_jacobian().  Phi and J stay on the device: N doubles per evaluation come back.  out_cost /
        out_jacobian: caller-owned result arrays (C-contiguous float64, shapes (B,) and X.shape) -- pinned ones (PinnedArray(...).array,
        for X too) let the chunks' copies run asynchronously at link speed instead of through the driver's pageable staging."""
        X = np.ascontiguousarray(X, dtype=np.float64)
        want = (self.T, self.N) if self.time_indexed else (self.N,)
        if X.ndim != len(want) + 1 or tuple(X.shape[1:]) != want:
            raise ExoticaError("Wrong state vector size! Got %s expected (B, %s)" % (X.shape, ", ".join(str(v) for v in want)))
        for name, arr, shape in (("out_cost", out_cost, (X.shape[0],)), ("out_jacobian", out_jacobian, X.shape)):
            if arr is not None and (not isinstance(arr, np.ndarray) or arr.dtype != np.float64 or not arr.flags.c_contiguous or arr.shape != tuple(shape)):
                raise ExoticaError("%s must be a C-contiguous float64 array of shape %s" % (name, tuple(shape)))
        cost = np.empty(X.shape[0]) if out_cost is None else out_cost
        jac = np.empty_like(X) if out_jacobian is None else out_jacobian
        _check(_host.lib().exo_host_problem_cost_jacobian_batch(self._h, X.ctypes.data, X.shape[0], cost.ctypes.data, jac.ctypes.data))
        return cost, jac

    def hessian_batch(self, X):
        """exotica.Hessian of every evaluation: X[E][N] -> H[E][length_jacobian][N][N] (KinematicTree::ComputeH and the task maps'
        Hessian overloads)."""
        X = np.ascontiguousarray(X, dtype=np.float64).reshape(-1, self.N)
        H = np.empty((X.shape[0], self.length_jacobian, self.N, self.N))
        _check(_host.lib().exo_host_problem_hessian_batch(self._h, X.ctypes.data, X.shape[0], H.ctypes.data))
        return H

    def cost_batch_device(self, d_x, d_phi, n_trajectories, d_cost, d_cost_eval=None, layout=LAYOUT_SOA, stream=None):
        """Same on the device buffers of update_batch_device (asynchronous on `stream`)."""
        _check(_host.lib().exo_host_problem_cost_batch_device(self._h, d_x, d_phi, int(n_trajectories), int(layout), d_cost, d_cost_eval, stream))

    def update_cost_batch_device(self, d_x, n_trajectories, d_phi, d_jacobian, d_cost, d_cost_eval=None, layout=LAYOUT_SOA, stream=None):
        """update_batch_device and cost_batch_device in one call: Update(x, t) of every timestep and get_cost() of every trajectory.
        SoA buffers of problems whose cost is made of EffFrame / EffPosition / JointLimit maps run one generated kernel (the cost of
        an evaluation is formed while its Phi is still in registers) plus the per-trajectory sum."""
        _check(_host.lib().exo_host_problem_update_cost_batch_device(self._h, d_x, int(n_trajectories), int(layout), d_phi, d_jacobian, d_cost, d_cost_eval,
                                                                     stream))

    def last_kernel_ms(self):
        ms = ctypes.c_float(0)
        _check(_host.lib().exo_host_problem_last_kernel_ms(self._h, ctypes.byref(ms)))
        return ms.value

    def launch_count(self):
        return int(_host.lib().exo_host_problem_launch_count(self._h))

    def last_kernel(self):
        """Name of the CUDA kernel the most recent batched update launched (bench.py's roofline names it)."""
        from . import _capi
        return _capi.lib().exo_problem_last_kernel(_host.lib().exo_host_problem_backend(self._h)).decode()


class Trajectory(object):
    """exotica::Trajectory (exotica_core/src/trajectory.cpp): get_position(t) -> frame[12] (rotation row-major, then position),
    get_duration().  Built from the reference's text format or from (data[n][4|7|8], radius)."""

    def __init__(self, data, radius=1.0):
        if not isinstance(data, str):
            data = np.asarray(data, dtype=np.float64)
            data = "%r\n%d %d\n%s" % (float(radius), data.shape[0], data.shape[1], "\n".join(" ".join(repr(float(v)) for v in row) for row in data))
        self.text = data
        self._duration = None

    def get_position(self, t):
        times = np.atleast_1d(np.asarray(t, dtype=np.float64))
        out = np.empty((times.size, 12))
        dur = ctypes.c_double(0.0)
        _check(_host.lib().exo_host_trajectory_sample(_s(self.text), int(times.size), _d(times), _d(out), ctypes.byref(dur)))
        self._duration = dur.value
        return out if np.ndim(t) else out[0]

    def get_duration(self):
        if self._duration is None:
            self.get_position(0.0)
        return self._duration


class IKSolver(object):
    """exotica_ik_solver's IKSolver (exotations/solvers/exotica_ik_solver): specify_problem / solve as in pyexotica, plus
    solve_batch(Q0) for multi-start IK -- every start state is solved independently on the GPU."""

    TERMINATION = {0: "IterationLimit", 2: "StepTolerance", 3: "FunctionTolerance", 4: "GradientTolerance", 5: "Divergence"}

    def __init__(self, props=None):
        props = dict(props or {})
        self.max_iterations = int(props.get("MaxIterations", 100))                      # exotica_core/init/motion_solver.in
        self.tolerance = float(props.get("Tolerance", 1e-5))                            # init/ik_solver.in
        self.regularization_rate = float(props.get("RegularizationRate", 1e-9))
        self.maximum_regularization = float(props.get("MaximumRegularization", 1e9))
        self.gradient_tolerance = float(props.get("GradientToleranceConvergenceThreshold", 1e-9))
        self.step_tolerance = float(props.get("StepToleranceConvergenceThreshold", 1e-5))
        self.max_step = float(props.get("MaxStep", 0.0))  # only used if MaxIterations == 1 (ik_solver.in)
        self.minimum_regularization = float(props.get("MinimumRegularization", 1e-12))
        self.threshold_increase = float(props.get("ThresholdRegularizationIncrease", 0.1))
        self.threshold_decrease = float(props.get("ThresholdRegularizationDecrease", 0.5))
        self.problem = None

    def specify_problem(self, problem):
        if problem.time_indexed:
            raise ExoticaError("This IKSolver can't solve problem of type '%s'!" % problem.type)
        self.problem = problem

    def solve_batch(self, Q0):
        """Q0[E][N] -> (solutions[E][N], cost[E], iterations[E], termination[E])."""
        if self.problem is None:
            raise ExoticaError("Solver has not been initialized!")
        Q0 = np.ascontiguousarray(Q0, dtype=np.float64).reshape(-1, self.problem.N)
        E = Q0.shape[0]
        prm = _capi.IkParams(self.max_iterations, self.tolerance, self.regularization_rate, self.maximum_regularization,
                             self.gradient_tolerance, self.step_tolerance, self.max_step, self.threshold_increase, self.threshold_decrease,
                             self.minimum_regularization)
        Q, cost = np.empty_like(Q0), np.empty(E)
        it, term = np.zeros(E, dtype=np.int32), np.zeros(E, dtype=np.int32)
        _check(_host.lib().exo_host_problem_ik_solve_batch(self.problem._h, Q0.ctypes.data, E, ctypes.byref(prm), Q.ctypes.data, cost.ctypes.data,
                                                           it.ctypes.data, term.ctypes.data))
        return Q, cost, it, term

    def solve(self):
        """pyexotica: solution = solver.solve() from the problem's start state -> [1][N]."""
        Q, cost, it, term = self.solve_batch(self.problem.start_state[None, :])
        self.termination_criterion = self.TERMINATION.get(int(term[0]), str(int(term[0])))
        return Q


class Setup(object):
    """pyexotica.Setup (pyexotica.cpp:683-706)."""

    @staticmethod
    def create_solver(initializer):
        name, props = initializer
        if name.split("/")[-1] != "IKSolver":
            raise ExoticaError("Unable to load solver of type '%s' (available on the GPU path: exotica/IKSolver)" % name)
        return IKSolver(props)

    @staticmethod
    def set_package_path(package, directory):
        _check(_host.lib().exo_host_set_package_path(_s(package), _s(directory)))

    @staticmethod
    def get_maps():
        return _split(_host.lib().exo_host_registered_maps())

    @staticmethod
    def get_problems():
        return _split(_host.lib().exo_host_registered_problems())

    @staticmethod
    def load_problem(file_or_xml):
        text = file_or_xml
        if not text.lstrip().startswith("<"):
            if not os.path.exists(text):
                raise ExoticaError("File does not exist: '%s'" % text)
            with open(text) as f:
                text = f.read()
        h = ctypes.c_void_p()
        _check(_host.lib().exo_host_load_problem_xml(_s(text), ctypes.byref(h)))
        return PlanningProblem(h.value)

    @staticmethod
    def create_problem(initializer):
        """initializer: ("exotica/<ProblemType>", {...}) as in exotica_examples/scripts/example_ik_manual_initialization."""
        return Setup.load_problem(initializer_to_xml(initializer))

    @staticmethod
    def create_scene(initializer):
        name, props = initializer
        return Scene.create(props["URDF"], props.get("SRDF", ""), props.get("JointGroup", ""))

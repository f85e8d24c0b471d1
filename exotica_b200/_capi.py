"""ctypes binding of the C-ABI (include/exotica_b200.h).  Pure forwarding: no torch, no CPU fallback.

Loading fails loudly if exotica_b200/lib/libexotica_b200.so is missing (build it with
`make -C exotica_b200/csrc` or `python -c "import __graft_entry__ as g; g.build()"`).
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("EXO_B200_LIB") or os.path.join(_HERE, "lib", "libexotica_b200.so")  # env override: kernel tuning builds

EXO_OK = 0
LAYOUT_SOA, LAYOUT_AOS_ROWMAJOR, LAYOUT_AOS_COLMAJOR = 0, 1, 2
MAP_EFF_FRAME, MAP_EFF_POSITION, MAP_JOINT_LIMIT, MAP_SPHERE_COLLISION, MAP_FRAMES = 0, 1, 2, 3, 4
MAP_EFF_ORIENTATION, MAP_EFF_POSITION_XY, MAP_DISTANCE, MAP_JOINT_POSE = 5, 6, 7, 8
MAP_EFF_AXIS_ALIGNMENT, MAP_CENTER_OF_MASS = 9, 10
ROTATION_TYPES = {"Quaternion": 0, "RPY": 1, "ZYX": 2, "ZYZ": 3, "AngleAxis": 4, "Matrix": 5}
ROTATION_LENGTH = [4, 3, 3, 3, 3, 9]

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int32)
_vp = ctypes.c_void_p


class TreeDesc(ctypes.Structure):
    _fields_ = [("n_elements", ctypes.c_int32), ("id", _ip), ("parent", _ip), ("joint_type", _ip), ("joint_origin", _dp),
                ("joint_axis", _dp), ("f_tip", _dp), ("control_id", _ip), ("mimic_element", _ip), ("mimic_multiplier", _dp),
                ("mimic_offset", _dp), ("generated_slot", _ip), ("n_controlled", ctypes.c_int32), ("lower", _dp), ("upper", _dp)]


class TaskMapDesc(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int32), ("n_frames", ctypes.c_int32), ("elem_a", _ip), ("off_a", _dp), ("elem_b", _ip),
                ("off_b", _dp), ("rotation_type", ctypes.c_int32), ("safe_percentage", ctypes.c_double),
                ("precision", ctypes.c_double), ("radius", _dp), ("group", _ip), ("is_used", ctypes.c_int32),
                ("n_joint_map", ctypes.c_int32), ("joint_map", _ip), ("joint_ref", _dp), ("axis", _dp), ("direction", _dp), ("mass", _dp),
                ("enable_z", ctypes.c_int32)]


class CostDesc(ctypes.Structure):
    _fields_ = [("T", ctypes.c_int32), ("T_goal", ctypes.c_int32), ("rho", _dp), ("goal", _dp), ("ct", ctypes.c_double), ("W", _dp)]


class IkParams(ctypes.Structure):
    _fields_ = [("max_iterations", ctypes.c_int32), ("tolerance", ctypes.c_double), ("regularization_rate", ctypes.c_double),
                ("maximum_regularization", ctypes.c_double), ("gradient_tolerance", ctypes.c_double), ("step_tolerance", ctypes.c_double),
                ("max_step", ctypes.c_double), ("threshold_regularization_increase", ctypes.c_double),
                ("threshold_regularization_decrease", ctypes.c_double), ("minimum_regularization", ctypes.c_double)]


class ExoticaError(RuntimeError):
    """Mirrors exotica::Exception (exotica_core/include/exotica_core/tools/exception.h:36-47)."""


# every symbol include/exotica_b200.h declares: (name, restype, argtypes)
SYMBOLS = [
    ("exo_last_error", ctypes.c_char_p, []),
    ("exo_device_count", ctypes.c_int, [ctypes.POINTER(ctypes.c_int)]),
    ("exo_set_device", ctypes.c_int, [ctypes.c_int]),
    ("exo_version", ctypes.c_char_p, []),
    ("exo_device_alloc", ctypes.c_int, [ctypes.POINTER(_vp), ctypes.c_size_t]),
    ("exo_device_free", ctypes.c_int, [_vp]),
    ("exo_host_alloc_pinned", ctypes.c_int, [ctypes.POINTER(_vp), ctypes.c_size_t]),
    ("exo_host_free_pinned", ctypes.c_int, [_vp]),
    ("exo_memcpy_h2d", ctypes.c_int, [_vp, _vp, ctypes.c_size_t, _vp]),
    ("exo_memcpy_d2h", ctypes.c_int, [_vp, _vp, ctypes.c_size_t, _vp]),
    ("exo_memset_device", ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_size_t, _vp]),
    ("exo_stream_create", ctypes.c_int, [ctypes.POINTER(_vp)]),
    ("exo_stream_destroy", ctypes.c_int, [_vp]),
    ("exo_stream_synchronize", ctypes.c_int, [_vp]),
    ("exo_device_synchronize", ctypes.c_int, []),
    ("exo_tree_create", ctypes.c_int, [ctypes.POINTER(TreeDesc), ctypes.POINTER(_vp)]),
    ("exo_tree_destroy", ctypes.c_int, [_vp]),
    ("exo_tree_set_model_state", ctypes.c_int, [_vp, _dp]),
    ("exo_tree_set_joint_limits", ctypes.c_int, [_vp, _dp, _dp]),
    ("exo_tree_num_controlled", ctypes.c_int, [_vp]),
    ("exo_problem_create", ctypes.c_int, [_vp, ctypes.POINTER(_vp)]),
    ("exo_problem_destroy", ctypes.c_int, [_vp]),
    ("exo_problem_add_task_map", ctypes.c_int, [_vp, ctypes.POINTER(TaskMapDesc)]),
    ("exo_problem_finalize", ctypes.c_int, [_vp]),
    ("exo_problem_length_phi", ctypes.c_int, [_vp]),
    ("exo_problem_length_jacobian", ctypes.c_int, [_vp]),
    ("exo_problem_num_maps", ctypes.c_int, [_vp]),
    ("exo_problem_map_indexing", ctypes.c_int, [_vp, ctypes.c_int, _ip]),
    ("exo_problem_set_generated", ctypes.c_int, [_vp, ctypes.c_int32, ctypes.c_int32, _dp]),
    ("exo_problem_update_batch", ctypes.c_int, [_vp, _vp, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, _vp, _vp, _vp]),
    ("exo_problem_update_batch_host", ctypes.c_int, [_vp, _vp, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, _vp, _vp]),
    ("exo_problem_launch_count", ctypes.c_int64, [_vp]),
    ("exo_problem_last_kernel_ms", ctypes.c_int, [_vp, ctypes.POINTER(ctypes.c_float)]),
    ("exo_problem_program_info", ctypes.c_int, [_vp, _ip]),
    ("exo_problem_chain_info", ctypes.c_int, [_vp, _ip]),
    ("exo_problem_set_kernel", ctypes.c_int, [_vp, ctypes.c_int]),
    ("exo_problem_generated_kernel_source", ctypes.c_int, [_vp, ctypes.c_char_p, ctypes.c_int64, ctypes.POINTER(ctypes.c_int64), ctypes.c_int32]),
    ("exo_problem_set_cost", ctypes.c_int, [_vp, ctypes.POINTER(CostDesc)]),
    ("exo_problem_cost_batch", ctypes.c_int, [_vp, _vp, _vp, ctypes.c_int64, ctypes.c_int32, _vp, _vp, _vp]),
    ("exo_problem_cost_batch_host", ctypes.c_int, [_vp, _vp, ctypes.c_int64, ctypes.c_int32, _vp]),
    ("exo_problem_cost_jacobian_batch", ctypes.c_int, [_vp, _vp, _vp, _vp, ctypes.c_int64, ctypes.c_int32, _vp, _vp]),
    ("exo_problem_aico_task_messages_batch", ctypes.c_int, [_vp, _vp, _vp, _vp, ctypes.c_int64, ctypes.c_int32, _vp, _vp, _vp, _vp, _vp]),
    ("exo_problem_cost_gradient_batch", ctypes.c_int, [_vp, _vp, ctypes.c_int64, ctypes.c_int32, _vp, _vp, _vp]),
    ("exo_problem_update_cost_batch", ctypes.c_int, [_vp, _vp, ctypes.c_int64, ctypes.c_int32, _vp, _vp, _vp, _vp, _vp]),
    ("exo_problem_cost_jacobian_batch_host", ctypes.c_int, [_vp, _vp, ctypes.c_int64, ctypes.c_int32, _vp, _vp]),
    ("exo_problem_set_min_distance_output", ctypes.c_int, [_vp, _vp]),
    ("exo_problem_ik_solve_batch", ctypes.c_int, [_vp, _vp, ctypes.c_int64, ctypes.POINTER(IkParams), _vp, _vp, _vp, _vp]),
    ("exo_problem_hessian_batch", ctypes.c_int, [_vp, _vp, _vp, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, _vp, _vp]),
    ("exo_trajectory_reduce", ctypes.c_int, [_vp, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, _vp, _vp]),
    ("exo_problem_last_kernel", ctypes.c_char_p, [_vp]),
    ("exo_event_create", ctypes.c_int, [ctypes.POINTER(_vp)]),
    ("exo_event_destroy", ctypes.c_int, [_vp]),
    ("exo_event_record", ctypes.c_int, [_vp, _vp]),
    ("exo_event_synchronize", ctypes.c_int, [_vp]),
    ("exo_event_elapsed_ms", ctypes.c_int, [_vp, _vp, ctypes.POINTER(ctypes.c_float)]),
    ("exo_graph_begin_capture", ctypes.c_int, [_vp]),
    ("exo_graph_end_capture", ctypes.c_int, [_vp, ctypes.POINTER(_vp)]),
    ("exo_graph_launch", ctypes.c_int, [_vp, _vp]),
    ("exo_graph_destroy", ctypes.c_int, [_vp]),
    ("exo_comm_nccl_version", ctypes.c_int, [ctypes.POINTER(ctypes.c_int)]),
    ("exo_comm_get_unique_id", ctypes.c_int, [ctypes.c_char_p]),
    ("exo_comm_create", ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_char_p, ctypes.POINTER(_vp)]),
    ("exo_comm_destroy", ctypes.c_int, [_vp]),
    ("exo_comm_rank", ctypes.c_int, [_vp]),
    ("exo_comm_world_size", ctypes.c_int, [_vp]),
    ("exo_comm_collective_count", ctypes.c_int64, [_vp]),

    ("exo_comm_peer_collective_count", ctypes.c_int64, [_vp]),
    ("exo_comm_allgather", ctypes.c_int, [_vp, _vp, _vp, ctypes.c_int64, _vp]),
    ("exo_comm_allreduce", ctypes.c_int, [_vp, _vp, _vp, ctypes.c_int64, ctypes.c_int32, _vp]),
    ("exo_comm_barrier", ctypes.c_int, [_vp, _vp]),
    ("exo_problem_cost_allgather", ctypes.c_int, [_vp, _vp, _vp, ctypes.c_int64, ctypes.c_int64, ctypes.c_int32, _vp, _vp]),
    ("exo_min_distance_allgather", ctypes.c_int, [_vp, _vp, ctypes.c_int64, ctypes.c_int32, ctypes.c_int64, _vp, _vp, _vp]),
]
COMM_ID_BYTES = 128
REDUCE_MIN, REDUCE_SUM, REDUCE_MAX = 0, 1, 2

_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ExoticaError("CUDA extension missing: %s (run `make -C exotica_b200/csrc`); there is no CPU fallback" % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, res, args in SYMBOLS:
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def check(rc):
    if rc != EXO_OK:
        raise ExoticaError(lib().exo_last_error().decode("utf-8", "replace"))


def _d(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def _i(a):
    return a.ctypes.data_as(_ip) if a is not None else None


def checked_output(arr, shape, name):
    """A caller-supplied output array goes to C as a bare pointer: it must be float64, C-contiguous, writeable and of the exact
    shape, or the library would scribble over foreign memory / write through a temporary copy."""
    if not isinstance(arr, np.ndarray):
        raise ExoticaError("%s must be a numpy array" % name)
    if arr.dtype != np.float64:
        raise ExoticaError("%s must be float64, got %s" % (name, arr.dtype))
    if tuple(arr.shape) != tuple(int(v) for v in shape):
        raise ExoticaError("%s has shape %s, expected %s" % (name, tuple(arr.shape), tuple(shape)))
    if not arr.flags["C_CONTIGUOUS"] or not arr.flags["WRITEABLE"]:
        raise ExoticaError("%s must be C-contiguous and writeable" % name)
    return arr


def device_count():
    n = ctypes.c_int(0)
    rc = lib().exo_device_count(ctypes.byref(n))
    return n.value if rc == EXO_OK else 0


class DeviceBuffer(object):
    """Raw device allocation of doubles."""

    def __init__(self, n_doubles):
        self.n = int(n_doubles)
        p = _vp()
        check(lib().exo_device_alloc(ctypes.byref(p), self.n * 8))
        self.ptr = p.value

    def upload(self, arr):
        arr = np.ascontiguousarray(arr, dtype=np.float64)
        assert arr.size <= self.n
        check(lib().exo_memcpy_h2d(self.ptr, arr.ctypes.data, arr.size * 8, None))
        check(lib().exo_stream_synchronize(None))

    def download(self, shape, offset=0):
        """offset: first double to read (for sampling slices of a large buffer)."""
        out = np.empty(shape, dtype=np.float64)
        assert offset >= 0 and offset + out.size <= self.n
        check(lib().exo_memcpy_d2h(out.ctypes.data, self.ptr + 8 * int(offset), out.size * 8, None))
        check(lib().exo_stream_synchronize(None))
        return out

    def fill_bytes(self, value):
        check(lib().exo_memset_device(self.ptr, value, self.n * 8, None))

    def free(self):
        if self.ptr:
            lib().exo_device_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Stream(object):
    """A non-blocking CUDA stream (exo_stream_create)."""

    def __init__(self):
        p = _vp()
        check(lib().exo_stream_create(ctypes.byref(p)))
        self.ptr = _vp(p.value)

    def synchronize(self):
        check(lib().exo_stream_synchronize(self.ptr))

    def __del__(self):
        try:
            if self.ptr:
                lib().exo_stream_destroy(self.ptr)
                self.ptr = None
        except Exception:
            pass


class Event(object):
    """A CUDA event for device-side timing on the launching stream."""

    def __init__(self):
        p = _vp()
        check(lib().exo_event_create(ctypes.byref(p)))
        self.ptr = _vp(p.value)

    def record(self, stream=None):
        check(lib().exo_event_record(self.ptr, stream.ptr if isinstance(stream, Stream) else stream))

    def synchronize(self):
        check(lib().exo_event_synchronize(self.ptr))

    def elapsed_ms(self, later):
        ms = ctypes.c_float(0)
        check(lib().exo_event_elapsed_ms(self.ptr, later.ptr, ctypes.byref(ms)))
        return ms.value

    def __del__(self):
        try:
            if self.ptr:
                lib().exo_event_destroy(self.ptr)
                self.ptr = None
        except Exception:
            pass


class Graph(object):
    """Stream capture into a CUDA graph: `with Graph.capture(stream) as g: ...calls on stream...`, then g.launch(stream)."""

    def __init__(self):
        self.ptr = None
        self._stream = None

    @classmethod
    def capture(cls, stream):
        g = cls()
        g._stream = stream
        return g

    def __enter__(self):
        check(lib().exo_graph_begin_capture(self._stream.ptr))
        return self

    def __exit__(self, exc_type, exc, tb):
        p = _vp()
        rc = lib().exo_graph_end_capture(self._stream.ptr, ctypes.byref(p))
        if exc_type is None:
            check(rc)
            self.ptr = _vp(p.value)
        return False

    def launch(self, stream):
        check(lib().exo_graph_launch(self.ptr, stream.ptr if isinstance(stream, Stream) else stream))

    def __del__(self):
        try:
            if self.ptr:
                lib().exo_graph_destroy(self.ptr)
                self.ptr = None
        except Exception:
            pass


class Comm(object):
    """exo_comm: NCCL communicator of this process' GPU (one process per GPU).  `unique_id`: the 128 bytes rank 0 obtained from
    Comm.unique_id(), shipped to every rank by the launcher (exotica_b200.distributed.init_from_env does it for torchrun-style
    environments)."""

    def __init__(self, rank, world, unique_id=None):
        if world > 1 and (unique_id is None or len(unique_id) != COMM_ID_BYTES):
            raise ExoticaError("a %d-rank communicator needs the %d-byte unique id of rank 0" % (world, COMM_ID_BYTES))
        h = _vp()
        check(lib().exo_comm_create(int(rank), int(world), bytes(unique_id) if unique_id is not None else None, ctypes.byref(h)))
        self.h = _vp(h.value)
        self.rank, self.world = int(rank), int(world)

    @staticmethod
    def unique_id():
        buf = ctypes.create_string_buffer(COMM_ID_BYTES)
        check(lib().exo_comm_get_unique_id(buf))
        return buf.raw

    @staticmethod
    def nccl_version():
        v = ctypes.c_int(0)
        check(lib().exo_comm_nccl_version(ctypes.byref(v)))
        return v.value

    def allgather(self, d_send, d_recv, count_per_rank, stream=None):
        check(lib().exo_comm_allgather(self.h, d_send, d_recv, int(count_per_rank), _sptr(stream)))

    def allreduce(self, d_send, d_recv, count, op, stream=None):
        check(lib().exo_comm_allreduce(self.h, d_send, d_recv, int(count), int(op), _sptr(stream)))

    def barrier(self, stream=None):
        check(lib().exo_comm_barrier(self.h, _sptr(stream)))

    def collective_count(self):
        return int(lib().exo_comm_collective_count(self.h))

    def peer_collective_count(self):
        """Collectives that went through the peer-memory windows instead of NCCL; -1: the windows are not available."""
        return int(lib().exo_comm_peer_collective_count(self.h))

    def cost_allgather(self, problem, d_x, n_traj_local, n_traj_per_rank, d_cost_all, layout=LAYOUT_SOA, stream=None):
        """problem: a _capi.Problem or raw exo_problem handle."""
        h = problem.h if hasattr(problem, "h") else problem
        check(lib().exo_problem_cost_allgather(h, self.h, d_x, int(n_traj_local), int(n_traj_per_rank), int(layout), d_cost_all, _sptr(stream)))

    def min_distance_allgather(self, d_min_eval, n_traj_local, T, n_traj_per_rank, d_min_traj_all=None, d_min_global=None, stream=None):
        check(lib().exo_min_distance_allgather(self.h, d_min_eval, int(n_traj_local), int(T), int(n_traj_per_rank), d_min_traj_all, d_min_global,
                                               _sptr(stream)))

    def destroy(self):
        if self.h:
            lib().exo_comm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


def _sptr(stream):
    return stream.ptr if isinstance(stream, Stream) else stream


class PinnedArray(object):
    """numpy view over page-locked host memory (for overlapped copies in update_batch_host)."""

    def __init__(self, shape):
        self.shape = tuple(int(s) for s in np.atleast_1d(shape))
        n = int(np.prod(self.shape))
        p = _vp()
        check(lib().exo_host_alloc_pinned(ctypes.byref(p), max(n, 1) * 8))
        self.ptr = p.value
        buf = (ctypes.c_double * max(n, 1)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=np.float64, count=n).reshape(self.shape)

    def free(self):
        if self.ptr:
            self.array = None
            lib().exo_host_free_pinned(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Tree(object):
    """exo_tree handle built from flat arrays (see exo_tree_desc)."""

    def __init__(self, flat, n_controlled, lower, upper, tree_state=None):
        self._keep = {k: np.ascontiguousarray(v) for k, v in flat.items()}
        f = self._keep
        self.lower = np.ascontiguousarray(lower, dtype=np.float64)
        self.upper = np.ascontiguousarray(upper, dtype=np.float64)
        d = TreeDesc()
        d.n_elements = len(f["id"])
        d.id = _i(f["id"])
        d.parent = _i(f["parent"])
        d.joint_type = _i(f["joint_type"])
        d.joint_origin = _d(f["origin"])
        d.joint_axis = _d(f["axis"])
        d.f_tip = _d(f["f_tip"])
        d.control_id = _i(f["control_id"])
        d.mimic_element = _i(f["mimic_elem"])
        d.mimic_multiplier = _d(f["mimic_mult"])
        d.mimic_offset = _d(f["mimic_off"])
        d.generated_slot = _i(f["generated"])
        d.n_controlled = int(n_controlled)
        d.lower = _d(self.lower)
        d.upper = _d(self.upper)
        h = _vp()
        check(lib().exo_tree_create(ctypes.byref(d), ctypes.byref(h)))
        self.h = h.value
        self.n = int(n_controlled)
        self.n_elements = d.n_elements
        if tree_state is not None:
            self.set_model_state(tree_state)

    def set_model_state(self, tree_state):
        s = np.ascontiguousarray(tree_state, dtype=np.float64)
        if s.size != self.n_elements:
            raise ExoticaError("tree state has %d entries, expected %d" % (s.size, self.n_elements))
        check(lib().exo_tree_set_model_state(self.h, _d(s)))

    def set_joint_limits(self, lower, upper):
        self.lower = np.ascontiguousarray(lower, dtype=np.float64)
        self.upper = np.ascontiguousarray(upper, dtype=np.float64)
        check(lib().exo_tree_set_joint_limits(self.h, _d(self.lower), _d(self.upper)))

    def __del__(self):
        try:
            if self.h:
                lib().exo_tree_destroy(self.h)
                self.h = None
        except Exception:
            pass


class Problem(object):
    """exo_problem handle: ordered task maps over one tree."""

    def __init__(self, tree):
        self.tree = tree
        h = _vp()
        check(lib().exo_problem_create(tree.h, ctypes.byref(h)))
        self.h = h.value
        self.finalized = False

    def add_task_map(self, kind, elem_a=None, off_a=None, elem_b=None, off_b=None, rotation_type=1, safe_percentage=0.0,
                     precision=1.0, radius=None, group=None, is_used=True, joint_map=None, joint_ref=None, axis=None, direction=None, mass=None,
                     enable_z=True):
        d = TaskMapDesc()
        d.kind = kind
        n = 0 if elem_a is None else len(elem_a)
        d.n_frames = n
        keep = []
        if n:
            ea = np.ascontiguousarray(elem_a, dtype=np.int32)
            eb = np.ascontiguousarray(elem_b if elem_b is not None else np.zeros(n), dtype=np.int32)
            ident = np.tile(np.array([1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0], dtype=np.float64), (n, 1))
            oa = np.ascontiguousarray(off_a if off_a is not None else ident, dtype=np.float64).reshape(n, 12)
            ob = np.ascontiguousarray(off_b if off_b is not None else ident, dtype=np.float64).reshape(n, 12)
            keep += [ea, eb, oa, ob]
            d.elem_a, d.elem_b, d.off_a, d.off_b = _i(ea), _i(eb), _d(oa), _d(ob)
        d.rotation_type = int(rotation_type)
        d.safe_percentage = float(safe_percentage)
        d.precision = float(precision)
        if radius is not None:
            r = np.ascontiguousarray(radius, dtype=np.float64)
            g = np.ascontiguousarray(group, dtype=np.int32)
            keep += [r, g]
            d.radius, d.group = _d(r), _i(g)
        d.is_used = 1 if is_used else 0
        if joint_map is not None:
            jm = np.ascontiguousarray(joint_map, dtype=np.int32)
            keep.append(jm)
            d.n_joint_map, d.joint_map = int(jm.size), _i(jm)
        if joint_ref is not None:
            jr = np.ascontiguousarray(joint_ref, dtype=np.float64)
            keep.append(jr)
            d.joint_ref = _d(jr)
        for name, val in (("axis", axis), ("direction", direction), ("mass", mass)):
            if val is not None:
                a = np.ascontiguousarray(val, dtype=np.float64)
                keep.append(a)
                setattr(d, name, _d(a))
        d.enable_z = 1 if enable_z else 0
        check(lib().exo_problem_add_task_map(self.h, ctypes.byref(d)))

    def finalize(self):
        check(lib().exo_problem_finalize(self.h))
        self.finalized = True
        self.length_phi = lib().exo_problem_length_phi(self.h)
        self.length_jacobian = lib().exo_problem_length_jacobian(self.h)
        self.n = self.tree.n

    def map_indexing(self, i):
        out = np.zeros(4, dtype=np.int32)
        check(lib().exo_problem_map_indexing(self.h, i, _i(out)))
        return tuple(int(v) for v in out)

    def program_info(self):
        out = np.zeros(8, dtype=np.int32)
        check(lib().exo_problem_program_info(self.h, _i(out)))
        keys = ["joint_ops", "emit_ops", "frame_slots", "joint_slots", "scratch_bytes_per_thread", "image_bytes", "accumulators", "ops"]
        return dict(zip(keys, (int(v) for v in out)))

    def set_generated(self, frames):
        frames = np.ascontiguousarray(frames, dtype=np.float64)
        T, n_gen = frames.shape[0], frames.shape[1]
        check(lib().exo_problem_set_generated(self.h, T, n_gen, _d(frames)))

    def update_batch_device(self, d_x, n_evals, d_phi, d_jac, T=1, layout=LAYOUT_SOA, stream=None):
        check(lib().exo_problem_update_batch(self.h, d_x, int(n_evals), int(T), int(layout), d_phi, d_jac, stream))

    def last_kernel_ms(self):
        ms = ctypes.c_float(0)
        check(lib().exo_problem_last_kernel_ms(self.h, ctypes.byref(ms)))
        return ms.value

    def launch_count(self):
        return int(lib().exo_problem_launch_count(self.h))

    def chain_info(self):
        out = (ctypes.c_int32 * 4)()
        check(lib().exo_problem_chain_info(self.h, out))
        return {"lowered": bool(out[0]), "joints": out[1], "all_revolute": bool(out[2] & 1), "orthogonal_links": bool(out[2] & 2), "frames": out[3]}

    def set_cost(self, rho, T=1, goal=None, ct=1.0, W=None):
        """rho[T][n_maps] (broadcast from [n_maps]); goal[1 or T][length_Phi] or None (SetZero goal); W[N] or None."""
        n_maps = lib().exo_problem_num_maps(self.h)
        rho = np.ascontiguousarray(np.broadcast_to(np.asarray(rho, dtype=np.float64), (T, n_maps)))
        d = CostDesc()
        d.T, d.T_goal, d.ct = int(T), 1, float(ct)
        d.rho = _d(rho)
        keep = [rho]
        if goal is not None:
            goal = np.ascontiguousarray(goal, dtype=np.float64).reshape(-1, self.length_phi)
            d.T_goal = goal.shape[0]
            d.goal = _d(goal)
            keep.append(goal)
        if W is not None:
            W = np.ascontiguousarray(W, dtype=np.float64)
            d.W = _d(W)
            keep.append(W)
        check(lib().exo_problem_set_cost(self.h, ctypes.byref(d)))
        self.cost_T = int(T)

    def _check_trajectories(self, X):
        T = getattr(self, "cost_T", None)
        if T is None:
            raise ExoticaError("exo_problem_set_cost has not been called")
        want = (T, self.n) if T > 1 else (self.n,)
        if X.ndim == 3 and T == 1 and X.shape[1:] == (1, self.n):
            return
        if X.ndim != len(want) + 1 or tuple(X.shape[1:]) != want:
            raise ExoticaError("Wrong state vector size! Got %s expected (B, %s)" % (X.shape, ", ".join(str(v) for v in want)))

    def cost_batch_device(self, d_x, d_phi, n_traj, d_cost_traj=None, d_cost_eval=None, layout=LAYOUT_SOA, stream=None):
        check(lib().exo_problem_cost_batch(self.h, d_x, d_phi, int(n_traj), int(layout), d_cost_traj, d_cost_eval, stream))

    def cost_batch_host(self, X):
        """X[B][T][N] (or [B][N] for end-pose problems) -> cost[B] through the host pipeline."""
        X = np.ascontiguousarray(X, dtype=np.float64)
        self._check_trajectories(X)
        B = X.shape[0]
        out = np.empty(B)
        check(lib().exo_problem_cost_batch_host(self.h, X.ctypes.data, B, LAYOUT_AOS_ROWMAJOR, out.ctypes.data))
        return out

    def cost_jacobian_batch_device(self, d_x, d_phi, d_jac, n_traj, d_grad, layout=LAYOUT_SOA, stream=None):
        check(lib().exo_problem_cost_jacobian_batch(self.h, d_x, d_phi, d_jac, int(n_traj), int(layout), d_grad, stream))

    def aico_task_messages_device(self, d_qhat, d_phi, d_jac, n_traj, d_R, d_r, d_rhat, d_cost, layout=LAYOUT_SOA, stream=None):
        check(lib().exo_problem_aico_task_messages_batch(self.h, d_qhat, d_phi, d_jac, int(n_traj), int(layout), d_R, d_r, d_rhat, d_cost, stream))

    def cost_gradient_batch_device(self, d_x, n_traj, d_grad, d_cost_traj=None, layout=LAYOUT_SOA, stream=None):
        check(lib().exo_problem_cost_gradient_batch(self.h, d_x, int(n_traj), int(layout), d_cost_traj, d_grad, stream))

    def update_cost_batch_device(self, d_x, n_traj, d_phi, d_jac, d_cost_traj, d_cost_eval=None, layout=LAYOUT_SOA, stream=None):
        """update_batch_device + cost_batch_device in one call (one generated kernel + the per-trajectory sum where it applies)."""
        check(lib().exo_problem_update_cost_batch(self.h, d_x, int(n_traj), int(layout), d_phi, d_jac, d_cost_traj, d_cost_eval, stream))

    def cost_jacobian_batch_host(self, X, with_cost=True):
        """X[B][T][N] (or [B][N] for end-pose problems) -> (cost[B] or None, grad like X) through the host pipeline; Phi and J
        stay on the device."""
        X = np.ascontiguousarray(X, dtype=np.float64)
        self._check_trajectories(X)
        B = X.shape[0]
        cost = np.empty(B) if with_cost else None
        grad = np.empty_like(X)
        check(lib().exo_problem_cost_jacobian_batch_host(self.h, X.ctypes.data, B, LAYOUT_AOS_ROWMAJOR, cost.ctypes.data if with_cost else None,
                                                         grad.ctypes.data))
        return cost, grad

    def hessian_batch_device(self, d_x, d_jac, n_evals, d_hess, T=1, layout=LAYOUT_SOA, stream=None):
        check(lib().exo_problem_hessian_batch(self.h, d_x, d_jac, int(n_evals), int(T), int(layout), d_hess, stream))

    def ik_solve_batch(self, Q0, max_iterations=100, tolerance=1e-5, regularization_rate=1e-9, maximum_regularization=1e9,
                       gradient_tolerance=1e-9, step_tolerance=1e-5, max_step=0.0, threshold_increase=0.1, threshold_decrease=0.5,
                       minimum_regularization=1e-12):
        """Multi-start IK: Q0[E][N] start states -> (Q[E][N], cost[E], iterations[E], termination[E]).  Needs set_cost(T=1)."""
        Q0 = np.ascontiguousarray(Q0, dtype=np.float64)
        E, N = Q0.shape
        dq, dc = DeviceBuffer(E * N), DeviceBuffer(E)
        di, dt = DeviceBuffer((E + 1) // 2 + 1), DeviceBuffer((E + 1) // 2 + 1)  # int32 arrays in double-sized storage
        dq.upload(Q0.T)
        prm = IkParams(int(max_iterations), float(tolerance), float(regularization_rate), float(maximum_regularization),
                       float(gradient_tolerance), float(step_tolerance), float(max_step), float(threshold_increase), float(threshold_decrease),
                       float(minimum_regularization))
        check(lib().exo_problem_ik_solve_batch(self.h, dq.ptr, E, ctypes.byref(prm), dc.ptr, di.ptr, dt.ptr, None))
        check(lib().exo_device_synchronize())
        Q = np.ascontiguousarray(dq.download((N, E)).T)
        cost = dc.download((E,))
        it = np.zeros(E, dtype=np.int32)
        tm = np.zeros(E, dtype=np.int32)
        check(lib().exo_memcpy_d2h(it.ctypes.data, di.ptr, E * 4, None))
        check(lib().exo_memcpy_d2h(tm.ctypes.data, dt.ptr, E * 4, None))
        check(lib().exo_stream_synchronize(None))
        for b in (dq, dc, di, dt):
            b.free()
        return Q, cost, it, tm

    def set_min_distance_output(self, d_ptr):
        check(lib().exo_problem_set_min_distance_output(self.h, d_ptr))

    def set_kernel(self, choice):
        check(lib().exo_problem_set_kernel(self.h, int(choice)))

    def generated_kernel_source(self, compile=False, cost=False, update_cost=False):
        """CUDA source of the per-robot generated tree kernel (cost=True: of the fused cost + gradient kernel, update_cost=True: of the
        update kernel with the cost epilogue -- both after set_cost); compile=True also runs NVRTC (works without a GPU)."""
        need = ctypes.c_int64(0)
        kind = 2 if cost else (4 if update_cost else 0)
        check(lib().exo_problem_generated_kernel_source(self.h, None, 0, ctypes.byref(need), kind))
        buf = ctypes.create_string_buffer(need.value)
        check(lib().exo_problem_generated_kernel_source(self.h, buf, need.value, None, kind | (1 if compile else 0)))
        return buf.value.decode()

    def last_kernel(self):
        return lib().exo_problem_last_kernel(self.h).decode()

    def update_batch_host(self, X, T=1, layout=LAYOUT_AOS_ROWMAJOR, want_jacobian=True, out_phi=None, out_jac=None):
        """Host arrays in, host arrays out (chunked copy/compute pipeline inside the library).
        AoS layouts: X[E][N] -> Phi[E][LP], J[E][LJ][N] (row-major) or J[E][N][LJ] (col-major).
        SoA: X[N][E] -> Phi[LP][E], J[LJ][N][E]."""
        X = np.ascontiguousarray(X, dtype=np.float64)
        N, LP, LJ = self.n, self.length_phi, self.length_jacobian
        if layout == LAYOUT_SOA:
            if X.ndim != 2 or X.shape[0] != N:
                raise ExoticaError("Wrong state vector size! Got %s expected (%d, E)" % (X.shape, N))
            E = X.shape[1]
            pshape, jshape = (LP, E), (LJ, N, E)
        else:
            if X.ndim != 2 or X.shape[1] != N:
                raise ExoticaError("Wrong state vector size! Got %s expected (E, %d)" % (X.shape, N))
            E = X.shape[0]
            pshape = (E, LP)
            jshape = (E, LJ, N) if layout == LAYOUT_AOS_ROWMAJOR else (E, N, LJ)
        phi = checked_output(out_phi, pshape, "out_phi") if out_phi is not None else np.empty(pshape, dtype=np.float64)
        jac = None
        if want_jacobian:
            jac = checked_output(out_jac, jshape, "out_jac") if out_jac is not None else np.empty(jshape, dtype=np.float64)
        check(lib().exo_problem_update_batch_host(self.h, X.ctypes.data, E, int(T), int(layout), phi.ctypes.data,
                                                  jac.ctypes.data if jac is not None else None))
        return phi, jac

    def __del__(self):
        try:
            if self.h:
                lib().exo_problem_destroy(self.h)
                self.h = None
        except Exception:
            pass

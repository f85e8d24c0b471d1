"""ctypes declarations for include/exotica_b200_host.h (the C wrappers over the C++ host mirror)."""
import ctypes

from . import _capi

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int32)
_vp = ctypes.c_void_p
_cp = ctypes.c_char_p

HOST_SYMBOLS = [
    ("exo_host_set_package_path", ctypes.c_int, [_cp, _cp]),
    ("exo_host_registered_maps", _cp, []),
    ("exo_host_registered_problems", _cp, []),
    ("exo_host_get_frame", ctypes.c_int, [_dp, ctypes.c_int, _dp]),
    ("exo_host_scene_create", ctypes.c_int, [_cp, _cp, _cp, ctypes.POINTER(_vp)]),
    ("exo_host_scene_destroy", ctypes.c_int, [_vp]),
    ("exo_host_scene_names", _cp, [_vp, ctypes.c_int]),
    ("exo_host_scene_num_controlled", ctypes.c_int, [_vp]),
    ("exo_host_scene_num_elements", ctypes.c_int, [_vp]),
    ("exo_host_scene_base_types", ctypes.c_int, [_vp, _ip]),
    ("exo_host_scene_tree", ctypes.c_int, [_vp, _ip, _ip, _ip, _ip, _dp, _dp, _dp]),
    ("exo_host_scene_element_index", ctypes.c_int, [_vp, _cp]),
    ("exo_host_scene_joint_limits", ctypes.c_int, [_vp, _dp, _dp]),
    ("exo_host_scene_set_joint_limits", ctypes.c_int, [_vp, _dp, _dp]),
    ("exo_host_scene_set_seed", ctypes.c_int, [_vp, ctypes.c_uint]),
    ("exo_host_scene_random_state", ctypes.c_int, [_vp, _dp]),
    ("exo_host_scene_get_model_state", ctypes.c_int, [_vp, _dp]),
    ("exo_host_scene_set_model_state", ctypes.c_int, [_vp, _dp, ctypes.c_int]),
    ("exo_host_scene_num_model_joints", ctypes.c_int, [_vp]),
    ("exo_host_scene_add_object", ctypes.c_int, [_vp, _cp, _dp, _cp]),
    ("exo_host_scene_load_scene", ctypes.c_int, [_vp, _cp, _dp]),
    ("exo_host_scene_load_scene_file", ctypes.c_int, [_vp, _cp, _dp]),
    ("exo_host_scene_clean", ctypes.c_int, [_vp]),
    ("exo_host_scene_add_trajectory", ctypes.c_int, [_vp, _cp, ctypes.c_int, _dp]),
    ("exo_host_scene_add_trajectory_string", ctypes.c_int, [_vp, _cp, _cp]),
    ("exo_host_scene_get_trajectory", ctypes.c_int, [_vp, _cp, ctypes.c_char_p, ctypes.c_int]),
    ("exo_host_trajectory_sample", ctypes.c_int, [_cp, ctypes.c_int, _dp, _dp, _dp]),
    ("exo_host_scene_request_frames", ctypes.c_int, [_vp, ctypes.c_int, _cp, _dp, _cp, _dp, ctypes.c_int]),
    ("exo_host_scene_update", ctypes.c_int, [_vp, _dp, ctypes.c_int, ctypes.c_double, _dp, _dp]),
    ("exo_host_scene_update_batch", ctypes.c_int, [_vp, _vp, ctypes.c_int64, ctypes.c_int, _vp, _vp]),
    ("exo_host_scene_fk", ctypes.c_int, [_vp, _cp, _dp, _cp, _dp, _dp]),
    ("exo_host_scene_jacobian", ctypes.c_int, [_vp, _cp, _dp, _cp, _dp, _dp]),
    ("exo_host_load_problem_xml", ctypes.c_int, [_cp, ctypes.POINTER(_vp)]),
    ("exo_host_problem_destroy", ctypes.c_int, [_vp]),
    ("exo_host_problem_dims", ctypes.c_int, [_vp, _ip]),
    ("exo_host_problem_type", _cp, [_vp]),
    ("exo_host_problem_num_tasks", ctypes.c_int, [_vp]),
    ("exo_host_problem_task_name", _cp, [_vp, ctypes.c_int]),
    ("exo_host_problem_task_indexing", ctypes.c_int, [_vp, ctypes.c_int, _ip]),
    ("exo_host_problem_scene", _vp, [_vp]),
    ("exo_host_problem_start_state", ctypes.c_int, [_vp, _dp]),
    ("exo_host_problem_update", ctypes.c_int, [_vp, _dp, ctypes.c_int, ctypes.c_int, _dp, _dp]),
    ("exo_host_problem_update_trajectory", ctypes.c_int, [_vp, _dp, ctypes.c_int, _dp, _dp]),
    ("exo_host_problem_update_batch", ctypes.c_int, [_vp, _vp, ctypes.c_int64, ctypes.c_int, _vp, _vp]),
    ("exo_host_problem_update_batch_device", ctypes.c_int, [_vp, _vp, ctypes.c_int64, ctypes.c_int, _vp, _vp, _vp]),
    ("exo_host_problem_cost_batch", ctypes.c_int, [_vp, _vp, ctypes.c_int64, _vp]),
    ("exo_host_problem_cost_jacobian_batch", ctypes.c_int, [_vp, _vp, ctypes.c_int64, _vp, _vp]),
    ("exo_host_problem_cost_batch_device", ctypes.c_int, [_vp, _vp, _vp, ctypes.c_int64, ctypes.c_int, _vp, _vp, _vp]),
    ("exo_host_problem_update_cost_batch_device", ctypes.c_int, [_vp, _vp, ctypes.c_int64, ctypes.c_int, _vp, _vp, _vp, _vp, _vp]),
    ("exo_host_problem_ik_solve_batch", ctypes.c_int, [_vp, _vp, ctypes.c_int64, ctypes.POINTER(_capi.IkParams), _vp, _vp, _vp, _vp]),
    ("exo_host_problem_hessian_batch", ctypes.c_int, [_vp, _vp, ctypes.c_int64, _vp]),
    ("exo_host_problem_last_kernel_ms", ctypes.c_int, [_vp, ctypes.POINTER(ctypes.c_float)]),
    ("exo_host_problem_launch_count", ctypes.c_int64, [_vp]),
    ("exo_host_problem_update_count", ctypes.c_int64, [_vp]),
    ("exo_host_problem_backend", _vp, [_vp]),
]

_READY = False


def lib():
    global _READY
    L = _capi.lib()
    if not _READY:
        for name, res, args in HOST_SYMBOLS:
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _READY = True
    return L

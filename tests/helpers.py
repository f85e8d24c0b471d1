"""Shared test helpers: build the same problem on the CPU oracle and on the CUDA backend (through the C-ABI)."""
import numpy as np

from exotica_b200 import _capi


def gpu_problem_from_maps(orc, maps, tree_state=None):
    """orc: oracle.Oracle (its RobotTree provides the flat tree and name->index resolution);
    maps: the same list of dicts given to orc.set_problem()."""
    tree = _capi.Tree(orc.tree.flat(), orc.tree.n_controlled, orc.lower, orc.upper,
                      tree_state if tree_state is not None else orc.tree.tree_state)
    prob = _capi.Problem(tree)
    for m in maps:
        kind = {"EffFrame": _capi.MAP_EFF_FRAME, "EffPosition": _capi.MAP_EFF_POSITION, "JointLimit": _capi.MAP_JOINT_LIMIT,
                "SphereCollision": _capi.MAP_SPHERE_COLLISION, "Frames": _capi.MAP_FRAMES, "EffOrientation": _capi.MAP_EFF_ORIENTATION,
                "EffPositionXY": _capi.MAP_EFF_POSITION_XY, "Distance": _capi.MAP_DISTANCE, "JointPose": _capi.MAP_JOINT_POSE,
                "EffAxisAlignment": _capi.MAP_EFF_AXIS_ALIGNMENT, "CenterOfMass": _capi.MAP_CENTER_OF_MASS}[m["type"]]
        frames = m.get("EndEffector", [])
        kw = dict(rotation_type=_capi.ROTATION_TYPES[m.get("Type", "RPY")], safe_percentage=m.get("SafePercentage", 0.0),
                  precision=m.get("Precision", 1.0), is_used=m.get("used", True))
        if frames:
            ea, oa, eb, ob = orc.resolve_frames(frames)
            kw.update(elem_a=ea, off_a=oa, elem_b=eb, off_b=ob)
        if kind == _capi.MAP_JOINT_POSE:
            kw["joint_map"] = m.get("JointMap")
            kw["joint_ref"] = m.get("JointRef")
        if kind == _capi.MAP_EFF_AXIS_ALIGNMENT:
            kw["axis"] = [f.get("Axis", [1.0, 0.0, 0.0]) for f in frames]
            kw["direction"] = [f.get("Direction", [1.0, 0.0, 0.0]) for f in frames]
        if kind == _capi.MAP_CENTER_OF_MASS:
            kw["mass"] = [float(f["Mass"]) for f in frames]
            kw["enable_z"] = m.get("EnableZ", True)
        if kind == _capi.MAP_SPHERE_COLLISION:
            names = sorted(set(f.get("Group", "default") for f in frames))
            kw["group"] = [names.index(f.get("Group", "default")) for f in frames]
            kw["radius"] = [float(f["Radius"]) for f in frames]
        prob.add_task_map(kind, **kw)
    prob.finalize()
    return tree, prob


def record(name, value):
    """Measured maximum errors of the parity tests, kept next to the other GPU-run artefacts (gpurun_out/parity_measured.jsonl;
    copied into profiles/ per round) so that tolerances are justified by numbers, not by guesses."""
    import json
    import os
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_measured.jsonl"), "a") as f:
            f.write(json.dumps({"test": name, "max_err": float(value)}) + "\n")
    except OSError:
        pass
    return float(value)


GUARD = 64  # doubles behind every output buffer that the kernels must leave untouched


def _guard_intact(buf, n):
    tail = buf.download((n + GUARD,))[n:]
    return bool((tail.view(np.uint64) == np.uint64(0xFFFFFFFFFFFFFFFF)).all())


def run_gpu(prob, X, layout=_capi.LAYOUT_SOA, T=1, want_jacobian=True):
    """X[E][N] (always evaluation-major here) -> Phi[E][LP], J[E][LJ][N] via device buffers in `layout`."""
    X = np.ascontiguousarray(X, dtype=np.float64)
    E, N = X.shape
    LP, LJ = prob.length_phi, prob.length_jacobian
    dx = _capi.DeviceBuffer(max(E * N, 1))
    dphi = _capi.DeviceBuffer(E * LP + GUARD)
    djac = _capi.DeviceBuffer(E * LJ * N + GUARD) if want_jacobian else None
    # poison the outputs: the kernel must write every entry itself -- and nothing behind the end of the buffers
    dphi.fill_bytes(0xFF)
    if djac is not None:
        djac.fill_bytes(0xFF)
    dx.upload(X.T if layout == _capi.LAYOUT_SOA else X)
    prob.update_batch_device(dx.ptr, E, dphi.ptr, djac.ptr if djac is not None else None, T=T, layout=layout)
    _capi.check(_capi.lib().exo_device_synchronize())
    assert _guard_intact(dphi, E * LP), "write behind the end of Phi"
    assert djac is None or _guard_intact(djac, E * LJ * N), "write behind the end of J"
    if layout == _capi.LAYOUT_SOA:
        phi = dphi.download((LP, E)).T
        jac = djac.download((LJ, N, E)).transpose(2, 0, 1) if want_jacobian else None
    elif layout == _capi.LAYOUT_AOS_ROWMAJOR:
        phi = dphi.download((E, LP))
        jac = djac.download((E, LJ, N)) if want_jacobian else None
    else:
        phi = dphi.download((E, LP))
        jac = djac.download((E, N, LJ)).transpose(0, 2, 1) if want_jacobian else None
    return np.ascontiguousarray(phi), (np.ascontiguousarray(jac) if jac is not None else None)


def angle_entries(maps, orc):
    """Indices of Phi that hold Euler angles / angle-axis components (branch cuts: compare modulo 2 pi)."""
    idx = []
    start = 0
    for m in maps:
        t = m["type"]
        nf = len(m.get("EndEffector", []))
        if t == "EffFrame":
            rot = m.get("Type", "RPY")
            rl = _capi.ROTATION_LENGTH[_capi.ROTATION_TYPES[rot]]
            if rot in ("RPY", "ZYX", "ZYZ"):
                for f in range(nf):
                    idx += [start + f * (3 + rl) + 3 + k for k in range(3)]
            start += nf * (3 + rl)
        elif t == "EffOrientation":
            rot = m.get("Type", "RPY")
            rl = _capi.ROTATION_LENGTH[_capi.ROTATION_TYPES[rot]]
            if rot in ("RPY", "ZYX", "ZYZ"):
                for f in range(nf):
                    idx += [start + f * rl + k for k in range(3)]
            start += nf * rl
        elif t == "EffPositionXY":
            start += 2 * nf
        elif t in ("Distance", "EffAxisAlignment"):
            start += nf
        elif t == "CenterOfMass":
            start += 3 if m.get("EnableZ", True) else 2
        elif t == "JointPose":
            start += len(m.get("JointMap", range(orc.n)))
        elif t == "EffPosition":
            start += 3 * nf
        elif t == "JointLimit":
            start += orc.n
        elif t == "SphereCollision":
            g = len(set(f.get("Group", "default") for f in m["EndEffector"]))
            start += g * (g - 1) // 2
        elif t == "Frames":
            start += 12 * nf
    return np.array(idx, dtype=np.int64)


def max_abs_err(a, b, wrap_idx=None):
    d = np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64)
    if wrap_idx is not None and len(wrap_idx):
        w = d[..., wrap_idx]
        d = d.copy()
        d[..., wrap_idx] = (w + np.pi) % (2 * np.pi) - np.pi
    if d.size == 0:
        return 0.0
    if np.isnan(d).any():
        return float("nan")
    return float(np.max(np.abs(d)))


def random_states(orc, E, seed, spread=0.0):
    """Uniform in the joint limits (KinematicTree::GetRandomControlledState, kinematic_tree.cpp:1025-1033);
    unbounded joints (|limit| = 1e20) draw from [-1, 1]; `spread` widens the box to exercise limit violations."""
    rng = np.random.default_rng(seed)
    lo = np.where(orc.lower < -1e19, -1.0, orc.lower) - spread
    hi = np.where(orc.upper > 1e19, 1.0, orc.upper) + spread
    return rng.uniform(lo, hi, size=(E, orc.n))

"""GPU tests of the C-ABI's collective / event / graph entry points (include/exotica_b200.h, SURVEY.md section 8 row e).

One GPU is enough for everything but the wire: a world-size-1 communicator takes the same code path up to the NCCL call
(slot layout, padding, per-trajectory reductions, graph capture).  bench.py's multi_gpu section runs the same calls over real
NCCL ranks at 2 / 4 / 8 GPUs and checks what was gathered against the oracle.
"""
import numpy as np
import pytest

from conftest import robot_paths
from exotica_b200 import _capi
from helpers import gpu_problem_from_maps, random_states

pytestmark = pytest.mark.gpu


def _lwr(oracle_lib):
    return oracle_lib.Oracle(oracle_lib.load_robot(*robot_paths("lwr_simplified"), joint_group="arm"))


def _trajectories(orc, B, T, seed):
    rng = np.random.default_rng(seed)
    q0 = random_states(orc, B, seed=seed + 1)
    return np.clip(q0[:, None, :] + np.cumsum(rng.normal(0, 0.05, size=(B, T, orc.n)), axis=1), orc.lower - 0.1, orc.upper + 0.1)


def test_events_and_streams():
    s = _capi.Stream()
    a, b = _capi.Event(), _capi.Event()
    buf = _capi.DeviceBuffer(1 << 20)
    a.record(s)
    _capi.check(_capi.lib().exo_memset_device(buf.ptr, 0, 8 << 20, s.ptr))
    b.record(s)
    b.synchronize()
    assert a.elapsed_ms(b) >= 0.0


def test_world_size_1_collectives_are_copies():
    comm = _capi.Comm(0, 1)
    src, dst = _capi.DeviceBuffer(16), _capi.DeviceBuffer(16)
    v = np.arange(16, dtype=np.float64)
    src.upload(v)
    comm.allgather(src.ptr, dst.ptr, 16)
    assert np.array_equal(dst.download((16,)), v)
    dst.fill_bytes(0)
    comm.allreduce(src.ptr, dst.ptr, 16, _capi.REDUCE_MIN)
    _capi.check(_capi.lib().exo_device_synchronize())
    assert np.array_equal(dst.download((16,)), v)
    comm.barrier()
    assert comm.collective_count() == 0  # no NCCL call was needed
    with pytest.raises(_capi.ExoticaError):
        _capi.Comm(0, 2, None)
    with pytest.raises(_capi.ExoticaError):
        comm.allreduce(src.ptr, dst.ptr, 16, 7)


@pytest.mark.parametrize("captured", [False, True])
def test_cost_allgather_matches_the_oracle(oracle_lib, captured):
    """exo_problem_cost_allgather: this rank's GetCost() values land in its slot of the gathered array (zero padded); replaying
    the call from a CUDA graph gives the same bits."""
    orc = _lwr(oracle_lib)
    maps = [{"type": "EffFrame", "EndEffector": [{"Link": "lwr_arm_7_link"}]}, {"type": "JointLimit"}]
    orc.set_problem(maps)
    tree, prob = gpu_problem_from_maps(orc, maps)
    B, T, per = 37, 25, 40
    X = _trajectories(orc, B, T, 31)
    W = np.linspace(1.0, 2.0, 7)
    ct = 1.0 / 0.05 / T
    prob.set_cost([3.0, 5.0], T=T, ct=ct, W=W)
    comm = _capi.Comm(0, 1)
    stream = _capi.Stream()
    dx, dall = _capi.DeviceBuffer(B * T * 7), _capi.DeviceBuffer(per)
    dx.upload(X.reshape(B * T, 7).T)
    dall.fill_bytes(0xFF)
    comm.cost_allgather(prob, dx.ptr, B, per, dall.ptr, stream=stream)
    stream.synchronize()
    if captured:
        first = dall.download((per,))
        dall.fill_bytes(0xFF)
        with _capi.Graph.capture(stream) as g:
            comm.cost_allgather(prob, dx.ptr, B, per, dall.ptr, stream=stream)
        g.launch(stream)
        g.launch(stream)
        stream.synchronize()
        assert np.array_equal(dall.download((per,)), first)
    got = dall.download((per,))
    ref = np.array([orc.trajectory_cost(X[b], [3.0, 5.0], ct=ct, W=W)[0] for b in range(B)])
    assert np.max(np.abs(got[:B] - ref) / np.maximum(1.0, np.abs(ref))) <= 1e-10
    assert not got[B:].any()
    assert prob.last_kernel() == "exo_chain_cost_kernel"


def test_min_distance_allgather(oracle_lib):
    comm = _capi.Comm(0, 1)
    B, T, per = 11, 7, 16
    rng = np.random.default_rng(4)
    d = rng.uniform(-0.2, 1.0, size=(B, T))
    dmin, dall, dg = _capi.DeviceBuffer(B * T), _capi.DeviceBuffer(per), _capi.DeviceBuffer(1)
    dmin.upload(d)
    comm.min_distance_allgather(dmin.ptr, B, T, per, dall.ptr, dg.ptr)
    _capi.check(_capi.lib().exo_device_synchronize())
    got = dall.download((per,))
    assert np.array_equal(got[:B], d.min(axis=1)) and np.isinf(got[B:]).all() and (got[B:] > 0).all()
    assert dg.download((1,))[0] == d.min()
    dg.fill_bytes(0)
    comm.min_distance_allgather(dmin.ptr, B, T, per, None, dg.ptr)  # global minimum only
    _capi.check(_capi.lib().exo_device_synchronize())
    assert dg.download((1,))[0] == d.min()
    comm.min_distance_allgather(None, 0, T, per, dall.ptr, dg.ptr)  # a rank without trajectories contributes +inf
    _capi.check(_capi.lib().exo_device_synchronize())
    assert np.isinf(dall.download((per,))).all() and np.isinf(dg.download((1,))[0])

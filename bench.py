#!/usr/bin/env python3
"""bench.py -- batched FK + Jacobian evaluations/sec (LWR 7-DoF) on N B200s, with parity, roofline and CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--configs E]

Workload (BASELINE.json configs[1], "cfg2a"): LWR 7-DoF, EffFrame(RPY) on lwr_arm_7_link, E = 2^20 seeded random
configurations per GPU per step, FP64.  One step = one pass of the hot path over that batch = one kernel launch.
Weak scaling: every rank evaluates its own 2^20 configurations, no data-path collective.
No torch anywhere: device memory, streams, events and the NCCL communicator all come from the C-ABI
(include/exotica_b200.h); under torchrun the ranks find each other through exotica_b200.distributed.init_from_env.

  value     whole-job evaluations/s with X already resident in HBM in the kernel's native SoA layout, timed with CUDA
            events on the launching stream over exactly K steps, max over ranks.  The steps rotate through 4 distinct X
            buffers (4 x 58.7 MB > the 126 MB L2), so every launch reads its inputs from HBM.
  parity    a seeded sample of the Phi / J buffers the LAST TIMED step wrote, compared with the CPU oracle (checker only,
            outside the timed region); the run fails above 1e-10.
  e2e       the same metric through the public API with HOST buffers (pinned, evaluation-major as the reference exposes
            them): H2D of X and D2H of Phi and J inside the timed region, every step.
  roofline  achieved = 440 algorithmic bytes/eval (read q 56 + write Phi 48 + write J 336)
            * evals per launch / average launch duration, against the measured HBM copy bandwidth in MEASURED_PEAKS.json;
            traffic = DRAM bytes per launch from the committed ncu capture of back-to-back launches (profiles/traffic.json).
  cpu_baseline  the CPU oracle (restated reference algorithm, one configuration at a time) on this box's host cores.
  cfg1      BASELINE.json configs[0] (example_ik.xml's exact shape: EffFrame on lwr_arm_6_link + quaternion LinkOffset, one
            configuration): CPU oracle microseconds per evaluation beside the latency of one problem.update(x) on the GPU.
  multi_gpu (N > 1) the configurations with a real exchange step, strong scaling of 4096 trajectories x T = 100: cfg3
            (EffFrame + JointLimit; update + cost + all-gather of the 4096 costs) and cfg5 (SphereCollision; update + min
            distance all-gather), each one CUDA graph per step on the stream that carries kernels and collective.

--impl reference times that CPU oracle alone with all host threads (the reference itself cannot be built in this
image: it needs catkin/ROS/MoveIt/orocos_kdl/Eigen) and prints the same JSON line with "impl": "reference".
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ROBOTS = os.path.join(ROOT, "tests", "fixtures", "robots")
METRIC = "batched FK+Jacobian evals/sec (LWR 7-DoF)"
UNIT = "evals/s"
BYTES_PER_EVAL = 56 + 48 + 336  # read q (7 f64) + write Phi (6 f64) + write J (6x7 f64)
HBM_FALLBACK_GBS = 6650.0       # B200_PROFILING.md fallback if MEASURED_PEAKS.json is absent
N_X_BUFFERS = 4                 # distinct input buffers rotated through the timed loop
PARITY_SAMPLE = 8192
PARITY_TOL = 1e-10

_emit = print  # replaced in main(): the JSON line goes to the original stdout


def workload_name(E):
    return "cfg2a: LWR 7-DoF EffFrame(RPY) lwr_arm_7_link, %d random configurations per GPU per step" % E


def lwr_scene():
    return ("exotica/Scene", {"Name": "Scene", "JointGroup": "arm", "URDF": os.path.join(ROBOTS, "lwr_simplified.urdf"),
                              "SRDF": os.path.join(ROBOTS, "lwr_simplified.srdf")})


def lwr_initializer():
    eff = ("exotica/EffFrame", {"Name": "Position", "Type": "RPY", "EndEffector": [("exotica/Frame", {"Link": "lwr_arm_7_link"})]})
    return ("exotica/UnconstrainedEndPoseProblem", {"Name": "bench", "PlanningScene": lwr_scene(), "Maps": [eff],
                                                    "Cost": [("exotica/Task", {"Task": "Position"})]})


CFG1_OFFSET = [0, 0, 0, 0.7071067811865476, -4.3297802811774664e-17, 0.7071067811865475, 4.3297802811774664e-17]


def cfg1_initializer():
    """exotica_examples/resources/configs/example_ik.xml:9-33 (BASELINE.json configs[0])."""
    eff = ("exotica/EffFrame", {"Name": "Position", "EndEffector": [("exotica/Frame", {"Link": "lwr_arm_6_link", "LinkOffset": CFG1_OFFSET})]})
    return ("exotica/UnconstrainedEndPoseProblem", {"Name": "MyProblem", "PlanningScene": lwr_scene(), "Maps": [eff],
                                                    "Cost": [("exotica/Task", {"Task": "Position"})], "StartState": [0] * 7,
                                                    "W": [7, 6, 5, 4, 3, 2, 1]})


def random_configurations(lower, upper, E, seed):
    import numpy as np
    rng = np.random.default_rng(seed)
    return rng.uniform(lower, upper, size=(E, len(lower)))


def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            with open(path) as f:
                return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst: the kernel is timed alone)"
        except Exception:
            pass
    return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


def recorded_traffic():
    """DRAM bytes per launch of the dominant kernel (dram__bytes_read.sum + dram__bytes_write.sum) from the committed ncu capture
    of back-to-back launches of this same command, and where it came from."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        try:
            with open(path) as f:
                d = json.load(f)
            return d.get("dram_bytes_per_launch"), d
        except Exception:
            return None, None
    return None, None


class ClockSampler(object):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, device_index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nv = None

    def _run(self):
        nv = self._nv
        names = {"hw_slowdown": "nvmlClocksThrottleReasonHwSlowdown", "hw_thermal_slowdown": "nvmlClocksThrottleReasonHwThermalSlowdown",
                 "sw_thermal_slowdown": "nvmlClocksThrottleReasonSwThermalSlowdown", "sw_power_cap": "nvmlClocksThrottleReasonSwPowerCap"}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for key, attr in names.items():
                    if mask & getattr(nv, attr, 0):
                        self.reasons.add(key)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self._nv is not None:
            self._stop.clear()
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self):
        if self._thread is not None:
            self._stop.set()
            self._thread.join()
            self._thread = None

    def summary(self, note):
        out = {"sm_mhz": (statistics.median(self.samples) if self.samples else None), "sm_max_mhz": self.max_mhz,
               "reasons": sorted(self.reasons), "samples": len(self.samples), "how": note}
        return out


def lwr_oracle(maps, add_collision_elements=False):
    from oracle import oracle
    tree = oracle.load_robot(os.path.join(ROBOTS, "lwr_simplified.urdf"), os.path.join(ROBOTS, "lwr_simplified.srdf"), "arm",
                             add_collision_elements=add_collision_elements)
    orc = oracle.Oracle(tree)
    orc.set_problem(maps)
    return orc


CFG2A_MAPS = [{"type": "EffFrame", "Type": "RPY", "EndEffector": [{"Link": "lwr_arm_7_link"}]}]


def time_cpu_oracle(E_request, threads, target_seconds=12.0, seed=1234 + 2):
    """Times the CPU oracle (restated reference algorithm incl. the per-collision-shape tree elements the reference
    updates, scene.cpp:795-821) on a bounded sample of the workload.  Returns (evals/s, sample description, E used)."""
    orc = lwr_oracle(CFG2A_MAPS, add_collision_elements=True)
    calib = random_configurations(orc.lower, orc.upper, 20000 * threads, seed)
    t0 = time.perf_counter()
    orc.problem_update_batch(calib, n_threads=threads, discard=True)
    rate = calib.shape[0] / max(time.perf_counter() - t0, 1e-9)
    # bounded sample: as many passes over (a prefix of) the step's batch as fit the time target
    E = int(min(E_request, max(50000, rate * target_seconds)))
    passes = max(1, int(round(rate * target_seconds / E)))
    X = random_configurations(orc.lower, orc.upper, E, seed)
    t0 = time.perf_counter()
    for _ in range(passes):
        orc.problem_update_batch(X, n_threads=threads, discard=True)
    dt = time.perf_counter() - t0
    sample = ("%d pass(es) over %d of the step's configurations, %d thread(s), Phi and J computed per configuration and not stored, %.2f s"
              % (passes, E, threads, dt))
    return passes * E / dt, sample, E


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle
    flags = oracle.use_native_build()
    threads = host_threads()
    E = args.configs
    for _ in range(max(args.warmup, 0)):
        time_cpu_oracle(E, threads, target_seconds=0.5)
    values, samples, used = [], [], []
    t_begin = time.perf_counter()
    for _ in range(args.steps):
        v, s, e = time_cpu_oracle(E, threads, target_seconds=min(10.0, 120.0 / max(args.steps, 1)))
        values.append(v)
        samples.append(s)
        used.append(e)
    wall = time.perf_counter() - t_begin
    value = statistics.median(values)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(E), "note": "CPU oracle = restated reference algorithm (the reference cannot be built here); "
                       "each step evaluates a bounded sample of the step's batch on all host threads; value = median evals/s"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": samples[-1], "build": flags},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    _emit(json.dumps(line))


# ---- device helpers (C-ABI only) ---------------------------------------------------------------------------------------------
class DeviceScalar(object):
    """One double on the device, for max / min over ranks of host-side numbers."""

    def __init__(self, capi):
        self.capi = capi
        self.buf = capi.DeviceBuffer(1)

    def reduce(self, comm, value, op):
        import numpy as np
        # collectives of one communicator are issued in one order on every rank: drain the streams around this one
        self.capi.check(self.capi.lib().exo_device_synchronize())
        self.buf.upload(np.array([value], dtype=np.float64))
        comm.allreduce(self.buf.ptr, self.buf.ptr, 1, op, None)
        self.capi.check(self.capi.lib().exo_device_synchronize())
        return float(self.buf.download((1,))[0])


def wrap_angles(d, idx):
    import numpy as np
    d = d.copy()
    d[..., idx] = (d[..., idx] + np.pi) % (2 * np.pi) - np.pi
    return d


def check_parity(capi, X_last, d_phi, d_jac, E, N, LP, LJ, seed):
    """Oracle parity of the buffers the last timed step wrote (SoA on the device): a seeded sample of evaluations."""
    import numpy as np
    n = min(PARITY_SAMPLE, E)
    sel = np.sort(np.random.default_rng(seed).choice(E, size=n, replace=False))
    phi = d_phi.download((LP, E))[:, sel].T
    jac = d_jac.download((LJ, N, E))[:, :, sel].transpose(2, 0, 1)
    orc = lwr_oracle(CFG2A_MAPS)
    phi_ref, jac_ref = orc.problem_update_batch(X_last[sel])
    dphi = wrap_angles(phi - phi_ref, [3, 4, 5])  # RPY entries: compare modulo 2 pi (the LWR base yaw sits on the atan2 cut)
    e_phi, e_jac = float(np.max(np.abs(dphi))), float(np.max(np.abs(jac - jac_ref)))
    out = {"max_abs_phi": e_phi, "max_abs_j": e_jac, "n": int(n), "tol": PARITY_TOL, "checked": "Phi / J device buffers of the last timed step vs the CPU oracle"}
    if not (e_phi <= PARITY_TOL and e_jac <= PARITY_TOL):
        raise RuntimeError("parity failure on the timed buffers: %r" % (out,))
    return out


def measure_cfg1(exo, capi):
    """BASELINE.json configs[0]: one configuration of example_ik.xml's problem -- CPU oracle us / evaluation beside the latency of
    problem.update(x) (host x in, Phi and J back on the host: H2D + one launch + D2H + synchronise)."""
    import numpy as np
    maps = [{"type": "EffFrame", "Type": "RPY", "EndEffector": [{"Link": "lwr_arm_6_link", "LinkOffset": CFG1_OFFSET}]}]
    orc = lwr_oracle(maps, add_collision_elements=True)
    X = random_configurations(orc.lower, orc.upper, 20000, 77)
    orc.problem_update_batch(X[:2000], discard=True)
    t0 = time.perf_counter()
    orc.problem_update_batch(X, discard=True)
    cpu_us = (time.perf_counter() - t0) / X.shape[0] * 1e6
    prob = exo.Setup.create_problem(cfg1_initializer())
    for i in range(20):
        prob.update(X[i])
    lat = []
    for i in range(200):
        t0 = time.perf_counter()
        prob.update(X[i])
        lat.append((time.perf_counter() - t0) * 1e6)
    phi_ref, jac_ref = lwr_oracle(maps).problem_update(X[199])
    err = max(float(np.max(np.abs(wrap_angles(prob.Phi - phi_ref, [3, 4, 5])))), float(np.max(np.abs(prob.jacobian - jac_ref))))
    if not err <= PARITY_TOL:
        raise RuntimeError("cfg1 parity failure: %g" % err)
    return {"workload": "example_ik.xml: EffFrame(RPY) lwr_arm_6_link + quaternion LinkOffset, one configuration",
            "cpu_us_per_eval": cpu_us, "cpu": "oracle, 1 thread, per-collision-shape tree elements included",
            "gpu_update_us": statistics.median(lat), "gpu_update_us_min": min(lat), "gpu": "problem.update(x): H2D + 1 launch + D2H + sync",
            "max_abs_err_vs_oracle": err}


# ---- multi-GPU configurations with a real exchange step (strong scaling) -----------------------------------------------------
def cfg3_initializer():
    frame = ("exotica/Frame", {"Link": "lwr_arm_7_link"})
    maps = [("exotica/EffFrame", {"Name": "Eff", "EndEffector": [frame]}), ("exotica/JointLimit", {"Name": "JL"})]
    return ("exotica/TimeIndexedProblem", {"Name": "cfg3", "PlanningScene": lwr_scene(), "Maps": maps, "T": 100, "tau": 0.05,
                                           "Cost": [("exotica/Task", {"Task": "Eff"}), ("exotica/Task", {"Task": "JL"})]})


def cfg5_spheres():
    import numpy as np
    sph = []
    rng = np.random.default_rng(5)
    for i in range(32):
        sph.append({"Link": "lwr_arm_%d_link" % (1 + i % 7), "LinkOffset": [0.0, 0.0, 0.02 * (i // 7)], "Radius": 0.05, "Group": "Robot"})
    for i in range(32):
        sph.append({"Link": "", "LinkOffset": [float(v) for v in rng.uniform([-0.8, -0.8, 0.0], [0.8, 0.8, 1.2])], "Radius": 0.05, "Group": "Obstacle"})
    return sph


def cfg5_initializer():
    maps = [("exotica/SphereCollision", {"Name": "Col", "Precision": 0.1, "EndEffector": [("exotica/Frame", s) for s in cfg5_spheres()]})]
    return ("exotica/TimeIndexedProblem", {"Name": "cfg5", "PlanningScene": lwr_scene(), "Maps": maps, "T": 100, "tau": 0.05,
                                           "Cost": [("exotica/Task", {"Task": "Col"})]})


def multi_start_trajectories(lim, B, T, N, seed=1234):
    import numpy as np
    rng = np.random.default_rng(seed)
    q0 = rng.uniform(lim[:, 0], lim[:, 1], size=(B, 1, N))
    return np.clip(q0 + np.cumsum(rng.normal(0, 0.05, size=(B, T, N)), axis=1), lim[:, 0] - 0.1, lim[:, 1] + 0.1)


def time_graph(capi, comm, stream, graph, reps, scalar):
    """Average milliseconds per launch of a captured step: CUDA events on the stream, max over ranks."""
    for _ in range(3):
        graph.launch(stream)
    comm.barrier(stream)
    ev0, ev1 = capi.Event(), capi.Event()
    ev0.record(stream)
    for _ in range(reps):
        graph.launch(stream)
    ev1.record(stream)
    ev1.synchronize()
    ms = ev0.elapsed_ms(ev1) / reps
    return scalar.reduce(comm, ms, capi.REDUCE_MAX)


def measure_multi_gpu(exo, capi, exd, comm, rank, world, reps=50, B=4096):
    """cfg3 / cfg5 sharded by trajectory over `world` GPUs (strong scaling: B x T evaluations in total).  Every step is one CUDA
    graph on one stream: the update kernel(s), the per-trajectory reduction and the NCCL collective, no host synchronisation."""
    import numpy as np
    host = exo._host
    out = {}
    scalar = DeviceScalar(capi)
    stream = capi.Stream()
    solo = capi.Comm(0, 1)  # rank 0 also runs the whole problem alone: the 1-GPU time of the same step in the same process
    for name, make in (("cfg3", cfg3_initializer), ("cfg5", cfg5_initializer)):
        prob = exo.Setup.create_problem(make())
        N, LP, LJ, T = prob.N, prob.length_Phi, prob.length_jacobian, prob.T
        lim = prob.get_scene().get_kinematic_tree().get_joint_limits()
        X = multi_start_trajectories(lim, B, T, N)
        backend = host.lib().exo_host_problem_backend(prob._h)
        prob.cost_batch(X[:2])  # pushes the cost description (rho, goal, W) into the backend through the host API

        def build(c, b0, b1, per_rank):
            nb = b1 - b0
            E = max(nb * T, 1)
            d_x = capi.DeviceBuffer(E * N)
            d_x.upload(np.ascontiguousarray(X[b0:b1].reshape(nb * T, N).T))
            d_phi, d_jac = capi.DeviceBuffer(E * LP), capi.DeviceBuffer(E * LJ * N)
            d_all = capi.DeviceBuffer(per_rank * c.world)
            d_min_eval, d_gmin = capi.DeviceBuffer(E), capi.DeviceBuffer(1)
            mine = d_all.ptr + 8 * c.rank * per_rank  # this rank's slot of the gathered array

            def step(with_collective):
                if name == "cfg5":
                    capi.check(capi.lib().exo_problem_set_min_distance_output(backend, d_min_eval.ptr))
                if nb > 0 and name == "cfg5":
                    prob.update_batch_device(d_x.ptr, nb * T, d_phi.ptr, d_jac.ptr, layout=exo.LAYOUT_SOA, stream=stream.ptr)
                if name == "cfg3":
                    # Phi and J of every (trajectory, timestep) stay in HBM for the caller; the costs are gathered.  Update + GetCost in
                    # one call: the generated kernel forms every evaluation's cost while its Phi is in registers
                    if nb > 0:
                        prob.update_cost_batch_device(d_x.ptr, nb, d_phi.ptr, d_jac.ptr, mine, layout=exo.LAYOUT_SOA, stream=stream.ptr)
                    if with_collective:
                        c.allgather(mine, d_all.ptr, per_rank, stream)
                elif with_collective:
                    c.min_distance_allgather(d_min_eval.ptr, nb, T, per_rank, d_all.ptr, d_gmin.ptr, stream)
                elif nb > 0:
                    capi.check(capi.lib().exo_trajectory_reduce(d_min_eval.ptr, nb, T, 0, capi.REDUCE_MIN, mine, stream.ptr))
            return (d_x, d_phi, d_jac, d_all, d_min_eval, d_gmin), step

        def capture(step, with_collective):
            with capi.Graph.capture(stream) as g:
                step(with_collective)
            return g

        b0, b1 = exd.shard_range(B, rank, world)
        per = exd.units_per_rank(B, world)
        keep, step = build(comm, b0, b1, per)
        step(True)  # allocates the library's scratch buffers and NCCL's connections before the capture
        stream.synchronize()
        g_full, g_nocoll = capture(step, True), capture(step, False)
        ms_full = time_graph(capi, comm, stream, g_full, reps, scalar)
        ms_nocoll = time_graph(capi, comm, stream, g_nocoll, reps, scalar)
        g_full.launch(stream)
        stream.synchronize()
        gathered = keep[3].download((per * world,))[:B]
        gmin = float(keep[5].download((1,))[0])
        res = {"trajectories": B, "T": T, "evals": B * T, "scaling": "strong", "ms_per_step": ms_full, "evals_per_s": B * T / (ms_full * 1e-3),
               "ms_without_collective": ms_nocoll, "collective_us": 1e3 * (ms_full - ms_nocoll),
               "step": ("update (Phi, J) with the cost epilogue + per-trajectory sum + all-gather(%d costs), one CUDA graph" % B) if name == "cfg3" else
                       ("update (park + pair kernel) + per-trajectory min + all-gather(%d distances) + global min, one CUDA graph" % B),
               "collective_path": "peer windows (one kernel over NVLink peer memory)" if comm.peer_collective_count() >= 0 and world > 1 else "NCCL"}
        comm.barrier(stream)
        if rank == 0:
            # the same step on ONE GPU (this one) for all B trajectories: the strong-scaling reference, measured in the same process
            keep1, step1 = build(solo, 0, B, B)
            step1(True)
            stream.synchronize()
            g1 = capture(step1, True)
            res["ms_per_step_1gpu"] = time_graph(capi, solo, stream, g1, reps, scalar)
            res["strong_scaling_efficiency"] = res["ms_per_step_1gpu"] / (world * ms_full)
            # parity of what was gathered, against the CPU oracle on a sample of trajectories
            sel = list(range(0, B, max(1, B // 32)))
            if name == "cfg3":
                orc = lwr_oracle([{"type": "EffFrame", "EndEffector": [{"Link": "lwr_arm_7_link"}]}, {"type": "JointLimit"}])
                ref = np.array([orc.trajectory_cost(X[b], [1.0, 1.0], ct=1.0 / 0.05 / T, W=np.ones(N))[0] for b in sel])
                res["gathered_cost_max_rel_err_vs_oracle"] = float(np.max(np.abs(gathered[sel] - ref) / np.maximum(1.0, np.abs(ref))))
            else:
                sph = cfg5_spheres()
                orc = lwr_oracle([])
                orc.request_frames(sph)
                rad = np.array([s["Radius"] for s in sph])
                ia = [i for i, s in enumerate(sph) if s["Group"] == "Obstacle"]
                ib = [i for i, s in enumerate(sph) if s["Group"] == "Robot"]
                ref = []
                for b in sel[:8]:
                    m = np.inf
                    for t in range(T):
                        orc.update(X[b, t])
                        p = orc.phi()[:, 9:12]
                        d = np.linalg.norm(p[ia][:, None, :] - p[ib][None, :, :], axis=2) - rad[ia][:, None] - rad[ib][None, :]
                        m = min(m, float(d.min()))
                    ref.append(m)
                res["gathered_min_distance_max_abs_err_vs_oracle"] = float(np.max(np.abs(gathered[sel[:8]] - np.array(ref))))
                res["global_min_distance"] = gmin
                res["global_min_equals_min_of_gathered"] = bool(gmin == float(np.min(gathered)))
        comm.barrier(stream)
        out[name] = res
    return out


def run_ours(args):
    import numpy as np
    import exotica_b200 as exo
    from exotica_b200 import _capi as capi
    from exotica_b200 import distributed as exd

    if capi.device_count() < 1:
        raise RuntimeError("bench.py needs a CUDA device: the product has no CPU path (use --impl reference for the CPU oracle)")
    comm, rank, world, local_rank = exd.init_from_env()
    # several ranks on one node: keep every rank's pinned pools and copy-issuing threads on its GPU's NUMA node (EXO_BENCH_NUMA=0: off)
    numa = exd.bind_to_gpu_numa_node(local_rank) if world > 1 and os.environ.get("EXO_BENCH_NUMA", "1") != "0" else None
    if rank == 0:
        from oracle import oracle
        oracle.use_native_build()  # CPU-baseline build flags; bit-identical results to the generic build

    E = args.configs
    problem = exo.Setup.create_problem(lwr_initializer())
    N, LP, LJ = problem.N, problem.length_Phi, problem.length_jacobian
    limits = problem.get_scene().get_kinematic_tree().get_joint_limits()
    # N_X_BUFFERS distinct device-resident SoA input buffers: together larger than L2, so no launch finds its inputs cached
    Xs = [random_configurations(limits[:, 0], limits[:, 1], E, seed=1234 + 2 + 1000 * rank + 17 * i) for i in range(N_X_BUFFERS)]
    d_xs = []
    for X in Xs:
        b = capi.DeviceBuffer(E * N)
        b.upload(X.T)  # [N][E]
        d_xs.append(b)
    d_phi = capi.DeviceBuffer(E * LP)       # [LP][E]
    d_jac = capi.DeviceBuffer(E * LJ * N)   # [LJ][N][E]
    stream = capi.Stream()
    scalar = DeviceScalar(capi)
    counter = [0]

    def step():
        problem.update_batch_device(d_xs[counter[0] % N_X_BUFFERS].ptr, E, d_phi.ptr, d_jac.ptr, layout=exo.LAYOUT_SOA, stream=stream.ptr)
        counter[0] += 1

    def barrier():
        comm.barrier(stream)
        capi.check(capi.lib().exo_device_synchronize())

    warmup = max(args.warmup, 3)
    for _ in range(warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    launches0 = problem.launch_count()
    ev0, ev1 = capi.Event(), capi.Event()
    sampler.start()
    barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        step()
    ev1.record(stream)
    barrier()
    sampler.stop()
    elapsed_ms = ev0.elapsed_ms(ev1)
    launches = problem.launch_count() - launches0
    timed_kernel = problem.last_kernel()
    last_buffer = (counter[0] - 1) % N_X_BUFFERS
    clock_note = "NVML polled every ~2 ms during the timed region"
    if len(sampler.samples) < 3:
        # the timed region was too short for NVML: sample during an untimed 0.4 s loop of the same kernel
        sampler.start()
        t_end = time.perf_counter() + 0.4
        while time.perf_counter() < t_end:
            for _ in range(50):
                step()
            stream.synchronize()
        sampler.stop()
        clock_note = "timed region shorter than the NVML period; sampled during an untimed 0.4 s loop of the same kernel right after it"
        # leave the buffers as the last TIMED step wrote them
        problem.update_batch_device(d_xs[last_buffer].ptr, E, d_phi.ptr, d_jac.ptr, layout=exo.LAYOUT_SOA, stream=stream.ptr)
        stream.synchronize()
    max_ms = scalar.reduce(comm, elapsed_ms, capi.REDUCE_MAX)
    value = world * E * args.steps / (max_ms * 1e-3)
    kernel_ms = max_ms / max(args.steps, 1)

    # ---- parity on the timed buffers (every rank checks its own; the oracle is the checker, outside the timed region) ----
    parity = check_parity(capi, Xs[last_buffer], d_phi, d_jac, E, N, LP, LJ, seed=99 + rank)
    worst = scalar.reduce(comm, max(parity["max_abs_phi"], parity["max_abs_j"]), capi.REDUCE_MAX)
    parity["max_abs_over_ranks"] = worst

    # ---- e2e: public API, host buffers, copies inside the timed region --------------------------------------------------
    px = exo.PinnedArray((E, N))
    pphi = exo.PinnedArray((E, LP))
    pjac = exo.PinnedArray((E, LJ, N))
    px.array[...] = Xs[0]
    for _ in range(2):
        problem.update_batch(px.array, layout=exo.LAYOUT_AOS_ROWMAJOR, out_phi=pphi.array, out_jacobian=pjac.array)
    e2e_steps = max(3, min(args.steps, 20))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        problem.update_batch(px.array, layout=exo.LAYOUT_AOS_ROWMAJOR, out_phi=pphi.array, out_jacobian=pjac.array)
    e2e_dt = time.perf_counter() - t0
    e2e_value = world * E * e2e_steps / scalar.reduce(comm, e2e_dt, capi.REDUCE_MAX)
    sel = np.arange(0, E, max(1, E // 1024))
    phi_ref, jac_ref = lwr_oracle(CFG2A_MAPS).problem_update_batch(Xs[0][sel])
    e2e_err = max(float(np.max(np.abs(wrap_angles(pphi.array[sel] - phi_ref, [3, 4, 5])))), float(np.max(np.abs(pjac.array[sel] - jac_ref))))
    if not e2e_err <= PARITY_TOL:
        raise RuntimeError("parity failure on the e2e host buffers: %g" % e2e_err)

    multi = None
    if world > 1:
        multi = measure_multi_gpu(exo, capi, exd, comm, rank, world)

    if rank != 0:
        comm.barrier(stream)
        comm.destroy()
        return

    peak, peak_src = measured_peak()
    achieved = BYTES_PER_EVAL * E / (kernel_ms * 1e-3) / 1e9
    traffic, traffic_info = recorded_traffic()
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "kernel": timed_kernel, "algorithmic_bytes_per_launch": BYTES_PER_EVAL * E, "launch_ms": kernel_ms, "peak_source": peak_src}
    if traffic:
        roofline["frac_on_dram_bytes"] = traffic / (kernel_ms * 1e-3) / 1e9 / peak
        roofline["traffic_source"] = (traffic_info or {}).get("source")
    from oracle import oracle
    flags = oracle.BUILD_FLAGS
    threads = host_threads()
    cpu_value, cpu_sample, _ = time_cpu_oracle(E, threads)
    cpu1_value, _, _ = time_cpu_oracle(E, 1, target_seconds=3.0)
    cfg1 = measure_cfg1(exo, capi)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": kernel_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(E), "layout": "SoA device buffers X[N][E] -> Phi[6][E], J[6][7][E]",
                       "l2": "no flush needed: every step streams %.0f MB of outputs and reads one of %d distinct %.1f MB input buffers in rotation "
                             "(inputs alone exceed the 126 MB L2)" % ((BYTES_PER_EVAL - 56) * E / 1e6, N_X_BUFFERS, 56 * E / 1e6),
                       "parallelism": "configurations sharded across %d GPU(s), no collective in the timed region" % world},
            "clocks": sampler.summary(clock_note), "parity": parity,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": E * N * 8, "d2h_bytes_per_step": E * (LP + LJ * N) * 8,
                    "steps": e2e_steps, "api": "exotica_b200.PlanningProblem.update_batch (pinned host X[E][N] -> Phi[E][6], J[E][6][7])", "numa_binding": numa,
                    "max_abs_err_vs_oracle": e2e_err},
            "gpu_launches": int(launches), "roofline": roofline,
            "cpu_baseline": {"value": cpu_value, "unit": UNIT, "cores": threads, "kind": "port", "sample": cpu_sample,
                             "single_thread_value": cpu1_value, "build": flags},
            "cfg1": cfg1}
    if multi is not None:
        line["multi_gpu"] = multi
    _emit(json.dumps(line))
    comm.barrier(stream)
    comm.destroy()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--configs", type=int, default=1 << 20, help="configurations per GPU per step")
    args = ap.parse_args()
    # stdout carries exactly one line, the JSON result: libraries that write to file descriptor 1 on their own (NCCL prints its
    # version banner there at communicator creation) are sent to stderr for the whole run
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    global _emit
    _emit = lambda text: os.write(json_fd, (text + "\n").encode())
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""bench.py -- batched FK + Jacobian evaluations/sec (LWR 7-DoF) on N B200s, with roofline and CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--configs E]

Workload (BASELINE.json configs[1], "cfg2a"): LWR 7-DoF, EffFrame(RPY) on lwr_arm_7_link, E = 2^20 seeded random
configurations per GPU per step, FP64.  One step = one pass of the hot path over that batch = one kernel launch.
Weak scaling: every rank evaluates its own 2^20 configurations, no data-path collective.

  value     whole-job evaluations/s with X already resident in HBM in the kernel's native SoA layout, timed with CUDA
            events on the launching stream over exactly K steps, max over ranks.
  e2e       the same metric through the public API with HOST buffers (pinned, evaluation-major as the reference exposes
            them): H2D of X and D2H of Phi and J inside the timed region, every step.
  roofline  achieved = 440 algorithmic bytes/eval (read q 56 + write Phi 48 + write J 336)
            * evals per launch / average launch duration, against the measured HBM copy bandwidth in MEASURED_PEAKS.json.
  cpu_baseline  the CPU oracle (restated reference algorithm, one configuration at a time) on this box's host cores.

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


_emit = print  # replaced in main(): the JSON line goes to the original stdout


def workload_name(E):
    return "cfg2a: LWR 7-DoF EffFrame(RPY) lwr_arm_7_link, %d random configurations per GPU per step" % E


def lwr_initializer():
    scene = ("exotica/Scene", {"Name": "Scene", "JointGroup": "arm", "URDF": os.path.join(ROBOTS, "lwr_simplified.urdf"),
                               "SRDF": os.path.join(ROBOTS, "lwr_simplified.srdf")})
    eff = ("exotica/EffFrame", {"Name": "Position", "Type": "RPY", "EndEffector": [("exotica/Frame", {"Link": "lwr_arm_7_link"})]})
    return ("exotica/UnconstrainedEndPoseProblem", {"Name": "bench", "PlanningScene": scene, "Maps": [eff],
                                                    "Cost": [("exotica/Task", {"Task": "Position"})]})


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
                return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


def recorded_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture, if any."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        try:
            with open(path) as f:
                return json.load(f).get("dram_bytes_per_launch")
        except Exception:
            return None
    return None


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


def time_cpu_oracle(E_request, threads, target_seconds=12.0, seed=1234 + 2):
    """Times the CPU oracle (restated reference algorithm incl. the per-collision-shape tree elements the reference
    updates, scene.cpp:795-821) on a bounded sample of the workload.  Returns (evals/s, sample description, E used)."""
    import numpy as np
    from oracle import oracle
    tree = oracle.load_robot(os.path.join(ROBOTS, "lwr_simplified.urdf"), os.path.join(ROBOTS, "lwr_simplified.srdf"), "arm",
                             add_collision_elements=True)
    orc = oracle.Oracle(tree)
    orc.set_problem([{"type": "EffFrame", "Type": "RPY", "EndEffector": [{"Link": "lwr_arm_7_link"}]}])
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
    sample = "%d pass(es) over %d of the step's configurations, %d thread(s), outputs discarded, %.2f s" % (passes, E, threads, dt)
    return passes * E / dt, sample, E


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
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
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": samples[-1]},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    _emit(json.dumps(line))


def run_ours(args):
    import numpy as np
    import torch
    import exotica_b200 as exo

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product has no CPU path (use --impl reference for the CPU oracle)")
    torch.cuda.set_device(local_rank)
    exo._capi.check(exo._capi.lib().exo_set_device(local_rank))
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # NCCL's version / debug lines must not share stdout with the JSON line
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))

    E = args.configs
    problem = exo.Setup.create_problem(lwr_initializer())
    N, LP, LJ = problem.N, problem.length_Phi, problem.length_jacobian
    limits = problem.get_scene().get_kinematic_tree().get_joint_limits()
    X = random_configurations(limits[:, 0], limits[:, 1], E, seed=1234 + 2 + 1000 * rank)

    # device-resident SoA buffers (torch is plumbing: memory, stream, events)
    dev = torch.device("cuda", local_rank)
    d_x = torch.from_numpy(np.ascontiguousarray(X.T)).to(dev)          # [N][E]
    d_phi = torch.empty((LP, E), dtype=torch.float64, device=dev)      # [LP][E]
    d_jac = torch.empty((LJ, N, E), dtype=torch.float64, device=dev)   # [LJ][N][E]
    stream = torch.cuda.current_stream()
    sptr = ctypes_ptr(stream.cuda_stream)

    def step():
        problem.update_batch_device(d_x.data_ptr(), E, d_phi.data_ptr(), d_jac.data_ptr(), layout=exo.LAYOUT_SOA, stream=sptr)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    launches0 = problem.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.start()
    barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        step()
    ev1.record(stream)
    barrier()
    sampler.stop()
    elapsed_ms = ev0.elapsed_time(ev1)
    launches = problem.launch_count() - launches0
    clock_note = "NVML polled every ~2 ms during the timed region"
    if len(sampler.samples) < 3:
        # the timed region was too short for NVML: sample during an untimed 0.4 s loop of the same kernel
        sampler.start()
        t_end = time.perf_counter() + 0.4
        while time.perf_counter() < t_end:
            for _ in range(50):
                step()
            torch.cuda.synchronize()
        sampler.stop()
        clock_note = "timed region shorter than the NVML period; sampled during an untimed 0.4 s loop of the same kernel right after it"
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    max_ms = float(t.item())
    value = world * E * args.steps / (max_ms * 1e-3)
    kernel_ms = max_ms / max(args.steps, 1)

    # parity spot-check on the timed buffers is part of tests/ (tests/test_gpu_parity.py); here only a sanity check
    phi_host = d_phi[:, :4].cpu().numpy()
    if not np.isfinite(phi_host).all():
        raise RuntimeError("non-finite outputs")

    # ---- e2e: public API, host buffers, copies inside the timed region --------------------------------------------------
    px = exo.PinnedArray((E, N))
    pphi = exo.PinnedArray((E, LP))
    pjac = exo.PinnedArray((E, LJ, N))
    px.array[...] = X
    for _ in range(2):
        problem.update_batch(px.array, layout=exo.LAYOUT_AOS_ROWMAJOR, out_phi=pphi.array, out_jacobian=pjac.array)
    e2e_steps = max(3, min(args.steps, 20))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        problem.update_batch(px.array, layout=exo.LAYOUT_AOS_ROWMAJOR, out_phi=pphi.array, out_jacobian=pjac.array)
    e2e_dt = time.perf_counter() - t0
    te = torch.tensor([e2e_dt], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * E * e2e_steps / float(te.item())
    e2e_checksum = float(pphi.array[:8].sum())

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak()
    achieved = BYTES_PER_EVAL * E / (kernel_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": recorded_traffic(),
                "kernel": problem.last_kernel(), "algorithmic_bytes_per_launch": BYTES_PER_EVAL * E, "launch_ms": kernel_ms, "peak_source": peak_src}
    threads = host_threads()
    cpu_value, cpu_sample, _ = time_cpu_oracle(E, threads)
    cpu1_value, _, _ = time_cpu_oracle(E, 1, target_seconds=3.0)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": kernel_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(E), "layout": "SoA device buffers X[N][E] -> Phi[6][E], J[6][7][E]",
                       "l2": "no flush: every step streams %.0f MB (inputs+outputs), larger than the 126 MB L2" % (BYTES_PER_EVAL * E / 1e6),
                       "parallelism": "configurations sharded across %d GPU(s), no collective" % world},
            "clocks": sampler.summary(clock_note),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": E * N * 8, "d2h_bytes_per_step": E * (LP + LJ * N) * 8,
                    "steps": e2e_steps, "api": "exotica_b200.PlanningProblem.update_batch (pinned host X[E][N] -> Phi[E][6], J[E][6][7])",
                    "checksum": e2e_checksum},
            "gpu_launches": int(launches), "roofline": roofline,
            "cpu_baseline": {"value": cpu_value, "unit": UNIT, "cores": threads, "kind": "port", "sample": cpu_sample,
                             "single_thread_value": cpu1_value}}
    _emit(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def ctypes_ptr(v):
    import ctypes
    return ctypes.c_void_p(int(v))


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
    global _emit
    _emit = lambda text: os.write(json_fd, (text + "\n").encode())
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

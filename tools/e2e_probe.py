import sys, time, os
sys.path.insert(0, '/root/repo')
import numpy as np, exotica_b200 as exo, bench
problem = exo.Setup.create_problem(bench.lwr_initializer())
E = 1 << 20
N, LP, LJ = problem.N, problem.length_Phi, problem.length_jacobian
lim = problem.get_scene().get_kinematic_tree().get_joint_limits()
X = bench.random_configurations(lim[:, 0], lim[:, 1], E, 3)
px, pphi, pjac = exo.PinnedArray((E, N)), exo.PinnedArray((E, LP)), exo.PinnedArray((E, LJ, N))
px.array[...] = X
for _ in range(3):
    problem.update_batch(px.array, layout=exo.LAYOUT_AOS_ROWMAJOR, out_phi=pphi.array, out_jacobian=pjac.array)
t0 = time.perf_counter()
for _ in range(10):
    problem.update_batch(px.array, layout=exo.LAYOUT_AOS_ROWMAJOR, out_phi=pphi.array, out_jacobian=pjac.array)
dt = (time.perf_counter() - t0) / 10
print("chunk MiB", os.environ.get("EXO_PIPE_CHUNK_MIB"), "ms/step %.3f" % (dt * 1e3), "evals/s %.4g" % (E / dt), "GB/s %.1f" % (E * (N + LP + LJ * N) * 8 / dt / 1e9))
# raw D2H bandwidth for reference
import torch
d = torch.empty(402653184 // 8, dtype=torch.float64, device="cuda")
h = torch.empty(402653184 // 8, dtype=torch.float64).pin_memory()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    h.copy_(d, non_blocking=True)
torch.cuda.synchronize()
print("raw D2H 402 MB: %.3f ms  %.1f GB/s" % ((time.perf_counter() - t0) / 5 * 1e3, 0.402653184 / ((time.perf_counter() - t0) / 5)))

"""PCIe probe: D2H bandwidth of one 402 MB copy vs. the same bytes split over 2 / 4 concurrent streams, with and without a
concurrent 58 MB H2D copy (what one e2e step of bench.py moves)."""
import time
import torch

n = 402653184 // 8
d = torch.empty(n, dtype=torch.float64, device="cuda")
h = torch.empty(n, dtype=torch.float64).pin_memory()
dx = torch.empty(58720256 // 8, dtype=torch.float64, device="cuda")
hx = torch.empty(58720256 // 8, dtype=torch.float64).pin_memory()
streams = [torch.cuda.Stream() for _ in range(4)]
up = torch.cuda.Stream()


def run(k, with_h2d, reps=6):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        if with_h2d:
            with torch.cuda.stream(up):
                dx.copy_(hx, non_blocking=True)
        step = n // k
        for i in range(k):
            with torch.cuda.stream(streams[i]):
                h[i * step:(i + 1) * step].copy_(d[i * step:(i + 1) * step], non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    print("D2H 402 MB over %d stream(s)%s: %.3f ms  %.1f GB/s" % (k, " + H2D 58 MB" if with_h2d else "", dt * 1e3, 0.402653184 / dt))


for k in (1, 2, 4):
    run(k, False)
for k in (1, 2, 4):
    run(k, True)

"""Host cost of one eager per-op call (Python + ctypes + launch), and the eager decode step rate."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from autoawq_b200 import ext  # noqa: E402

dev = torch.device("cuda:0")
rep = bench.Replica(dev, 1, layers=32)
for _ in range(3):
    rep.step(rep.h)
torch.cuda.synchronize()
t0 = time.perf_counter()
n = 20
for _ in range(n):
    rep.step(rep.h)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / n
print(f"eager per-op step: {dt * 1e3:.3f} ms = {1 / dt:.1f} tok/s, {dt / rep.launches_per_step * 1e6:.2f} us per call (224 calls)")
x = torch.randn((1, 128), device=dev, dtype=torch.float16)
qw = torch.zeros((128, 32), dtype=torch.int32, device=dev)
qz = torch.zeros((1, 32), dtype=torch.int32, device=dev)
sc = torch.ones((1, 256), dtype=torch.float16, device=dev)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3000):
    ext.gemm_forward_cuda(x, qw, sc, qz, 8)
torch.cuda.synchronize()
print(f"tiny gemm_forward_cuda call: {(time.perf_counter() - t0) / 3000 * 1e6:.2f} us host time per call")

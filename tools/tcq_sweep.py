"""Small-M kernel (gemm_tcq_kernel) against the paths it replaces: per shape and M, device time of
  old = register-staged tcgen05 kernel (knob 19 = 1; M <= 8: the persistent GEMV) and new = TMA-staged kernel
(knob 19 = 0, GEMV threshold 0 so that M <= 8 runs it too); plus the new kernel's per-CTA phase timeline (knob 3 = 9).
CUDA graph over a rotating weight pool > L2, CUDA events.  Prints one JSON."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from autoawq_b200 import ext  # noqa: E402
from autoawq_b200._cabi import lib  # noqa: E402
from tools.gpu_probe_lib import time_kernel  # noqa: E402

dev = torch.device("cuda:0")
peaks = bench.measured_peaks()
G = 128
out = {"peaks": peaks, "rows": {}, "timeline": {}}
shapes = [(4096, 4096), (4096, 14336), (4096, 28672), (14336, 4096)]
Ms = [int(v) for v in os.environ.get("TCQ_MS", "2,4,8,16,32,64,128").split(",")]


def timeline(x, qw, sc, qz):
    import ctypes

    ext.set_knob(3, 9)
    ext.linear_forward("gemm", x, qw, sc, qz, G)
    torch.cuda.synchronize()
    buf = (ctypes.c_uint64 * (148 * 8))()
    lib.b200awq_debug_read(buf, ctypes.sizeof(buf))
    ext.set_knob(3, 0)
    a = np.frombuffer(buf, dtype=np.uint64).reshape(148, 8).astype(np.int64)
    t0 = a[:, 0].min()
    rel = (a[:, :7] - t0) / 1e3
    names = ["entry", "setup", "first_q", "producers_done", "mma_done", "acc_drained", "epilogue_done"]
    return {n: {"min": round(float(rel[:, i].min()), 2), "med": round(float(np.median(rel[:, i])), 2),
                "max": round(float(rel[:, i].max()), 2)} for i, n in enumerate(names)} | {
        "segments_max": int(a[:, 7].max())}


for (K, N) in shapes:
    wbytes = K * N // 2 + (K // G) * N * 2 + (K // G) * N // 2
    nbuf = max(3, int(400e6 // wbytes) + 1)
    qw = [torch.randint(-2**31, 2**31 - 1, (K, N // 8), dtype=torch.int32, device=dev) for _ in range(nbuf)]
    qz = [torch.randint(-2**31, 2**31 - 1, (K // G, N // 8), dtype=torch.int32, device=dev) for _ in range(nbuf)]
    sc = [(torch.rand((K // G, N), device=dev) * 0.01 + 0.001).half() for _ in range(nbuf)]
    for M in Ms:
        x = torch.randn((M, K), device=dev, dtype=torch.float16)
        row = {}
        for name, k19, k2, k21 in (("old", 1, 8, 0), ("new", 0, 0, 1), ("aligned", 0, 0, 2)):
            ext.set_knob(19, k19)
            ext.set_knob(2, k2)
            ext.set_knob(21, k21)
            us = time_kernel(torch, lambda i: ext.linear_forward("gemm", x, qw[i], sc[i], qz[i], G), nbuf, iters=100, warm=5)
            b = bench.linear_bytes(K, N, M)
            row[name] = {"us": round(us, 2), "gbs": round(b / us / 1e3, 1), "frac_hbm": round(b / us / 1e3 / peaks["hbm_gbs"], 4),
                         "tflops": round(2.0 * M * K * N / us / 1e6, 1)}
        out["rows"][f"{K}x{N} M={M}"] = row
        if M in (16, 128):
            ext.set_knob(19, 0)
            ext.set_knob(2, 0)
            out["timeline"][f"{K}x{N} M={M}"] = timeline(x, qw[0], sc[0], qz[0])
    ext.set_knob(19, 0)
    ext.set_knob(2, 8)
    ext.set_knob(21, 0)
    del qw, qz, sc
    torch.cuda.empty_cache()
print(json.dumps(out, indent=1))

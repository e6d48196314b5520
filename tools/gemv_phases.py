"""Per-CTA phase timeline of the persistent GEMV (knob 3 + b200awq_debug_read): where do the ~5 us of fixed
per-launch cost go?  Prints median / max of each phase over the 148 CTAs for the four Llama-3-8B shapes."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autoawq_b200 import ext  # noqa: E402
from autoawq_b200._cabi import lib  # noqa: E402

dev = torch.device("cuda:0")
G = 128
MS = [int(a) for a in sys.argv[1:]] or [1]
for M, (K, N) in [(m, s) for m in MS for s in [(4096, 4096), (4096, 6144), (4096, 28672), (14336, 4096)]]:
    nbuf = max(3, int(400e6 // (K * N // 2)) + 1)
    qw = [torch.randint(-2**31, 2**31 - 1, (K, N // 8), dtype=torch.int32, device=dev) for _ in range(nbuf)]
    qz = [torch.randint(-2**31, 2**31 - 1, (K // G, N // 8), dtype=torch.int32, device=dev) for _ in range(nbuf)]
    sc = [(torch.rand((K // G, N), device=dev) * 0.01 + 0.001).half() for _ in range(nbuf)]
    x = torch.randn((M, K), device=dev, dtype=torch.float16)
    ext.set_knob(3, 1)
    rows = []
    for it in range(3 * nbuf):
        ext.linear_forward("gemm", x, qw[it % nbuf], sc[it % nbuf], qz[it % nbuf], G)
        buf = np.zeros((256, 8), dtype=np.uint64)
        lib.b200awq_debug_read(buf.ctypes.data_as(ctypes.c_void_p), buf.nbytes)
        if it >= nbuf:
            rows.append(buf[:148].astype(np.int64))
    ext.set_knob(3, 0)
    a = np.stack(rows)  # [iters, 148, 8]
    t0 = a[:, :, 0].min(axis=1, keepdims=True)  # earliest CTA entry of the launch
    names = ["entry(skew)", "pdl_wait done", "first tile landed", "warp0 loop done", "all warps done",
             "REDs issued+barrier", "tickets+barrier", "finalise done"]
    print(f"M={M} K={K} N={N}: ns since the first CTA entered (median over CTAs, max over CTAs), median over {a.shape[0]} launches")
    for i, nm in enumerate(names):
        d = a[:, :, i] - t0
        print(f"   {nm:20s} median {np.median(np.median(d, axis=1)):8.0f}   max {np.median(d.max(axis=1)):8.0f}")

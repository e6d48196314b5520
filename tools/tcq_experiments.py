"""Timing decomposition of the small-M kernel (knob 20; outputs invalid while set): full kernel, without the producers'
dequant + stores, without the proxy fence; M = 16 on the four Llama-3-8B shapes.  Prints one JSON."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autoawq_b200 import ext  # noqa: E402
from tools.gpu_probe_lib import time_kernel  # noqa: E402

dev = torch.device("cuda:0")
G = 128
out = {}
for (K, N) in [(4096, 4096), (4096, 28672), (14336, 4096)]:
    wbytes = K * N // 2
    nbuf = max(3, int(400e6 // wbytes) + 1)
    qw = [torch.randint(-2**31, 2**31 - 1, (K, N // 8), dtype=torch.int32, device=dev) for _ in range(nbuf)]
    qz = [torch.randint(-2**31, 2**31 - 1, (K // G, N // 8), dtype=torch.int32, device=dev) for _ in range(nbuf)]
    sc = [(torch.rand((K // G, N), device=dev) * 0.01 + 0.001).half() for _ in range(nbuf)]
    for M in (16, 64):
        x = torch.randn((M, K), device=dev, dtype=torch.float16)
        row = {}
        for name, k20 in (("full", 0), ("no_dequant_no_sts", 1), ("no_mma", 4), ("one_mma", 8), ("no_mma_no_dequant", 5)):
            ext.set_knob(20, k20)
            row[name] = round(time_kernel(torch, lambda i: ext.linear_forward("gemm", x, qw[i], sc[i], qz[i], G), nbuf,
                                          iters=100, warm=5), 2)
        ext.set_knob(20, 0)
        for pf in (8, 16, 32, 64):      # HBM -> L2 prefetch distance in k-step pairs (knob 22)
            ext.set_knob(22, pf)
            row[f"l2_prefetch_{pf}"] = round(time_kernel(torch, lambda i: ext.linear_forward("gemm", x, qw[i], sc[i], qz[i], G),
                                                         nbuf, iters=100, warm=5), 2)
        ext.set_knob(22, 0)
        out[f"{K}x{N} M={M}"] = row
    del qw, qz, sc
    torch.cuda.empty_cache()
print(json.dumps(out, indent=1))

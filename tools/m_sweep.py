"""M sweep of the per-op kernels (BASELINE config 3: M in {8..4096} on 4096x4096 and 4096x14336) + the other layouts
+ the dequant kernel.  CUDA graph of launches rotating over a weight pool > L2, CUDA events.  Prints one JSON."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from autoawq_b200 import ext  # noqa: E402
from tools.gpu_probe_lib import time_kernel  # noqa: E402

dev = torch.device("cuda:0")
peaks = bench.measured_peaks()
G = 128
out = {"peaks": peaks, "gemm_layout": {}, "other_layouts": {}, "dequant": {}}
for (K, N) in [(4096, 4096), (4096, 14336), (4096, 28672), (14336, 4096)]:
    wbytes = K * N // 2 + (K // G) * N * 2 + (K // G) * N // 2
    nbuf = max(3, int(400e6 // wbytes) + 1)
    qw = [torch.randint(-2**31, 2**31 - 1, (K, N // 8), dtype=torch.int32, device=dev) for _ in range(nbuf)]
    qz = [torch.randint(-2**31, 2**31 - 1, (K // G, N // 8), dtype=torch.int32, device=dev) for _ in range(nbuf)]
    sc = [(torch.rand((K // G, N), device=dev) * 0.01 + 0.001).half() for _ in range(nbuf)]
    Ms = [1, 2, 4, 8, 16, 32, 64, 128, 256, 1024, 4096] if N != 28672 else [1, 8, 16, 64]
    for M in Ms:
        x = torch.randn((M, K), device=dev, dtype=torch.float16)
        us = time_kernel(torch, lambda i: ext.linear_forward("gemm", x, qw[i], sc[i], qz[i], G), nbuf,
                         iters=100 if M <= 256 else 12, warm=5)
        b, fl = bench.linear_bytes(K, N, M), 2.0 * M * K * N
        out["gemm_layout"][f"{K}x{N} M={M}"] = {
            "us": round(us, 2), "gbs": round(b / us / 1e3, 1), "frac_hbm": round(b / us / 1e3 / peaks["hbm_gbs"], 4),
            "tflops": round(fl / us / 1e6, 1), "frac_tensor_sustained": round(fl / us / 1e6 / peaks["bf16_tflops_sustained"], 4)}
    us = time_kernel(torch, lambda i: ext.dequantize_weights_cuda(qw[i], sc[i], qz[i]), nbuf, iters=50, warm=5)
    db = wbytes + 2 * K * N
    out["dequant"][f"{K}x{N}"] = {"us": round(us, 2), "gbs": round(db / us / 1e3, 1),
                                  "frac_hbm": round(db / us / 1e3 / peaks["hbm_gbs"], 4)}
    del qw, qz, sc
    torch.cuda.empty_cache()
# the other two checkpoint layouts at decode sizes (random bit patterns are valid packed tensors; scales kept tiny)
from autoawq_b200.packing import calculate_zeros_width  # noqa: E402

for (K, N) in [(4096, 4096), (14336, 4096)]:
    zw = calculate_zeros_width(K, G)
    wbytes = K * N // 2 + (K // G) * N * 2 + (K // G) * N // 2
    nbuf = max(3, int(400e6 // wbytes) + 1)
    vw = [torch.randint(-2**31, 2**31 - 1, (N, K // 8), dtype=torch.int32, device=dev) for _ in range(nbuf)]
    vz = [torch.randint(-2**31, 2**31 - 1, (N, zw), dtype=torch.int32, device=dev) for _ in range(nbuf)]
    vs = [(torch.rand((N, zw * 8), device=dev) * 0.01 + 0.001).half() for _ in range(nbuf)]
    fw = [torch.randint(-2**15, 2**15 - 1, (N // 4, K), dtype=torch.int16, device=dev) for _ in range(nbuf)]
    fs = [(torch.rand((zw * 8, N), device=dev) * 0.01 + 0.001).half() for _ in range(nbuf)]
    fz = [(-torch.rand((zw * 8, N), device=dev) * 0.05).half() for _ in range(nbuf)]
    for M in (1, 8, 64):
        x = torch.randn((M, K), device=dev, dtype=torch.float16)
        b = bench.linear_bytes(K, N, M)
        for name, fn in (("gemv", lambda i: ext.linear_forward("gemv", x, vw[i], vs[i], vz[i], G)),
                         ("fast", lambda i: ext.linear_forward("fast", x, fw[i], fs[i], fz[i], G))):
            us = time_kernel(torch, fn, nbuf, iters=100, warm=5)
            out["other_layouts"][f"{name} {K}x{N} M={M}"] = {"us": round(us, 2), "gbs": round(b / us / 1e3, 1),
                                                             "frac_hbm": round(b / us / 1e3 / peaks["hbm_gbs"], 4)}
    del vw, vz, vs, fw, fs, fz
    torch.cuda.empty_cache()
print(json.dumps(out, indent=1))

"""Small launch sequences for ncu captures (run under `ncu ... python tools/ncu_target.py <what>`).
what = gemv : 4096x4096 g128 M=1 GEMV over a rotating 400 MB weight pool (DRAM-resident)
       gemv_big : 4096x28672 M=1
       gemm : 4096x4096x4096 tcgen05 GEMM
       program : a 4-layer Llama-3-8B-shaped decode program (16 linears + glue in one program_kernel launch), run 3x
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autoawq_b200 import ext  # noqa: E402

dev = torch.device("cuda:0")
what = sys.argv[1] if len(sys.argv) > 1 else "gemv"
if what == "program":
    from autoawq_b200.program import DecodeProgram

    G, H, I = 128, 4096, 14336
    LIN = [("qkv", H, 6144), ("o", H, H), ("gate_up", H, 2 * I), ("down", I, H)]

    def rand_linear(K, N):
        return (torch.randint(-2**31, 2**31 - 1, (K, N // 8), dtype=torch.int32, device=dev),
                ((torch.rand((K // G, N), device=dev) * 0.5 + 0.75) / (6.1 * K**0.5)).half(),
                torch.randint(-2**31, 2**31 - 1, (K // G, N // 8), dtype=torch.int32, device=dev))

    ws = [{n: rand_linear(K, N) for n, K, N in LIN} for _ in range(4)]
    nw = torch.ones(H, dtype=torch.float16, device=dev)
    x = torch.randn((1, H), device=dev, dtype=torch.float16)
    xn = torch.empty((1, H), dtype=torch.float16, device=dev)
    act = torch.empty((1, I), dtype=torch.float16, device=dev)
    prog = DecodeProgram()
    for lw in ws:
        prog.layernorm_forward_cuda(x, nw, xn, 1e-5)
        qkv = prog.gemm_forward_cuda(xn, *lw["qkv"], 8)
        o = prog.gemm_forward_cuda(qkv[:, :H], *lw["o"], 8)
        prog.layernorm_forward_cuda(o, nw, xn, 1e-5)
        gu = prog.gemm_forward_cuda(xn, *lw["gate_up"], 8)
        prog.silu_and_mul(act, gu)
        x = prog.gemm_forward_cuda(act, *lw["down"], 8)
    prog.build()
    assert prog.fused
    for _ in range(3):
        prog.run()
    torch.cuda.synchronize()
    sys.exit(0)
if what == "layer":
    # one Llama-3-8B layer's four linears at M = 1, weights rotated through a pool > L2: the launch mix bench.py times
    G = 128
    shapes = [(4096, 6144), (4096, 4096), (4096, 28672), (14336, 4096)]
    pool = []
    for rep in range(3):
        for (K, N) in shapes:
            pool.append((torch.randint(-2**31, 2**31 - 1, (K, N // 8), dtype=torch.int32, device=dev),
                         (torch.rand((K // G, N), device=dev) * 0.01 + 0.001).half(),
                         torch.randint(-2**31, 2**31 - 1, (K // G, N // 8), dtype=torch.int32, device=dev),
                         torch.randn((1, K), device=dev, dtype=torch.float16)))
    for qw, sc, qz, x in pool:
        ext.linear_forward("gemm", x, qw, sc, qz, G)
    torch.cuda.synchronize()
    sys.exit(0)
K, N, M = {"gemv": (4096, 4096, 1), "gemv_big": (4096, 28672, 1), "gemm": (4096, 4096, 4096),
           "gemm64": (4096, 14336, 64), "gemm16": (4096, 4096, 16), "gemv8": (4096, 4096, 8),
           "prefill_big": (4096, 14336, 4096), "gemm16_big": (4096, 28672, 16), "gemm64_big": (4096, 28672, 64)}[what]
G = 128
wbytes = K * N // 2
nbuf = max(3, int(400e6 // wbytes) + 1)
qw = [torch.randint(-2**31, 2**31 - 1, (K, N // 8), dtype=torch.int32, device=dev) for _ in range(nbuf)]
qz = [torch.randint(-2**31, 2**31 - 1, (K // G, N // 8), dtype=torch.int32, device=dev) for _ in range(nbuf)]
sc = [(torch.rand((K // G, N), device=dev) * 0.01 + 0.001).half() for _ in range(nbuf)]
x = torch.randn((M, K), device=dev, dtype=torch.float16)
for i in range(2 * nbuf):
    ext.linear_forward("gemm", x, qw[i % nbuf], sc[i % nbuf], qz[i % nbuf], G)
torch.cuda.synchronize()

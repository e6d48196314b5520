"""Per-op phase timeline of the decode-program kernel (knob 3 = 2 + b200awq_debug_read): where does the time
between two linears go?  Builds `--layers` Llama-3-8B-shaped blocks (random packed weights), runs the program a few
times and prints, per kernel op, the phase boundaries (ns, relative to the op's begin on CTA 0; median over the
first 8 CTAs) plus the op's duration."""
import argparse
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autoawq_b200 import ext  # noqa: E402
from autoawq_b200._cabi import lib  # noqa: E402
from autoawq_b200.program import DecodeProgram  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--layers", type=int, default=4)
ap.add_argument("--kind", type=int, default=0, help="knob 14: 0 = default (stream), 1 = split-K kernel")
a = ap.parse_args()
ext.set_knob(14, a.kind)
dev = torch.device("cuda:0")
G, H, I = 128, 4096, 14336
LIN = [("qkv", H, 6144), ("o", H, H), ("gate_up", H, 2 * I), ("down", I, H)]


def rand_linear(K, N):
    qw = torch.randint(-2**31, 2**31 - 1, (K, N // 8), dtype=torch.int32, device=dev)
    qz = torch.randint(-2**31, 2**31 - 1, (K // G, N // 8), dtype=torch.int32, device=dev)
    s = ((torch.rand((K // G, N), device=dev) * 0.5 + 0.75) / (6.1 * K**0.5)).half()
    return qw, s, qz


ws = [{n: rand_linear(K, N) for n, K, N in LIN} for _ in range(a.layers)]
nw = torch.ones(H, dtype=torch.float16, device=dev)
h = torch.randn((1, H), device=dev, dtype=torch.float16)
xn = torch.empty((1, H), dtype=torch.float16, device=dev)
act = torch.empty((1, I), dtype=torch.float16, device=dev)
prog = DecodeProgram()
x = h
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
nops = prog.kernel_ops
ext.set_knob(3, 2)
runs = []
for it in range(6):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    prog.run()
    e1.record()
    torch.cuda.synchronize()
    print(f"run {it}: {e0.elapsed_time(e1) * 1e3:.1f} us, abort record {DecodeProgram.abort_record()}")
    ext.set_knob(3, 2)
    buf = np.zeros((32, 8, 8), dtype=np.uint64)
    lib.b200awq_debug_read(buf.ctypes.data_as(ctypes.c_void_p), buf.nbytes)
    if it >= 2:
        runs.append(buf[: min(nops, 32)].astype(np.int64))
ext.set_knob(3, 0)
r = np.stack(runs)  # [runs, ops, cta, 8]
stream = prog.kind == "stream"
names = (["begin", "row polled", "x staged", "1st chunk", "warp0 done", "all warps", "published", "-"] if stream else
         ["begin", "prev done", "x staged", "1st tile", "warp0 done", "all warps", "sums added", "published"])
print("program kind:", prog.kind)
print(f"{nops} kernel ops; ns relative to the op's begin on its earliest CTA (median over 8 CTAs, median over runs)")
print("op  shape        " + " ".join(f"{n:>10s}" for n in names) + "   next-begin")
for op in range(min(nops, 32)):
    t0 = r[:, op, :, 0].min(axis=1)[:, None, None]
    d = np.median(np.median(r[:, op] - t0, axis=1), axis=0)
    nb = ""
    if op + 1 < min(nops, 32):
        nb = f"{np.median(np.median(r[:, op + 1, :, 0] - t0[:, :, 0], axis=1)):10.0f}"
    n, K, N = LIN[op % 4]
    print(f"{op:2d}  {n:8s}     " + " ".join(f"{v:10.0f}" for v in d) + "   " + nb)
tot = r[:, min(nops, 32) - 1, :, 6 if stream else 7].max(axis=1) - r[:, 0, :, 0].min(axis=1)
print("span first begin -> last published (ns):", np.median(tot), " per op:", np.median(tot) / min(nops, 32))

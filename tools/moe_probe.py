"""Mixtral-shaped MoE block at decode (bs = 1, 8 experts, top-2; awq/modules/fused/moe.py apply_moe_weights):
time route -> align -> gate|up grouped GEMM -> silu*mul -> down grouped GEMM (x routing weight) -> sum, as a CUDA
graph over NL distinct layers (weights >> L2), and report us per block and the HBM rate over the ACTIVE experts."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import awq_ext  # noqa: E402

dev = torch.device("cuda:0")
E, H, I, G, topk, NL = 8, 4096, 14336, 128, 2, 3
g = torch.Generator(device=dev).manual_seed(0)


def stacked(K, N):
    qw = torch.randint(-2**31, 2**31 - 1, (E, K, N // 8), dtype=torch.int32, device=dev, generator=g)
    qz = torch.randint(-2**31, 2**31 - 1, (E, K // G, N // 8), dtype=torch.int32, device=dev, generator=g)
    sc = ((torch.rand((E, K // G, N), device=dev, generator=g) * 0.5 + 0.75) / (6.1 * K**0.5)).half()
    return qw, sc, qz


layers = [(stacked(H, 2 * I), stacked(I, H), torch.randn((1, E), device=dev, generator=g)) for _ in range(NL)]
x = torch.randn((1, H), device=dev, dtype=torch.float16, generator=g)
tw = torch.empty((1, topk), dtype=torch.float32, device=dev)
tid = torch.empty((1, topk), dtype=torch.int32, device=dev)
src = torch.empty((1, topk), dtype=torch.int32, device=dev)
numel = topk
s_ids = torch.empty((numel + E * 15,), dtype=torch.int32, device=dev)
e_ids = torch.empty((numel + E,), dtype=torch.int32, device=dev)
npost = torch.empty((1,), dtype=torch.int32, device=dev)
act = torch.empty((1, topk, I), dtype=torch.float16, device=dev)


def block(w1, w2, gating, h):
    awq_ext.topk_softmax(tw, tid, src, gating)
    s_ids.fill_(numel)
    awq_ext.moe_alig_block_size(tid, E, 16, s_ids, e_ids, npost)
    gu = awq_ext.grouped_gemm_forward(h.view(1, 1, H), *w1, tw, s_ids, e_ids, npost, False, 8)
    awq_ext.silu_and_mul(act, gu)
    out = awq_ext.grouped_gemm_forward(act, *w2, tw, s_ids, e_ids, npost, True, 8)
    return torch.sum(out, dim=1)


def step():
    h = x
    for w1, w2, gating in layers:
        h = block(w1, w2, gating, h)
    return h


s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(2):
        step()
    s.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr, stream=s):
        step()
torch.cuda.synchronize()
for _ in range(5):
    gr.replay()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
n = 30
for _ in range(n):
    gr.replay()
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / n / NL
active_bytes = topk * ((H * 2 * I) // 2 + (I * H) // 2) * (1 + 1 / 32 + 1 / 128)   # weights + scales + zeros
print(json.dumps({"us_per_moe_block": round(us, 2), "active_expert_MB": round(active_bytes / 1e6, 1),
                  "GBps_over_active_experts": round(active_bytes / us / 1e3, 1), "layers_rotated": NL,
                  "shape": "Mixtral-8x7B: E=8 top-2, hidden 4096, inter 14336, g128, bs=1"}))

"""Where does the stream program's time go?  Runs the 32-layer Llama-3-8B program under the kernel's experiment
modes (knob 3): 0 normal, 5 no unit math (pure weight stream + hand-off), 6 no waiting on producers' tags (no
hand-off), 7 both (pure stream)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from autoawq_b200 import ext  # noqa: E402
from autoawq_b200.program import DecodeProgram  # noqa: E402

dev = torch.device("cuda:0")
layers = int(sys.argv[1]) if len(sys.argv) > 1 else 32
rep = bench.Replica(dev, 1, layers=layers)
prog = DecodeProgram()
rep.step(rep.h, api=prog)
prog.build()
print("kind", prog.kind)
alg = sum(bench.linear_bytes(K, N, 1) for _, K, N in bench.LINEARS) * layers
sweep = [(0, 48, 0, 8), (0, 48, 0, 12), (0, 48, 0, 16), (0, -1, 0, 12), (0, -1, 0, 16), (0, 48, 1, 12), (0, 48, 2, 12),
         (6, 48, 0, 12), (7, 48, 0, 12), (6, 48, 0, 16), (7, 48, 0, 16)]
for mode, k8, k10, k9 in sweep:
    ext.set_knob(3, mode)
    ext.set_knob(8, k8)
    ext.set_knob(10, k10)
    ext.set_knob(9, k9)
    for _ in range(3):
        prog.run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        prog.run()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    print(f"mode {mode} l2_ahead {k8:4d} KB/lane gate {k10:2d} warps {k9:2d}: {us:8.1f} us per step, {alg / us / 1e3:7.1f} GB/s, abort {DecodeProgram.abort_record()}")
    ext.set_knob(3, 0)

"""Small-M kernel under programmatic dependent launch (knob 4, what bench.py runs with) inside a CUDA graph: 6 back-to-back
launches on different weights and token counts, each checked against the fp64 oracle; the split-K workspace must be
all-zero afterwards.  Prints PASS / FAIL lines; exit code 1 on failure."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autoawq_b200 import ext  # noqa: E402
from oracle import awq_oracle as O  # noqa: E402

dev = torch.device("cuda:0")
ok = True
for pdl in (0, 1):
    ext.set_knob(4, pdl)
    cases = []
    for i, (K, N, M) in enumerate([(1024, 1792, 16), (2048, 640, 64), (1024, 1792, 8), (512, 256, 100), (4096, 512, 33),
                                   (1024, 1792, 128)]):
        c = O.make_case(K, N, 128, seed=100 + i)
        x = np.random.default_rng(i).standard_normal((M, K)).astype(np.float16)
        w = O.dequantize_gemm(c["qweight"], c["qzeros"], c["scales"], 128)
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
        cases.append((t(x), t(c["qweight"]), t(c["scales"]), t(c["qzeros"]), O.gemm_f64(x, w),
                      np.abs(x.astype(np.float64)) @ np.abs(w.astype(np.float64)),
                      torch.empty((M, N), dtype=torch.float16, device=dev)))
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for xt, qw, sc, qz, _, _, y in cases:            # warm-up: allocates the stream's workspace outside the capture
            ext.linear_forward("gemm", xt, qw, sc, qz, 128, out=y)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        for rep in range(3):
            for xt, qw, sc, qz, _, _, y in cases:
                ext.linear_forward("gemm", xt, qw, sc, qz, 128, out=y)
    for y in (c[-1] for c in cases):
        y.zero_()
    for _ in range(5):
        g.replay()
    torch.cuda.synchronize()
    for i, (xt, _, _, _, ref, bud, y) in enumerate(cases):
        got = y.cpu().numpy().astype(np.float64)
        tol = 2.0**-10 * np.abs(ref) + 2.0**-16 * bud + 1e-6
        bad = int((np.abs(got - ref) > tol).sum()) + int((~np.isfinite(got)).sum())
        print(f"pdl={pdl} case {i} M={xt.shape[0]}: {'PASS' if bad == 0 else 'FAIL'} ({bad} bad)")
        ok &= bad == 0
    dirty = sum(int(ws.view(torch.int32).ne(0).sum()) for ws in ext._WS.values())
    print(f"pdl={pdl} workspace non-zero words: {dirty}")
    ok &= dirty == 0
ext.set_knob(4, 0)
sys.exit(0 if ok else 1)

"""Debug helper: single-linear stream programs at assorted (K, N, G) against the oracle; prints where they differ."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from autoawq_b200 import ext  # noqa: E402
from autoawq_b200.program import DecodeProgram  # noqa: E402
from oracle import awq_oracle as O  # noqa: E402

dev = torch.device("cuda:0")
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
ext.set_knob(14, 2)
for K, N, G in [(512, 1024, 32), (2048, 2048, 32), (512, 1024, 128), (512, 1024, 64), (512, 4096, 32), (256, 1024, 32),
                (1024, 1024, 32)]:
    c = O.make_case(K, N, G, seed=K)
    sc = (c["scales"].astype(np.float32) / (6.1 * 0.0108 * np.sqrt(K))).astype(np.float16)
    w = O.dequantize_gemm(c["qweight"], c["qzeros"], sc, G)
    x = np.random.default_rng(3).standard_normal((1, K)).astype(np.float16)
    prog = DecodeProgram()
    y = prog.gemm_forward_cuda(t(x), t(c["qweight"]), t(sc), t(c["qzeros"]), 8)
    prog.build()
    prog.run()
    torch.cuda.synchronize()
    ref = O.gemm_f64(x, w)[0]
    got = y.float().cpu().numpy()[0].astype(np.float64)
    bad = np.abs(got - ref) > 2e-3 * np.abs(ref) + 2e-3
    sets = bad.reshape(-1, 16)
    print(f"K={K} N={N} G={G} kind={prog.kind}: {bad.sum()} / {N} wrong; sets wrong {sets.any(1).sum()} / {N // 16};"
          f" lo-wrong {sets[:, :8].sum()} hi-wrong {sets[:, 8:].sum()}; abort {DecodeProgram.abort_record()}")
    if bad.any():
        i = np.where(bad)[0][:6]
        print("   first wrong cols", i, "got", got[i], "ref", ref[i])
        ws = np.where(sets.any(1))[0]
        print("   wrong sets", ws[:24])

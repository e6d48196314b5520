"""Is the Mixtral leg's number independent of what ran before it?  Fresh process: the leg alone, then the small-batch leg
+ one M = 4096 call (grows the split-K workspace), then the leg again.  Each run validates its own output (finite, two
distinct experts routed in the last layer).  Prints one JSON."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from autoawq_b200 import ext  # noqa: E402

dev = torch.device("cuda:0")
out = {}
r = bench.mixtral_leg(torch, dev, 20)
out["alone"] = {k: r[k] for k in ("tok_s", "ms_per_step", "checked")}
out["small_batch"] = bench.small_batch_leg(torch, ext, dev, 10, bench.measured_peaks())
K, N, G = 4096, 28672, 128
qw = torch.randint(-2**31, 2**31 - 1, (K, N // 8), dtype=torch.int32, device=dev)
qz = torch.randint(-2**31, 2**31 - 1, (K // G, N // 8), dtype=torch.int32, device=dev)
sc = (torch.rand((K // G, N), device=dev) * 0.01 + 0.001).half()
ext.linear_forward("gemm", torch.randn((4096, K), device=dev, dtype=torch.float16), qw, sc, qz, G)
torch.cuda.synchronize()
del qw, qz, sc
r = bench.mixtral_leg(torch, dev, 20)
out["after_other_legs"] = {k: r[k] for k in ("tok_s", "ms_per_step", "checked")}
print(json.dumps(out, indent=1))

"""2-GPU check of the column -> row sharded MLP with the NCCL all-reduce (skipped with fewer than 2 GPUs).
Compares the tensor-parallel result on the B200 kernels with the unsharded oracle contraction."""
import os
import socket

import numpy as np
import pytest
import torch

from oracle import awq_oracle as O

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from autoawq_b200 import shard as S

        K, I, G = 1024, 2048, 128
        dev = torch.device("cuda", rank)

        def packed(c):
            return S.PackedGemm(torch.from_numpy(c["qweight"]).to(dev), torch.from_numpy(c["qzeros"]).to(dev),
                                torch.from_numpy(c["scales"]).to(dev))

        cg, cu, cd = O.make_case(K, I, G, seed=1), O.make_case(K, I, G, seed=2), O.make_case(I, K, G, seed=3)
        mlp = S.TensorParallelMLP(packed(cg), packed(cu), packed(cd), rank, world)
        for M in (1, 24):
            x = (np.random.default_rng(M).standard_normal((M, K)) * 0.5).astype(np.float16)
            y = mlp(torch.from_numpy(x).to(dev)).float().cpu().numpy()
            wg = O.dequantize_gemm(cg["qweight"], cg["qzeros"], cg["scales"], G)
            wu = O.dequantize_gemm(cu["qweight"], cu["qzeros"], cu["scales"], G)
            wd = O.dequantize_gemm(cd["qweight"], cd["qzeros"], cd["scales"], G)
            g = O.gemm_f64(x, wg).astype(np.float16).astype(np.float64)
            u = O.gemm_f64(x, wu).astype(np.float16).astype(np.float64)
            act = (g / (1 + np.exp(-g)) * u).astype(np.float16)
            ref = O.gemm_f64(act, wd)
            # two fp16 roundings upstream + fp16 partial sums reduced across ranks
            tol = 4e-3 * np.abs(ref) + 2e-3 * np.sqrt(np.mean(ref**2)) + 1e-4
            assert np.all(np.abs(y - ref) <= tol), f"rank {rank} M={M}: max err {np.abs(y - ref).max()}"
        ret[rank] = True
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_tensor_parallel_mlp_nccl():
    import torch.multiprocessing as mp

    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
        assert ret.get(0) and ret.get(1)

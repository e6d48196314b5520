"""world_size-2 gloo test (CPU) of the multi-GPU host logic: column / row / qkv sharding of packed tensors
plus the all-reduce that finishes a row-parallel linear.  Per-rank partial products are computed with the
CPU oracle (test infrastructure) - the point here is the slicing arithmetic and the collective plumbing;
the kernels themselves are covered by the -m gpu tests."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import awq_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, K, N1, G, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from autoawq_b200 import shard as S

        c1 = O.make_case(K, N1, G, seed=1)       # column-parallel producer  K -> N1
        c2 = O.make_case(N1, K, G, seed=2)       # row-parallel consumer     N1 -> K
        p1 = S.PackedGemm(torch.from_numpy(c1["qweight"]), torch.from_numpy(c1["qzeros"]), torch.from_numpy(c1["scales"]))
        bias = torch.from_numpy((np.arange(K) % 7).astype(np.float16))
        p2 = S.PackedGemm(torch.from_numpy(c2["qweight"]), torch.from_numpy(c2["qzeros"]), torch.from_numpy(c2["scales"]), bias)
        x = np.random.default_rng(0).standard_normal((3, K)).astype(np.float16)
        # unsharded truth
        w1 = O.dequantize_gemm(c1["qweight"], c1["qzeros"], c1["scales"], G)
        w2 = O.dequantize_gemm(c2["qweight"], c2["qzeros"], c2["scales"], G)
        mid = O.gemm_f64(x, w1).astype(np.float16)
        full = O.gemm_f64(mid, w2) + bias.numpy().astype(np.float64)
        # sharded: column slice -> local mid -> row slice -> partial -> all-reduce
        s1 = S.shard_columns(p1, rank, world, 128 if N1 % (128 * world) == 0 else 8)
        s2 = S.shard_rows(p2, rank, world)
        w1s = O.dequantize_gemm(s1.qweight.numpy(), s1.qzeros.numpy(), s1.scales.numpy(), G)
        mid_local = O.gemm_f64(x, w1s).astype(np.float16)
        k0 = s2.in_features * rank  # equal split in this test
        assert np.array_equal(mid_local, mid[:, k0 : k0 + s2.in_features])
        assert np.array_equal(S.x_slice_for_rows(torch.from_numpy(mid), N1, G, rank, world).numpy(), mid_local)
        w2s = O.dequantize_gemm(s2.qweight.numpy(), s2.qzeros.numpy(), s2.scales.numpy(), G)
        part = O.gemm_f64(mid_local, w2s)
        if s2.bias is not None:
            part = part + s2.bias.numpy().astype(np.float64)
        assert (s2.bias is not None) == (rank == 0)
        y = torch.from_numpy(part)
        S.all_reduce_sum(y)
        np.testing.assert_allclose(y.numpy(), full, rtol=1e-9, atol=1e-9)
        # fused-qkv head-group split: the local columns are whole heads of q, k and v
        H, HKV, D = 8, 2, 16
        cq = O.make_case(K, (H + 2 * HKV) * D, G, seed=3)
        pq = S.PackedGemm(torch.from_numpy(cq["qweight"]), torch.from_numpy(cq["qzeros"]), torch.from_numpy(cq["scales"]))
        sq = S.shard_qkv(pq, H, HKV, D, rank, world)
        wq = O.dequantize_gemm(cq["qweight"], cq["qzeros"], cq["scales"], G)
        wqs = O.dequantize_gemm(sq.qweight.numpy(), sq.qzeros.numpy(), sq.scales.numpy(), G)
        hq, hk = H // world * D, HKV // world * D
        exp = np.concatenate([wq[:, rank * hq : (rank + 1) * hq],
                              wq[:, H * D + rank * hk : H * D + (rank + 1) * hk],
                              wq[:, (H + HKV) * D + rank * hk : (H + HKV) * D + (rank + 1) * hk]], axis=1)
        assert np.array_equal(wqs.view(np.uint16), exp.view(np.uint16))
        ret[rank] = True
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("K,N1,G", [(256, 512, 64), (128, 256, 32)])
def test_column_row_pair_world2(K, N1, G):
    world = 2
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), K, N1, G, ret), nprocs=world, join=True)
        assert all(ret.get(r) for r in range(world))


def test_bounds_and_errors():
    from autoawq_b200 import shard as S

    assert S._bounds(4096, 3, 8, 128) == (1536, 2048)
    assert [S._bounds(14336, r, 8, 128) for r in (0, 7)] == [(0, 1792), (12544, 14336)]
    with pytest.raises(ValueError):
        S._bounds(100, 0, 2, 8)
    c = O.make_case(64, 64, 32, seed=0)
    p = S.PackedGemm(torch.from_numpy(c["qweight"]), torch.from_numpy(c["qzeros"]), torch.from_numpy(c["scales"]))
    assert p.group_size == 32 and p.in_features == 64 and p.out_features == 64
    with pytest.raises(ValueError):
        S.shard_columns(p, 0, 2, quantum=4)
    with pytest.raises(ValueError):
        S.shard_qkv(p, 3, 1, 16, 0, 2)

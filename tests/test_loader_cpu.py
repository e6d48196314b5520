"""Checkpoint loader (autoawq_b200/loader.py) on the CPU box: a synthetic two-layer AWQ checkpoint in
the reference's buffer naming, split over two safetensors files, loaded whole and as tensor-parallel shards; every
shard must equal the slice autoawq_b200/shard.py takes from the whole tensors, and the concatenation of the shards'
outputs must reproduce the unsharded linear (oracle contraction)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import awq_oracle as O

safetensors = pytest.importorskip("safetensors")


@pytest.fixture(scope="module")
def ckpt(tmp_path_factory):
    from safetensors.torch import save_file

    d = tmp_path_factory.mktemp("ckpt")
    G, H, I = 64, 256, 512
    shapes = {"self_attn.q_proj": (H, H), "self_attn.k_proj": (H, 64), "self_attn.v_proj": (H, 64),
              "self_attn.o_proj": (H, H), "mlp.gate_proj": (H, I), "mlp.up_proj": (H, I), "mlp.down_proj": (I, H)}
    full, weight_map = {}, {}
    for layer in range(2):
        tensors = {}
        for i, (name, (K, N)) in enumerate(shapes.items()):
            c = O.make_case(K, N, G, seed=layer * 20 + i)
            p = f"model.layers.{layer}.{name}"
            tensors[p + ".qweight"] = torch.from_numpy(c["qweight"])
            tensors[p + ".qzeros"] = torch.from_numpy(c["qzeros"])
            tensors[p + ".scales"] = torch.from_numpy(c["scales"])
            if name.endswith("o_proj") or name.endswith("q_proj"):
                tensors[p + ".bias"] = torch.from_numpy((np.arange(N) % 7).astype(np.float16))
        tensors[f"model.layers.{layer}.input_layernorm.weight"] = torch.ones(H, dtype=torch.float16)
        fn = f"model-0000{layer + 1}-of-00002.safetensors"
        save_file(tensors, str(d / fn))
        full.update(tensors)
        weight_map.update({k: fn for k in tensors})
    (d / "model.safetensors.index.json").write_text(json.dumps({"weight_map": weight_map}))
    return str(d), full, G


def test_index_and_whole_load(ckpt):
    from autoawq_b200.loader import CheckpointIndex, load_packed_linears, split_mode

    path, full, G = ckpt
    idx = CheckpointIndex(path)
    pre = idx.linear_prefixes()
    assert len(pre) == 14 and "model.layers.1.mlp.down_proj" in pre
    assert split_mode("model.layers.0.self_attn.q_proj") == "column" and split_mode("x.mlp.down_proj") == "row"
    assert split_mode("lm_head") == "replicate"
    lin = load_packed_linears(path, "cpu")
    for p, pk in lin.items():
        assert torch.equal(pk.qweight, full[p + ".qweight"]) and torch.equal(pk.scales, full[p + ".scales"])
        assert pk.group_size == G
        assert (pk.bias is not None) == (p + ".bias" in full)
    # without the index file: the directory is scanned
    os.rename(os.path.join(path, "model.safetensors.index.json"), os.path.join(path, "index.bak"))
    try:
        assert CheckpointIndex(path).linear_prefixes() == pre
    finally:
        os.rename(os.path.join(path, "index.bak"), os.path.join(path, "model.safetensors.index.json"))


@pytest.mark.parametrize("world", [2, 4])
def test_shards_match_shard_py_and_recompose(ckpt, world):
    from autoawq_b200 import shard
    from autoawq_b200.loader import CheckpointIndex, fuse_columns, load_packed_linear

    path, full, G = ckpt
    idx = CheckpointIndex(path)
    rng = np.random.default_rng(0)
    for p in ["model.layers.0.self_attn.q_proj", "model.layers.1.mlp.gate_proj", "model.layers.1.self_attn.o_proj",
              "model.layers.0.mlp.down_proj"]:
        whole = shard.PackedGemm(full[p + ".qweight"], full[p + ".qzeros"], full[p + ".scales"], full.get(p + ".bias"))
        K, N = whole.in_features, whole.out_features
        w = O.dequantize_gemm(whole.qweight.numpy(), whole.qzeros.numpy(), whole.scales.numpy(), G)
        x = rng.standard_normal((3, K)).astype(np.float16)
        ref = O.gemm_f64(x, w) + (whole.bias.numpy().astype(np.float64) if whole.bias is not None else 0.0)
        mode = "row" if p.endswith(("o_proj", "down_proj")) else "column"
        parts = []
        for r in range(world):
            got = load_packed_linear(idx, p, "cpu", r, world)
            exp = shard.shard_rows(whole, r, world) if mode == "row" else shard.shard_columns(whole, r, world)
            for a, b in [(got.qweight, exp.qweight), (got.qzeros, exp.qzeros), (got.scales, exp.scales)]:
                assert torch.equal(a, b), (p, r)
            assert (got.bias is None) == (exp.bias is None)
            if got.bias is not None:
                assert torch.equal(got.bias, exp.bias)
            ws = O.dequantize_gemm(got.qweight.numpy(), got.qzeros.numpy(), got.scales.numpy(), G)
            if mode == "row":
                k0, k1 = shard._bounds(K, r, world, G)
                y = O.gemm_f64(x[:, k0:k1], ws)
                if got.bias is not None:
                    y = y + got.bias.numpy().astype(np.float64)
            else:
                y = O.gemm_f64(x, ws) + (got.bias.numpy().astype(np.float64) if got.bias is not None else 0.0)
            parts.append(y)
        out = sum(parts) if mode == "row" else np.concatenate(parts, axis=1)
        np.testing.assert_allclose(out, ref, rtol=1e-12, atol=1e-12)
    # fused q|k|v of one rank = the three column shards side by side
    r, world2 = 1, 2
    qkv = fuse_columns(load_packed_linear(idx, f"model.layers.0.self_attn.{n}_proj", "cpu", r, world2) for n in "qkv")
    assert qkv.out_features == (256 + 64 + 64) // world2 and qkv.bias is not None and qkv.bias.shape[0] == qkv.out_features

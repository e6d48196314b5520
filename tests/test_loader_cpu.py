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


def test_uneven_groups_column_and_row_shards_line_up(tmp_path):
    """I = 11008, G = 128, tp = 4 -> 86 groups do not divide by 4.  gate / up columns must be split on
    the boundaries of down's K split, or a rank's activation width differs from its down rows."""
    from safetensors.torch import save_file

    from autoawq_b200 import shard
    from autoawq_b200.loader import check_tp_alignment, load_packed_linears

    H, I, G, world = 128, 11008, 128, 4
    tensors = {}
    for i, (name, (K, N)) in enumerate({"mlp.gate_proj": (H, I), "mlp.up_proj": (H, I), "mlp.down_proj": (I, H)}.items()):
        rng = np.random.default_rng(i)
        p = f"model.layers.0.{name}"
        tensors[p + ".qweight"] = torch.from_numpy(rng.integers(-2**31, 2**31 - 1, (K, N // 8), dtype=np.int64).astype(np.int32))
        tensors[p + ".qzeros"] = torch.from_numpy(rng.integers(-2**31, 2**31 - 1, (K // G, N // 8), dtype=np.int64).astype(np.int32))
        tensors[p + ".scales"] = torch.from_numpy(rng.random((K // G, N)).astype(np.float16))
    save_file(tensors, str(tmp_path / "model.safetensors"))
    widths = []
    for r in range(world):
        sh = load_packed_linears(str(tmp_path), "cpu", r, world)     # check_tp_alignment runs inside
        g, u, d = (sh[f"model.layers.0.mlp.{n}_proj"] for n in ("gate", "up", "down"))
        assert g.out_features == u.out_features == d.in_features
        k0, k1 = shard._bounds(I, r, world, G)
        assert torch.equal(g.qweight, tensors["model.layers.0.mlp.gate_proj.qweight"][:, k0 // 8:k1 // 8])
        assert torch.equal(d.qweight, tensors["model.layers.0.mlp.down_proj.qweight"][k0:k1])
        widths.append(d.in_features)
    assert sum(widths) == I and len(set(widths)) == 2           # 21 / 22 groups: genuinely uneven
    # the old behaviour (8-column quantum for gate / up) is caught, not silently mis-paired
    bad = load_packed_linears(str(tmp_path), "cpu", 1, world, column_quantum=8,
                              prefixes=["model.layers.0.mlp.gate_proj", "model.layers.0.mlp.up_proj"])
    bad["model.layers.0.mlp.down_proj"] = load_packed_linears(
        str(tmp_path), "cpu", 1, world, prefixes=["model.layers.0.mlp.down_proj"])["model.layers.0.mlp.down_proj"]
    with pytest.raises(ValueError):
        check_tp_alignment(bad, 1, world)
    # TensorParallelMLP uses the same boundaries (and carries gate / up biases)
    whole = {n: shard.PackedGemm(tensors[f"model.layers.0.mlp.{n}_proj.qweight"], tensors[f"model.layers.0.mlp.{n}_proj.qzeros"],
                                 tensors[f"model.layers.0.mlp.{n}_proj.scales"]) for n in ("gate", "up", "down")}
    whole["gate"].bias = torch.arange(I, dtype=torch.float16)
    try:
        mlp = shard.TensorParallelMLP(whole["gate"], whole["up"], whole["down"], 1, world)
    except ImportError:
        pytest.skip("libb200awq.so not built")
    assert mlp.gu.out_features == 2 * mlp.down.in_features
    k0, k1 = shard._bounds(I, 1, world, G)
    assert torch.equal(mlp.gu.bias[: k1 - k0], whole["gate"].bias[k0:k1]) and float(mlp.gu.bias[k1 - k0:].abs().sum()) == 0


def test_fused_names_are_split_per_section(tmp_path):
    """qkv_proj / gate_up_proj (Phi-3-style fused checkpoints) must be split per section, not as one
    contiguous N range."""
    from safetensors.torch import save_file

    from autoawq_b200 import shard
    from autoawq_b200.loader import CheckpointIndex, load_packed_linear

    H, I, G, heads, kv, d = 256, 512, 64, 4, 2, 64
    tensors = {}
    for i, (name, (K, N)) in enumerate({"self_attn.qkv_proj": (H, (heads + 2 * kv) * d), "mlp.gate_up_proj": (H, 2 * I)}.items()):
        c = O.make_case(K, N, G, seed=i)
        p = f"model.layers.0.{name}"
        tensors[p + ".qweight"], tensors[p + ".qzeros"], tensors[p + ".scales"] = (torch.from_numpy(c[k]) for k in
                                                                                    ("qweight", "qzeros", "scales"))
    save_file(tensors, str(tmp_path / "model.safetensors"))
    idx = CheckpointIndex(str(tmp_path))
    world = 2
    for r in range(world):
        gu = load_packed_linear(idx, "model.layers.0.mlp.gate_up_proj", "cpu", r, world)
        whole = shard.PackedGemm(*(tensors[f"model.layers.0.mlp.gate_up_proj.{k}"] for k in ("qweight", "qzeros", "scales")))
        half = I // world
        assert gu.out_features == 2 * half
        assert torch.equal(gu.scales[:, :half], whole.scales[:, r * half:(r + 1) * half])              # gate share
        assert torch.equal(gu.scales[:, half:], whole.scales[:, I + r * half:I + (r + 1) * half])      # up share
        qkv = load_packed_linear(idx, "model.layers.0.self_attn.qkv_proj", "cpu", r, world, qkv_heads=(heads, kv, d))
        wq = shard.PackedGemm(*(tensors[f"model.layers.0.self_attn.qkv_proj.{k}"] for k in ("qweight", "qzeros", "scales")))
        exp = shard.shard_qkv(wq, heads, kv, d, r, world)
        assert torch.equal(qkv.qweight, exp.qweight) and torch.equal(qkv.scales, exp.scales)
    with pytest.raises(ValueError):
        load_packed_linear(idx, "model.layers.0.self_attn.qkv_proj", "cpu", 0, world)   # heads unknown: refuse

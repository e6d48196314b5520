"""Checkpoint -> (sharded) packed device buffers, without accelerate.

The reference loads an AWQ checkpoint by building the whole module tree on the meta device and letting accelerate's
`load_checkpoint_and_dispatch` place every tensor (awq/models/base.py:505-535); multi-GPU there means layer placement,
every rank would read every tensor.  For the tensor-parallel decode of BASELINE config 5 each rank needs only its
column / row slice of every packed tensor (autoawq_b200/shard.py): this loader reads exactly those bytes.  safetensors
files are memory-mapped and sliceable, so a column shard `qweight[:, a:b]` or a row shard `qweight[k0:k1]` is read
straight from the file into the rank's device buffer.

Tensor names follow the reference's module buffers (awq/modules/linear/gemm.py:135-170): `<prefix>.qweight`,
`<prefix>.qzeros`, `<prefix>.scales`, optional `<prefix>.bias`, GEMM layout (the checkpoint default,
awq/models/_config.py:14).
"""
from __future__ import annotations

import json
import os
from typing import Dict, Iterable, Optional

import torch

from .shard import PackedGemm, _bounds

# which way a linear is split, by the last component of its name (Llama / Mistral / Mixtral / Qwen naming)
COLUMN_PARALLEL = ("q_proj", "k_proj", "v_proj", "gate_proj", "up_proj", "w1", "w3", "qkv_proj", "gate_up_proj")
ROW_PARALLEL = ("o_proj", "down_proj", "w2", "dense", "out_proj")


class CheckpointIndex:
    """name -> file for a directory of *.safetensors shards (uses model.safetensors.index.json when present)."""

    def __init__(self, path: str):
        self.files: Dict[str, str] = {}
        if os.path.isfile(path):
            self._scan(path)
            return
        idx = os.path.join(path, "model.safetensors.index.json")
        if os.path.exists(idx):
            with open(idx) as f:
                wm = json.load(f)["weight_map"]
            self.files = {k: os.path.join(path, v) for k, v in wm.items()}
        else:
            for fn in sorted(os.listdir(path)):
                if fn.endswith(".safetensors"):
                    self._scan(os.path.join(path, fn))
        if not self.files:
            raise FileNotFoundError(f"no safetensors tensors under {path}")

    def _scan(self, file: str):
        from safetensors import safe_open

        with safe_open(file, framework="pt", device="cpu") as f:
            for k in f.keys():
                self.files[k] = file

    def linear_prefixes(self) -> list:
        """Prefixes that hold a complete packed linear (qweight + qzeros + scales)."""
        out = []
        for k in self.files:
            if k.endswith(".qweight"):
                p = k[: -len(".qweight")]
                if p + ".qzeros" in self.files and p + ".scales" in self.files:
                    out.append(p)
        return sorted(out)


def split_mode(prefix: str, column: Iterable[str] = COLUMN_PARALLEL, row: Iterable[str] = ROW_PARALLEL) -> str:
    leaf = prefix.rsplit(".", 1)[-1]
    if leaf in column:
        return "column"
    if leaf in row:
        return "row"
    return "replicate"


def _read(index: CheckpointIndex, name: str, device, rows: Optional[slice] = None, cols: Optional[slice] = None):
    from safetensors import safe_open

    with safe_open(index.files[name], framework="pt", device="cpu") as f:
        sl = f.get_slice(name)
        nd = len(sl.get_shape())
        if nd == 1:
            t = sl[cols] if cols is not None else sl[:]
        else:
            t = sl[rows if rows is not None else slice(None), cols if cols is not None else slice(None)]
    return t.contiguous().to(device, non_blocking=False)


def load_packed_linear(index: CheckpointIndex, prefix: str, device, tp_rank: int = 0, tp_world: int = 1,
                       mode: Optional[str] = None, column_quantum: int = 8) -> PackedGemm:
    """One packed linear, or this rank's shard of it.  mode: "column" (split N on `column_quantum`-column boundaries),
    "row" (split K on group boundaries; the bias stays on rank 0 so the all-reduce adds it once), "replicate",
    or None = by name (`split_mode`)."""
    from safetensors import safe_open

    mode = mode or split_mode(prefix)
    if tp_world == 1:
        mode = "replicate"
    with safe_open(index.files[prefix + ".qweight"], framework="pt", device="cpu") as f:
        K, NW = f.get_slice(prefix + ".qweight").get_shape()
    with safe_open(index.files[prefix + ".scales"], framework="pt", device="cpu") as f:
        KG, N = f.get_slice(prefix + ".scales").get_shape()
    if N != NW * 8 or K % KG != 0:
        raise ValueError(f"{prefix}: qweight {K}x{NW} and scales {KG}x{N} do not form a GEMM-layout linear")
    G = K // KG
    has_bias = prefix + ".bias" in index.files
    if mode == "column":
        if column_quantum % 8 != 0:
            raise ValueError("column quantum must be a multiple of 8 (one packed word)")
        n0, n1 = _bounds(N, tp_rank, tp_world, column_quantum)
        cw = slice(n0 // 8, n1 // 8)
        return PackedGemm(_read(index, prefix + ".qweight", device, cols=cw),
                          _read(index, prefix + ".qzeros", device, cols=cw),
                          _read(index, prefix + ".scales", device, cols=slice(n0, n1)),
                          _read(index, prefix + ".bias", device, cols=slice(n0, n1)) if has_bias else None)
    if mode == "row":
        k0, k1 = _bounds(K, tp_rank, tp_world, G)
        return PackedGemm(_read(index, prefix + ".qweight", device, rows=slice(k0, k1)),
                          _read(index, prefix + ".qzeros", device, rows=slice(k0 // G, k1 // G)),
                          _read(index, prefix + ".scales", device, rows=slice(k0 // G, k1 // G)),
                          _read(index, prefix + ".bias", device) if (has_bias and tp_rank == 0) else None)
    if mode != "replicate":
        raise ValueError(mode)
    return PackedGemm(_read(index, prefix + ".qweight", device), _read(index, prefix + ".qzeros", device),
                      _read(index, prefix + ".scales", device),
                      _read(index, prefix + ".bias", device) if has_bias else None)


def load_packed_linears(path: str, device, tp_rank: int = 0, tp_world: int = 1, prefixes: Optional[Iterable[str]] = None,
                        column_quantum: int = 8) -> Dict[str, PackedGemm]:
    """Every packed linear of a checkpoint directory (or the given prefixes), sharded for (tp_rank, tp_world)."""
    index = CheckpointIndex(path)
    out = {}
    for p in (list(prefixes) if prefixes is not None else index.linear_prefixes()):
        out[p] = load_packed_linear(index, p, device, tp_rank, tp_world, None, column_quantum)
    return out


def fuse_columns(parts: Iterable[PackedGemm]) -> PackedGemm:
    """Concatenate column-parallel linears that share their input (q|k|v, gate|up) along N - what the reference's
    fuse_qkv does for whole tensors (awq/utils/fused_utils.py:67-96), here for the rank's shards."""
    parts = list(parts)
    G = parts[0].group_size
    if any(p.group_size != G or p.in_features != parts[0].in_features for p in parts):
        raise ValueError("fused linears must share in_features and group size")
    bias = None
    if any(p.bias is not None for p in parts):
        bias = torch.cat([p.bias if p.bias is not None else
                          torch.zeros(p.out_features, dtype=torch.float16, device=p.qweight.device) for p in parts])
    return PackedGemm(torch.cat([p.qweight for p in parts], dim=1).contiguous(),
                      torch.cat([p.qzeros for p in parts], dim=1).contiguous(),
                      torch.cat([p.scales for p in parts], dim=1).contiguous(), bias)

"""Host-side producers of the three packed formats (torch, any device), vectorised.

These mirror what the reference's `from_linear` class methods emit (awq/modules/linear/gemm.py:171-251,
gemv.py:78-153, gemv_fast.py:26-65,120-181) so a model quantised by the reference's AwqQuantizer can be
packed without the O(K) Python loops.  Verified bit-for-bit against the reference's own outputs
(tests/golden/packers.npz) in tests/test_packing.py.
"""
from __future__ import annotations

import torch

PACK_NUM = 8
# nibble i of a GEMM-layout word <- column 8c + AWQ_ORDER[i]  (awq/utils/packing_utils.py:4)
AWQ_ORDER = (0, 2, 4, 6, 1, 3, 5, 7)
AWQ_REVERSE_ORDER = (0, 4, 1, 5, 2, 6, 3, 7)


def calculate_zeros_width(in_features: int, group_size: int = 128, pack_num: int = 8) -> int:
    """Padded width (int32 words) of a GEMV-layout zeros row (gemv.py:12-24)."""
    if group_size >= 128:
        mult = 1
    elif group_size == 64:
        mult = 2
    elif group_size == 32:
        mult = 4
    else:
        raise NotImplementedError
    base = (in_features // group_size + pack_num - 1) // pack_num
    return (base + mult - 1) // mult * mult


def _pack_words(vals: torch.Tensor, nibble_of_col) -> torch.Tensor:
    """[R, 8C] ints in 0..15 -> [R, C] int32; column j of each octet goes to nibble nibble_of_col[j]."""
    R, C8 = vals.shape
    v = vals.to(torch.int64).reshape(R, C8 // PACK_NUM, PACK_NUM) & 0xF
    shifts = torch.tensor([4 * nibble_of_col[j] for j in range(PACK_NUM)], dtype=torch.int64, device=vals.device)
    words = (v << shifts).sum(dim=-1)  # disjoint bit fields: sum == or
    words = torch.where(words >= 2**31, words - 2**32, words)
    return words.to(torch.int32)


def pack_gemm_words(vals: torch.Tensor) -> torch.Tensor:
    return _pack_words(vals, AWQ_REVERSE_ORDER)


def pack_seq_words(vals: torch.Tensor) -> torch.Tensor:
    return _pack_words(vals, tuple(range(PACK_NUM)))


def unpack_gemm_words(words: torch.Tensor) -> torch.Tensor:
    """[R, C] int32 (GEMM interleave) -> [R, 8C] uint8 in natural column order."""
    shifts = torch.tensor([4 * r for r in AWQ_REVERSE_ORDER], dtype=torch.int32, device=words.device)
    v = (words.unsqueeze(-1) >> shifts) & 0xF
    return v.reshape(words.shape[0], -1).to(torch.uint8)


def quantize_to_int(weight_nk: torch.Tensor, scales_ng: torch.Tensor, zeros_ng: torch.Tensor, group_size: int):
    """round((W + z*s) / s) per group, the integer recovery `from_linear` performs (gemm.py:196-203):
    W [N, K] (already pseudo-quantised), scales / zeros [N, K/G] -> ints [N, K] (no clamp, as the reference)."""
    N, K = weight_nk.shape
    s = scales_ng.to(torch.float16).repeat_interleave(group_size, dim=1)  # fp16 scales, as stored
    sz = (zeros_ng * scales_ng).repeat_interleave(group_size, dim=1)
    return torch.round((weight_nk + sz) / s).to(torch.int32)


def pack_gemm(intweight_nk: torch.Tensor, zeros_ng: torch.Tensor, scales_ng: torch.Tensor):
    """-> qweight [K, N/8] i32, qzeros [K/G, N/8] i32, scales [K/G, N] f16 (checkpoint default `version="gemm"`)."""
    qweight = pack_gemm_words(intweight_nk.t().contiguous())
    qzeros = pack_gemm_words(zeros_ng.t().contiguous().to(torch.int32))
    return qweight, qzeros, scales_ng.t().contiguous().to(torch.float16)


def pack_gemv(intweight_nk: torch.Tensor, zeros_ng: torch.Tensor, scales_ng: torch.Tensor, group_size: int):
    """-> qweight [N, K/8] i32 (sequential nibbles), qzeros [N, zw] i32, scales [N, 8 zw] f16 (zero padded)."""
    N, K = intweight_nk.shape
    zw = calculate_zeros_width(K, group_size)
    ng = K // group_size
    qweight = pack_seq_words(intweight_nk)
    z = torch.zeros((N, zw * PACK_NUM), dtype=torch.int32, device=intweight_nk.device)
    z[:, :ng] = zeros_ng.to(torch.int32)
    s = torch.zeros((N, zw * PACK_NUM), dtype=torch.float16, device=intweight_nk.device)
    s[:, :ng] = scales_ng.to(torch.float16)
    return qweight, pack_seq_words(z), s


def _fast_kperm(device) -> torch.Tensor:
    a = torch.arange(32, device=device).reshape(4, 4, 2).permute(1, 0, 2).reshape(32)
    return a.reshape(4, 4, 2).permute(0, 2, 1).reshape(32)


def pack_gemv_fast_weight(intweight_nk: torch.Tensor, interleave: int = 4, kstride: int = 64) -> torch.Tensor:
    """ints [N, K] -> int16 [N/4, K] (gemv_fast.py:26-65)."""
    N, K = intweight_nk.shape
    w = intweight_nk.to(torch.int32).reshape(N, K // 32, 32)[:, :, _fast_kperm(intweight_nk.device)].reshape(N, K)
    w = w.reshape(N // interleave, interleave, K // kstride, kstride).permute(0, 2, 1, 3).contiguous()
    w = w.reshape(N // interleave, K // kstride, kstride, interleave)
    packed = w[..., 0] | (w[..., 1] << 4) | (w[..., 2] << 8) | (w[..., 3] << 12)
    packed = torch.where(packed >= 2**15, packed - 2**16, packed)
    return packed.reshape(N // interleave, K).to(torch.int16)


def pack_gemv_fast(intweight_nk, zeros_ng, scales_ng, group_size: int):
    """-> qweight i16 [N/4, K], scales f16 [8 zw, N], scaled zeros f16 [8 zw, N] = -(s * z) (gemv_fast.py:146-181)."""
    N, K = intweight_nk.shape
    zw = calculate_zeros_width(K, group_size)
    ng = K // group_size
    qs = torch.zeros((N, zw * PACK_NUM), dtype=torch.float16, device=intweight_nk.device)
    qs[:, :ng] = scales_ng.to(torch.float16)
    qz = torch.zeros_like(qs)
    qz[:, :ng] = -(qs[:, :ng] * zeros_ng.to(torch.float32)).to(torch.float16)
    return pack_gemv_fast_weight(intweight_nk), qs.t().contiguous(), qz.t().contiguous()

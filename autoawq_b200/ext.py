"""Torch-facing operator layer: the functions the reference imports from `awq_ext` / `awq_v2_ext`
(call sites: awq/modules/linear/gemm.py:51-58, gemv.py:168-180, gemv_fast.py:192-205,
awq/modules/fused/norm.py:33-36, moe.py:76), implemented on the C ABI of libb200awq.so.

torch is plumbing only: device memory (caching allocator), the current stream, dtype/shape checks.
Every function launches on torch's current stream, never synchronises and is CUDA-graph capturable
once the per-(device, stream) workspace exists (first call on that stream allocates it).
"""
from __future__ import annotations

import torch

from . import _cabi
from ._cabi import B200AwqError, check, lib

__all__ = [
    "gemm_forward_cuda", "dequantize_weights_cuda", "gemv_forward_cuda", "gemmv2_forward_cuda",
    "gemv_forward_cuda_decode", "gemm_forward_cuda_prefill", "layernorm_forward_cuda", "silu_and_mul",
    "topk_softmax", "moe_alig_block_size", "grouped_gemm_forward",
    "linear_forward", "stream_pack", "set_knob", "get_knob", "B200AwqError",
]

_WS: dict = {}
_WS_MIN = 16 << 20
_WS_TICKETS = 16384          # kTicketBytes (csrc/kernels.h); tests/test_host_cpu.py checks it against the ABI


def _require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise B200AwqError("b200awq: tensors must live on a CUDA device (there is no CPU path)")


def _workspace(dev: torch.device, stream_ptr: int, need: int) -> torch.Tensor:
    key = (dev.index, stream_ptr)
    ws = _WS.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.zeros(max(need, _WS_MIN), dtype=torch.uint8, device=dev)
        _WS[key] = ws
    return ws


def _stream(dev: torch.device) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


class _DeviceGuard:
    """Launch on the tensor's device even when it is not the current one (accelerate places layers
    on several GPUs; the reference's Triton path does the same guard, awq/modules/triton/gemm.py:23-27)."""

    __slots__ = ("dev", "prev")

    def __init__(self, dev: torch.device):
        self.dev, self.prev = dev.index, None

    def __enter__(self):
        cur = torch.cuda.current_device()
        if cur != self.dev:
            self.prev = cur
            torch.cuda.set_device(self.dev)

    def __exit__(self, *exc):
        if self.prev is not None:
            torch.cuda.set_device(self.prev)


def _x2d(x: torch.Tensor, K: int) -> torch.Tensor:
    if x.dtype != torch.float16:
        raise B200AwqError(f"b200awq: activations must be float16, got {x.dtype}")
    if x.shape[-1] != K:
        raise B200AwqError(f"b200awq: activation feature dim {x.shape[-1]} != in_features {K}")
    x2 = x.reshape(-1, K)
    if x2.stride(-1) != 1 or (x2.shape[0] > 1 and (x2.stride(0) < K or x2.stride(0) % 8 != 0)) \
            or x2.data_ptr() % 16 != 0:
        x2 = x2.contiguous()  # TMA needs 16-byte aligned rows
    return x2


def _check_w(t: torch.Tensor, dtype, name: str):
    if t.dtype != dtype or not t.is_contiguous():
        raise B200AwqError(f"b200awq: {name} must be contiguous {dtype}, got {t.dtype} contiguous={t.is_contiguous()}")


_LAYOUT = {
    "gemm": (lib.b200awq_gemm_forward, torch.int32),
    "gemv": (lib.b200awq_gemv_forward, torch.int32),
    "fast": (lib.b200awq_fast_forward, torch.int16),
}
# validated weight triples: id(qweight) -> (weakref to qweight, data_ptrs, K, N, device).  A decode loop calls the
# same 160 linears every token: dtype / contiguity / device checks and the pointer reads are done once per tensor
# (the weakref guards against id() reuse after the tensor died; in-place updates keep the pointers valid).
_WCACHE: dict = {}
_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_cur_device = getattr(torch._C, "_cuda_getDevice", None)


def _weights(layout: str, qweight, scales, qzeros):
    import weakref

    key = (id(qweight), id(scales), id(qzeros))
    hit = _WCACHE.get(key)
    if hit is not None:
        refs = hit[7]
        # ids are only unique among LIVE objects: every one of the three must still be the tensor that was validated
        if refs[0]() is qweight and refs[1]() is scales and refs[2]() is qzeros and hit[1] == qweight.data_ptr():
            return hit
    _require_cuda(qweight, scales, qzeros)
    wdt = _LAYOUT[layout][1]
    if layout == "gemm":
        K, N = qweight.shape[0], qweight.shape[1] * 8
    elif layout == "gemv":
        N, K = qweight.shape[0], qweight.shape[1] * 8
    else:
        N, K = qweight.shape[0] * 4, qweight.shape[1]
    _check_w(qweight, wdt, "qweight")
    _check_w(scales, torch.float16, "scales")
    _check_w(qzeros, torch.float16 if layout == "fast" else torch.int32, "qzeros")
    if scales.device != qweight.device or qzeros.device != qweight.device:
        raise B200AwqError("b200awq: qweight / scales / qzeros must live on one device")
    refs = (weakref.ref(qweight), weakref.ref(scales), weakref.ref(qzeros))
    ent = (refs[0], qweight.data_ptr(), scales.data_ptr(), qzeros.data_ptr(), K, N, qweight.device, refs)
    if len(_WCACHE) > 16384:
        _WCACHE.clear()
    _WCACHE[key] = ent
    return ent


def linear_forward(layout: str, x, qweight, scales, qzeros, group_size: int, bias=None, out=None) -> torch.Tensor:
    """Y = X . deq(W) (+ bias) for layout in {"gemm", "gemv", "fast"}; returns [M, N] fp16 (written into `out`
    when given: a contiguous [M, N] fp16 tensor).  The hot call of an eager decode loop (160 per token): weights are
    validated once per tensor, no context-manager object, no extra ABI call (b200awq_workspace_bytes is
    16384 + min(M, 128) * N * 8, restated here)."""
    try:
        fn = _LAYOUT[layout][0]
    except KeyError:
        raise ValueError(layout) from None
    _, p_qw, p_sc, p_qz, K, N, dev, _refs = _weights(layout, qweight, scales, qzeros)
    if not x.is_cuda or (bias is not None and not bias.is_cuda):
        raise B200AwqError("b200awq: tensors must live on a CUDA device (there is no CPU path)")
    x2 = _x2d(x, K)
    M = x2.shape[0]
    if x2.device != dev:
        raise B200AwqError(f"b200awq: activations on {x2.device}, weights on {dev}")
    if out is None:
        y = torch.empty((M, N), dtype=torch.float16, device=dev)
    else:
        y = out
        if y.dtype != torch.float16 or not y.is_contiguous() or y.numel() != M * N or y.device != dev:
            raise B200AwqError("b200awq: `out` must be a contiguous float16 [M, N] tensor on the input's device")
    if M == 0:
        return y
    G = K if group_size in (-1, 0) else int(group_size)
    di = dev.index
    cur = _cur_device() if _cur_device is not None else torch.cuda.current_device()
    if cur != di:
        torch.cuda.set_device(di)
    try:
        st = _raw_stream(di) if _raw_stream is not None else torch.cuda.current_stream(dev).cuda_stream
        need = _WS_TICKETS + (M if M < 128 else 128) * N * 8
        ws = _WS.get((di, st))
        if ws is None or ws.numel() < need:
            ws = _workspace(dev, st, need)
        code = fn(x2.data_ptr(), x2.stride(0) if M > 1 else K, p_qw, p_sc, p_qz,
                  bias.data_ptr() if bias is not None else None, y.data_ptr(), M, K, N, G,
                  ws.data_ptr(), ws.numel(), st)
    finally:
        if cur != di:
            torch.cuda.set_device(cur)
    if code != 0:
        check(code, f"b200awq_{layout}_forward(M={M}, K={K}, N={N}, G={G})")
    return y


# ----------------------------------------------------------------------------- awq_ext surface
def gemm_forward_cuda(x, qweight, scales, qzeros, split_k_iters=8):
    """awq_ext.gemm_forward_cuda (gemm.py:56-58): x [M, K] f16, GEMM layout -> [M, N] f16.
    `split_k_iters` is a legacy hint of the reference kernels; accepted and ignored."""
    K = qweight.shape[0]
    G = K // scales.shape[0]
    out = linear_forward("gemm", x, qweight, scales, qzeros, G)
    return out.reshape(x.shape[:-1] + (out.shape[-1],))


def dequantize_weights_cuda(qweight, scales, qzeros, split_k_iters=0, thx=0, thy=0, dbg=False):
    """awq_ext.dequantize_weights_cuda (gemm.py:51-53, tests/test_dequantization.py:41-49) -> [K, N] f16."""
    _require_cuda(qweight, scales, qzeros)
    _check_w(qweight, torch.int32, "qweight")
    _check_w(scales, torch.float16, "scales")
    _check_w(qzeros, torch.int32, "qzeros")
    K, N = qweight.shape[0], qweight.shape[1] * 8
    G = K // scales.shape[0]
    dev = qweight.device
    out = torch.empty((K, N), dtype=torch.float16, device=dev)
    with _DeviceGuard(dev):
        code = lib.b200awq_dequantize_gemm(qweight.data_ptr(), scales.data_ptr(), qzeros.data_ptr(), out.data_ptr(),
                                           K, N, G, _stream(dev))
    check(code, f"b200awq_dequantize_gemm(K={K}, N={N}, G={G})")
    return out


def gemv_forward_cuda(x, qweight, scales, qzeros, group_size):
    """awq_ext.gemv_forward_cuda (gemv.py:177-180): GEMV layout, M <= 8."""
    out = linear_forward("gemv", x, qweight, scales, qzeros, group_size)
    return out.reshape(x.shape[:-1] + (out.shape[-1],))


def gemmv2_forward_cuda(x, qweight, scales, qzeros, group_size, split_k_iters=8):
    """awq_ext.gemmv2_forward_cuda (gemv.py:168-176): GEMV layout, M > 8."""
    out = linear_forward("gemv", x, qweight, scales, qzeros, group_size)
    return out.reshape(x.shape[:-1] + (out.shape[-1],))


def layernorm_forward_cuda(x, weight, out, eps):
    """awq_ext.layernorm_forward_cuda (fused/norm.py:33-36): RMSNorm written into `out`."""
    _require_cuda(x, weight, out)
    if x.dtype != torch.float16 or weight.dtype != torch.float16 or out.dtype != torch.float16:
        raise B200AwqError("b200awq: rmsnorm expects float16 tensors")
    hidden = x.shape[-1]
    xc = x if x.is_contiguous() else x.contiguous()
    if not out.is_contiguous():
        raise B200AwqError("b200awq: rmsnorm output must be contiguous")
    rows = xc.numel() // hidden
    with _DeviceGuard(x.device):
        code = lib.b200awq_rmsnorm(xc.data_ptr(), weight.data_ptr(), out.data_ptr(), rows, hidden, float(eps),
                                   _stream(x.device))
    check(code, "b200awq_rmsnorm")


def silu_and_mul(out, gate_up):
    """awq_ext.silu_and_mul (fused/moe.py:76): out[.., d] = silu(gate_up[.., :d]) * gate_up[.., d:]."""
    _require_cuda(out, gate_up)
    d = out.shape[-1]
    if gate_up.shape[-1] != 2 * d or not gate_up.is_contiguous() or not out.is_contiguous():
        raise B200AwqError("b200awq: silu_and_mul expects contiguous [.., 2d] -> [.., d]")
    rows = out.numel() // d
    with _DeviceGuard(out.device):
        code = lib.b200awq_silu_and_mul(gate_up.data_ptr(), out.data_ptr(), rows, d, _stream(out.device))
    check(code, "b200awq_silu_and_mul")


# ------------------------------------------------------------------------------------ MoE (awq_ext surface)
def topk_softmax(topk_weights, topk_ids, token_expert_indicies, gating_output):
    """awq_ext.topk_softmax (fused/moe.py:162-167): fills the three output tensors [M, topk] from gating_output
    [M, E] f32 (softmax over experts, top-k, not renormalised)."""
    _require_cuda(topk_weights, topk_ids, token_expert_indicies, gating_output)
    if gating_output.dtype != torch.float32 or topk_weights.dtype != torch.float32 \
            or topk_ids.dtype != torch.int32 or token_expert_indicies.dtype != torch.int32:
        raise B200AwqError("b200awq: topk_softmax expects f32 gating / weights and i32 index tensors")
    g = gating_output if gating_output.is_contiguous() else gating_output.contiguous()
    for t in (topk_weights, topk_ids, token_expert_indicies):
        if not t.is_contiguous():
            raise B200AwqError("b200awq: topk_softmax outputs must be contiguous")
    M, E = g.shape
    topk = topk_weights.shape[-1]
    with _DeviceGuard(g.device):
        code = lib.b200awq_topk_softmax(g.data_ptr(), topk_weights.data_ptr(), topk_ids.data_ptr(),
                                        token_expert_indicies.data_ptr(), M, E, topk, _stream(g.device))
    check(code, f"b200awq_topk_softmax(M={M}, E={E}, topk={topk})")


def moe_alig_block_size(topk_ids, num_experts, block_size, sorted_token_ids, expert_ids, num_tokens_post_pad):
    """awq_ext.moe_alig_block_size (sic; fused/moe.py:131-133): fills sorted_token_ids, expert_ids,
    num_tokens_post_pad from topk_ids [M, topk] i32."""
    _require_cuda(topk_ids, sorted_token_ids, expert_ids, num_tokens_post_pad)
    for t in (topk_ids, sorted_token_ids, expert_ids, num_tokens_post_pad):
        if t.dtype != torch.int32 or not t.is_contiguous():
            raise B200AwqError("b200awq: moe_alig_block_size expects contiguous int32 tensors")
    numel = topk_ids.numel()
    if sorted_token_ids.numel() < numel + num_experts * (block_size - 1) or expert_ids.numel() < numel + num_experts:
        raise B200AwqError("b200awq: moe_alig_block_size output tensors are too small")
    with _DeviceGuard(topk_ids.device):
        code = lib.b200awq_moe_align_block_size(topk_ids.data_ptr(), numel, int(num_experts), int(block_size),
                                                sorted_token_ids.data_ptr(), expert_ids.data_ptr(),
                                                num_tokens_post_pad.data_ptr(), _stream(topk_ids.device))
    check(code, "b200awq_moe_align_block_size")


def grouped_gemm_forward(x, qweight, scales, qzeros, topk_weights, sorted_token_ids, expert_ids,
                         num_tokens_post_padded, mul_weights, split_k_iters=8):
    """awq_ext.grouped_gemm_forward (fused/moe.py:60-89): x [T, 1 or topk, K] f16, stacked expert weights
    qweight [E, K, N/8] / scales [E, K/G, N] / qzeros [E, K/G, N/8] -> [T, topk, N] f16."""
    _require_cuda(x, qweight, scales, qzeros, topk_weights, sorted_token_ids, expert_ids, num_tokens_post_padded)
    _check_w(qweight, torch.int32, "qweight")
    _check_w(scales, torch.float16, "scales")
    _check_w(qzeros, torch.int32, "qzeros")
    if x.dim() != 3 or x.dtype != torch.float16:
        raise B200AwqError("b200awq: grouped_gemm_forward expects x [T, 1 or topk, K] float16")
    if topk_weights.dtype != torch.float32 or not topk_weights.is_contiguous():
        raise B200AwqError("b200awq: topk_weights must be contiguous float32")
    E, K, N = qweight.shape[0], qweight.shape[1], qweight.shape[2] * 8
    G = K // scales.shape[1]
    T, topk = topk_weights.shape
    xc = x if x.is_contiguous() else x.contiguous()
    if xc.shape[0] != T or xc.shape[1] not in (1, topk) or xc.shape[2] != K:
        raise B200AwqError(f"b200awq: x {tuple(x.shape)} does not match T={T}, topk={topk}, K={K}")
    y = torch.empty((T, topk, N), dtype=torch.float16, device=x.device)
    with _DeviceGuard(x.device):
        st = _stream(x.device)
        slen = sorted_token_ids.numel()
        # decode-sized calls get the split-K scratch of the persistent kernel (8 rows of fp32 per 8 sorted slots);
        # beyond 64 MB of scratch the library's workspace-free grouped kernel runs instead
        need = 16384 + slen * N * 4
        ws = _workspace(x.device, st, need) if need <= (64 << 20) else None
        code = lib.b200awq_grouped_gemm_forward(
            xc.data_ptr(), int(xc.shape[1]), qweight.data_ptr(), scales.data_ptr(), qzeros.data_ptr(),
            topk_weights.data_ptr(), sorted_token_ids.data_ptr(), expert_ids.data_ptr(),
            num_tokens_post_padded.data_ptr(), y.data_ptr(), T, topk, slen, E, K, N, G,
            1 if mul_weights else 0, 16, ws.data_ptr() if ws is not None else None,
            ws.numel() if ws is not None else 0, st)
    check(code, f"b200awq_grouped_gemm_forward(T={T}, topk={topk}, E={E}, K={K}, N={N}, G={G})")
    return y


# ---------------------------------------------------------------------------- awq_v2_ext surface
def _fast_group_size(K: int, rows: int) -> int:
    from .packing import calculate_zeros_width

    for g in (128, 64, 32):
        if K % g == 0 and calculate_zeros_width(K, g) * 8 == rows:
            return g
    raise B200AwqError(f"b200awq: cannot infer group size from scales rows={rows}, K={K}")


def gemv_forward_cuda_decode(x, qweight, scales, szeros, m, n, k, group_size):
    """awq_v2_ext.gemv_forward_cuda_decode (gemv_fast.py:192-201): x [B, 1, K] -> [B, 1, N]."""
    out = linear_forward("fast", x, qweight, scales, szeros, group_size)
    return out.reshape(x.shape[:-1] + (out.shape[-1],))


def gemm_forward_cuda_prefill(x, qweight, scales, szeros):
    """awq_v2_ext.gemm_forward_cuda_prefill (gemv_fast.py:203-205): x [B, S, K] -> [B, S, N]."""
    K = qweight.shape[1]
    out = linear_forward("fast", x, qweight, scales, szeros, _fast_group_size(K, scales.shape[0]))
    return out.reshape(x.shape[:-1] + (out.shape[-1],))


def stream_pack(qweight, scales, qzeros, mode: int = 0) -> torch.Tensor:
    """One-time re-layout of a GEMM-layout linear into the stream format the decode-program kernel reads
    (include/b200awq.h; the post_init-style hook, cf. awq/modules/linear/exllama.py:66-79).  Returns a uint8 tensor."""
    _require_cuda(qweight, scales, qzeros)
    _check_w(qweight, torch.int32, "qweight")
    _check_w(scales, torch.float16, "scales")
    _check_w(qzeros, torch.int32, "qzeros")
    K, N = qweight.shape[0], qweight.shape[1] * 8
    G = K // scales.shape[0]
    nbytes = lib.b200awq_stream_bytes(K, N, G)
    if nbytes == 0:
        raise B200AwqError(f"b200awq: no stream format for K={K}, N={N}, G={G}")
    out = torch.empty(nbytes, dtype=torch.uint8, device=qweight.device)
    with _DeviceGuard(qweight.device):
        code = lib.b200awq_stream_pack(qweight.data_ptr(), scales.data_ptr(), qzeros.data_ptr(), out.data_ptr(), K, N, G,
                                       int(mode), _stream(qweight.device))
    check(code, f"b200awq_stream_pack(K={K}, N={N}, G={G}, mode={mode})")
    return out


def set_knob(key: int, value: int) -> None:
    check(lib.b200awq_set_knob(key, value), "b200awq_set_knob")


def get_knob(key: int) -> int:
    return lib.b200awq_get_knob(key)

"""The oracle (oracle/awq_oracle.py) against vectors produced by the real reference
(tests/golden/make_golden.py).  CPU only.  Pins the oracle before anything trusts it."""
import hashlib
import os

import numpy as np
import pytest

from oracle import awq_oracle as O


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint16)


def test_known_answer_word(golden_dir):
    g = np.load(os.path.join(golden_dir, "dequant_small.npz"))
    qw = np.full((8, 2), 0x76543210, dtype=np.int32)
    w = O.dequantize_gemm(qw, np.zeros((1, 2), np.int32), np.ones((1, 16), np.float16), 8)
    assert np.array_equal(w[0, :8], np.array([0, 4, 1, 5, 2, 6, 3, 7], dtype=np.float16))
    assert np.array_equal(_bits(w), _bits(g["known_answer_w"]))


def test_dequant_small_bit_exact(golden_dir):
    g = np.load(os.path.join(golden_dir, "dequant_small.npz"))
    for m in g["meta"]:
        tag, K, N, G = str(m).split(",")
        w = O.dequantize_gemm(g[f"{tag}_qweight"], g[f"{tag}_qzeros"], g[f"{tag}_scales"], int(G))
        ref = g[f"{tag}_w"]
        assert w.shape == ref.shape == (int(K), int(N))
        assert np.array_equal(_bits(w), _bits(ref)), tag


@pytest.mark.parametrize("N", [1792, 4096])
def test_dequant_reference_test_shape(golden_dir, N):
    """tests/test_dequantization.py recipe (raw int32 words, randn scales, K=4096, g=128)."""
    g = np.load(os.path.join(golden_dir, "dequant_ref_shape.npz"))
    c = O.make_case(4096, N, 128, seed=0, raw=True)
    w = O.dequantize_gemm(c["qweight"], c["qzeros"], c["scales"], 128)
    assert _sha(w) == str(g[f"n{N}_sha256"])
    assert np.array_equal(_bits(w[g[f"n{N}_rows"]]), _bits(g[f"n{N}_w_rows"]))
    c2 = O.make_case(4096, N, 128, seed=1, raw=False)
    w2 = O.dequantize_gemm(c2["qweight"], c2["qzeros"], c2["scales"], 128)
    assert _sha(w2) == str(g[f"n{N}_can_sha256"])


def _canon_from_golden(g, tag):
    K, N, G = (int(v) for v in g[f"{tag}_meta"])
    w = g[f"{tag}_weight"].astype(np.float32)  # [N, K] pseudo-quantised fp16 weights
    s = g[f"{tag}_scales_ng"].astype(np.float32)  # [N, K/G]
    z = g[f"{tag}_zeros_ng"].astype(np.float32)
    iw = np.round((w + np.repeat(z * s, G, axis=1)) / np.repeat(s, G, axis=1)).astype(np.int64)
    return K, N, G, iw.T.astype(np.uint8), z.T.astype(np.uint8), s.T.astype(np.float16)


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_packers_match_from_linear(golden_dir, tag):
    g = np.load(os.path.join(golden_dir, "packers.npz"))
    K, N, G, iw, iz, s = _canon_from_golden(g, tag)
    qw, qz = O.pack_gemm(iw, iz)
    assert np.array_equal(qw, g[f"{tag}_gemm_qweight"])
    assert np.array_equal(qz, g[f"{tag}_gemm_qzeros"])
    assert np.array_equal(_bits(s), _bits(g[f"{tag}_gemm_scales"]))
    vw, vz, vs = O.pack_gemv(iw, iz, s, G)
    assert np.array_equal(vw, g[f"{tag}_gemv_qweight"])
    assert np.array_equal(vz, g[f"{tag}_gemv_qzeros"])
    assert np.array_equal(_bits(vs), _bits(g[f"{tag}_gemv_scales"]))
    if f"{tag}_fast_qweight" in g:
        fw, fs, fz = O.pack_gemv_fast(iw, iz, s, G)
        assert np.array_equal(fw, g[f"{tag}_fast_qweight"])
        assert np.array_equal(_bits(fs), _bits(g[f"{tag}_fast_scales"]))
        assert np.array_equal(_bits(fz), _bits(g[f"{tag}_fast_qzeros"]))
        assert np.array_equal(O.unpack_gemv_fast_weight(fw), iw)
    # unpack inverts pack, and the three layouts encode the same weights
    assert np.array_equal(O.unpack_gemm(qw), iw)
    wg = O.dequantize_gemm(qw, qz, s, G)
    wv = O.dequantize_gemv(vw, vz, vs, G)
    assert np.array_equal(_bits(wg), _bits(wv))


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_forward_matches_reference_module(golden_dir, tag):
    """WQLinear_GEMM.forward (naive CPU branch incl. fp16 bias add) vs the oracle forward.
    The reference rounds a torch CPU fp16 matmul; ours rounds an fp64 contraction: they may
    differ by one fp16 ulp of the pre-bias value, hence rtol 2^-9 with a small atol."""
    g = np.load(os.path.join(golden_dir, "packers.npz"))
    K, N, G = (int(v) for v in g[f"{tag}_meta"])
    for xi in range(3):
        x, yref = g[f"{tag}_x{xi}"], g[f"{tag}_y{xi}"]
        y = O.wqlinear_forward(x, g[f"{tag}_gemm_qweight"], g[f"{tag}_gemm_qzeros"], g[f"{tag}_gemm_scales"], G,
                               bias=g[f"{tag}_bias"])
        if x.ndim == 2:  # reference returns 3-D for 2-D input then reshapes back (gemm.py:83-84,287)
            assert yref.shape == (x.shape[0], N)
        assert y.shape == yref.shape
        np.testing.assert_allclose(y.astype(np.float32), yref.astype(np.float32), rtol=2**-9, atol=2e-3)


def test_zeros_width_table(golden_dir):
    g = np.load(os.path.join(golden_dir, "packers.npz"))
    for (k, gs), zw in zip(g["zw_in"], g["zw_out"]):
        assert O.calculate_zeros_width(int(k), int(gs)) == int(zw)


def test_quantize_rtn_roundtrip():
    rng = np.random.default_rng(3)
    w = rng.standard_normal((64, 256)).astype(np.float32) * 0.05
    iw, iz, s = O.quantize_rtn(w, 64)
    assert iw.max() <= 15 and iz.max() <= 15
    qw, qz = O.pack_gemm(iw, iz)
    wd = O.dequantize_gemm(qw, qz, s, 64).astype(np.float32)
    # round-to-nearest with step s: error at most s/2 (+ fp16 rounding)
    step = np.repeat(s.astype(np.float32), 64, axis=0)
    assert np.all(np.abs(wd - w.T) <= 0.5 * step * 1.01 + 1e-4)


def test_column_concat_is_format_preserving():
    """fuse_qkv concatenates packed tensors along N (fused_utils.py:87-96): legal because the
    interleave is per 8 columns. Same property licenses N-sharding on 8-column boundaries."""
    a = O.make_case(128, 64, 32, seed=5)
    b = O.make_case(128, 32, 32, seed=6)
    qw = np.concatenate([a["qweight"], b["qweight"]], axis=1)
    qz = np.concatenate([a["qzeros"], b["qzeros"]], axis=1)
    s = np.concatenate([a["scales"], b["scales"]], axis=1)
    w = O.dequantize_gemm(qw, qz, s, 32)
    wa = O.dequantize_gemm(a["qweight"], a["qzeros"], a["scales"], 32)
    wb = O.dequantize_gemm(b["qweight"], b["qzeros"], b["scales"], 32)
    assert np.array_equal(_bits(w), _bits(np.concatenate([wa, wb], axis=1)))


def test_moe_align_oracle_matches_reference_docstring_example():
    """awq/modules/fused/moe.py:104-113 spells out one worked example; it is the only vector the reference holds for
    the MoE path, so the oracle is pinned to it."""
    import numpy as np

    from oracle import awq_oracle as O

    ids = np.array([[2, 3, 4], [1, 2, 4], [1, 3, 4], [1, 2, 3]])
    s, e, n = O.moe_align_block_size(ids, 4, 5)
    assert n == 16 and s[:16].tolist() == [3, 6, 9, 12, 0, 4, 10, 12, 1, 7, 11, 12, 2, 5, 8, 12]
    assert e[:4].tolist() == [1, 2, 3, 4]


def test_moe_topk_and_grouped_gemm_oracle_self_consistency():
    import numpy as np

    from oracle import awq_oracle as O

    g = np.array([[0.0, 1.0, 1.0, -1.0]], dtype=np.float32)
    w, i, src = O.topk_softmax(g, 2)
    assert i.tolist() == [[1, 2]] and abs(float(w.sum()) - 2 * float(w[0, 0])) < 1e-7 and src.tolist() == [[0, 1]]
    # grouped GEMM == per-token dense matmul with the selected expert
    rng = np.random.default_rng(0)
    E, K, N, T, topk = 3, 8, 4, 4, 2
    W = rng.standard_normal((E, K, N))
    x = rng.standard_normal((T, 1, K))
    ids = np.array([[0, 1], [2, 0], [1, 2], [0, 2]])
    tw = rng.random((T, topk))
    s, e, n = O.moe_align_block_size(ids, 16, E)
    out = O.grouped_gemm_f64(x, W, tw, s, e, n, True)
    for t in range(T):
        for k in range(topk):
            np.testing.assert_allclose(out[t, k], x[t, 0] @ W[ids[t, k]] * tw[t, k], rtol=1e-12)

"""The stream format's numpy restatement (oracle/stream_format.py) pinned to the reference's semantics on the CPU:
a GEMV evaluated FROM THE STREAM BUFFER, nibble by nibble in the kernel's lane / fragment order, must equal the
contraction with the reference-defined dequantised weights (oracle/awq_oracle.py:dequantize_gemm, itself pinned to
awq/utils/packing_utils.py:87-102 by the golden vectors) - for every group size, both column maps, raw int32 words."""
import numpy as np
import pytest

from oracle import awq_oracle as O
from oracle import stream_format as SF


@pytest.mark.parametrize("K,N,G,mode", [(256, 32, 128, 0), (128, 32, 32, 0), (256, 48, 64, 0), (256, 64, 128, 1),
                                        (256, 32, -1, 0), (384, 16, 128, 0)])
def test_stream_buffer_reproduces_reference_dequant(K, N, G, mode):
    c = O.make_case(K, N, G, seed=K + N, raw=True)
    Gs = c["group_size"]
    st = SF.pack_stream(c["qweight"], c["qzeros"], c["scales"], Gs, mode)
    assert st.dtype == np.uint8 and st.size == SF.stream_bytes(K, N, Gs)
    iw, iz = SF.unpack_gemm_ints(c["qweight"], c["qzeros"])
    # exact real-number weights (q - z) * s; the reference rounds each to fp16 (dequantize_gemm) - compare both ways
    w_exact = (iw.astype(np.float64) - np.repeat(iz.astype(np.float64), Gs, axis=0)) * \
        np.repeat(c["scales"].astype(np.float64), Gs, axis=0)
    w_ref = O.dequantize_gemm(c["qweight"], c["qzeros"], c["scales"], Gs).astype(np.float64)
    fin = np.isfinite(w_ref)   # raw randn scales * 15 stay far below fp16 max, but be explicit
    assert fin.all()
    assert np.all(np.abs(w_exact - w_ref) <= 2.0**-11 * np.abs(w_exact) + 1e-12)
    x = np.random.default_rng(1).standard_normal(K).astype(np.float16)
    y = SF.simulate_gemv(st, K, N, Gs, x, mode)
    ref = x.astype(np.float64) @ w_exact
    np.testing.assert_allclose(y, ref, rtol=1e-9, atol=1e-9)


def test_set_columns_cover_every_column_once():
    for N, mode in [(32, 0), (4096, 0), (64, 1), (28672, 1)]:
        cols = SF.set_columns(N, mode)
        assert sorted(cols.reshape(-1).tolist()) == list(range(N))
        if mode == 1:   # gate column j and up column N/2 + j share tile rows g / g + 8
            assert np.array_equal(cols[:, 8:] - cols[:, :8], np.full((N // 16, 8), N // 2))


def test_unit_geometry():
    assert (SF.unit_k(128), SF.unit_bytes(128)) == (128, 1072)
    assert (SF.unit_k(64), SF.unit_bytes(64)) == (64, 560)
    assert (SF.unit_k(32), SF.unit_bytes(32)) == (32, 304)
    assert SF.unit_k(4096) == 128
    # algorithmic bytes per weight of a g128 linear: 0.5 + 2/128 + 0.5/128 = 0.51953; the stream format adds 8 pad
    # bytes per 1024-byte unit (0.0039 B/weight)
    assert abs(SF.stream_bytes(4096, 4096, 128) / (4096 * 4096) - (0.51953125 + 8 / 2048)) < 1e-9

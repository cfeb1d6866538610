"""CPU: KV-cache oracle self-checks (layout sizes, quantizer properties)."""
import numpy as np
import pytest

from oracle import kvcache_ref as KV


def test_span_bytes_formula():
    assert KV.span_bytes(KV.QUANT_NONE, 128, 4) == 128 * 4 * 128 * 2
    assert KV.span_bytes(KV.QUANT_I8, 128, 4) == 128 * 4 * 128 + 2 * 128 * 4 * 4
    assert KV.span_bytes(KV.QUANT_U4, 16, 8) == 16 * 8 * 64 + 2 * 16 * 8 * 4


@pytest.mark.parametrize("mode", [KV.QUANT_I8, KV.QUANT_U4])
def test_quant_rows_properties(mode):
    rng = np.random.default_rng(3)
    x = rng.standard_normal((50, 128)).astype(np.float32)
    x[7] = 0.25  # constant row: scale clamps to EPS
    q, z, s = KV.quant_rows(x, mode)
    lo, hi = (-128, 127) if mode == KV.QUANT_I8 else (0, 15)
    assert q.min() >= lo and q.max() <= hi
    assert np.all(z == np.rint(z))
    assert s[7] == np.float32(1e-5)
    xd = KV.dequant_rows(q, z, s)
    rows = np.arange(50) != 7
    assert np.max(np.abs(xd[rows] - x[rows]) / s[rows, None]) <= 1.0 + 1e-3  # zero rounding + value rounding


@pytest.mark.parametrize("mode", [KV.QUANT_NONE, KV.QUANT_I8, KV.QUANT_U4])
def test_span_cache_roundtrip(mode):
    rng = np.random.default_rng(4)
    c = KV.SpanCacheRef(mode, 16, 2)
    b = c.add_sequence()
    rows = KV.bits_to_f32(KV.bf16_bits(rng.standard_normal((40, 2, 128)).astype(np.float32)))
    for t in range(40):
        c.append(b, t, rows[t])
    assert len(c.spans[b]) == 3
    d = c.dense(b, 40)
    tol = {KV.QUANT_NONE: 0.0, KV.QUANT_I8: 0.05, KV.QUANT_U4: 0.6}[mode]
    assert np.max(np.abs(d - rows.transpose(1, 0, 2))) <= tol


def test_attention_ref_uniform():
    c_k = KV.SpanCacheRef(KV.QUANT_NONE, 16, 1)
    c_v = KV.SpanCacheRef(KV.QUANT_NONE, 16, 1)
    c_k.add_sequence(); c_v.add_sequence()
    for t in range(5):
        c_k.append(0, t, np.zeros((1, 128), np.float32))
        c_v.append(0, t, np.full((1, 128), float(t), np.float32))
    out = KV.attention_ref(np.ones((1, 2, 128), np.float32), c_k, c_v, [5], 2, 0.1)
    assert np.allclose(out, 2.0)  # uniform softmax over values 0..4

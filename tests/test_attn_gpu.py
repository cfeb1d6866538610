"""GPU parity: span cache append (vs the oracle's span bytes; bit-exact vs the reference kernel in test_ref_pin_gpu.py) and
paged attention through the C ABI.

Attention tolerance: <= 2e-3 abs against fp32/fp64 attention on the SAME (dequantized) cache contents
(BASELINE.md §3), inputs N(0,1) like the reference's span-attention tests (test_quant_none.cpp)."""
import numpy as np
import pytest
import torch

from oracle import kvcache_ref as KV

pytestmark = pytest.mark.gpu


def _bf16(x):
    return torch.from_numpy(x).to(torch.bfloat16)


def _build(mode, B, lens, nH, nG, span, seed, max_len=None, fill=0, mirror=False, dtype=torch.bfloat16):
    # mirror=True keeps the oracle's OWN quantization of the same rows (append parity tests)
    """Fill a device SpanCache token by token with the product append kernel and mirror it in the oracle."""
    from b200spark import ops
    rng = np.random.default_rng(seed)
    max_len = max_len or (max(lens) + 1)
    cache = ops.SpanCache(B, max_len, nH, nG, span, mode, fill=fill, dtype=dtype)
    ft = "fp16" if dtype == torch.float16 else "bf16"
    kref, vref = KV.SpanCacheRef(mode, span, nG, ft=ft), KV.SpanCacheRef(mode, span, nG, ft=ft)
    for _ in range(B):
        kref.add_sequence(); vref.add_sequence()
    T = max(lens)
    width = (nH + 2 * nG) * 128
    q_last = np.zeros((B, nH, 128), np.float32)
    cur = torch.zeros(B, dtype=torch.int32, device="cuda")
    for t in range(T):
        qkv = torch.from_numpy(rng.standard_normal((B, width)).astype(np.float32)).to(dtype)
        # sequences already at their final length keep re-writing a scratch position beyond their length
        pos = torch.tensor([min(t, lens[b]) for b in range(B)], dtype=torch.int32, device="cuda")
        q = ops.cache_append(cache, qkv.cuda(), pos)
        x = qkv.float().numpy().reshape(B, nH + 2 * nG, 128)
        for b in range(B):
            if t < lens[b]:
                kref.append(b, t, x[b, nH:nH + nG]); vref.append(b, t, x[b, nH + nG:])
                q_last[b] = x[b, :nH]
        if t == T - 1:
            assert torch.equal(q.cpu().float().reshape(B, nH, 128), qkv.float().reshape(B, -1, 128)[:, :nH])
    torch.cuda.synchronize()
    if mode != KV.QUANT_NONE and not mirror:
        # attention tests attend over IDENTICAL cache bytes: the oracle reads back what the append kernel stored (the kernel
        # follows the reference kernel's MUFU.RCP arithmetic bit for bit, the CPU oracle's quantizer uses an IEEE reciprocal
        # and lands on the other side of exact zero-point ties in ~1 % of the rows — test_append_against_oracle bounds that)
        for ref, which in ((kref, "k"), (vref, "v")):
            for b in range(B):
                for si in range(len(ref.spans[b])):
                    ref.spans[b][si] = cache.span_view(which, b, si).cpu().numpy().copy()
    return cache, kref, vref, q_last


def _codes(buf, mode, nG, span, n):
    row = {KV.QUANT_I8: 128, KV.QUANT_U4: 64}[mode]
    d = buf[:nG * span * row].reshape(nG, span, row)[:, :n]
    if mode == KV.QUANT_I8:
        return d.view(np.int8).astype(np.int32)
    return np.stack([d & 0xF, d >> 4], -1).reshape(nG, n, 128).astype(np.int32)


@pytest.mark.parametrize("mode", [KV.QUANT_NONE, KV.QUANT_I8, KV.QUANT_U4])
@pytest.mark.parametrize("span", [16, 128])
def test_append_against_oracle(mode, span):
    """bf16 spans: byte-identical.  I8 / U4: the kernel repeats the reference kernel's compiled arithmetic (bit-exact
    against the reference itself, tests/test_ref_pin_gpu.py); the CPU oracle repeats the same formula with an IEEE
    reciprocal where the GPU uses MUFU.RCP, so: scales identical, zero points identical except on exact-tie rows
    (|diff| = 1, < 3 % of rows), codes identical on every row whose zero point agrees up to isolated +-1 (< 0.5 %)."""
    B, nH, nG = 3, 8, 2
    lens = [37, 5, 130]
    cache, kref, vref, _ = _build(mode, B, lens, nH, nG, span, seed=span + mode, max_len=140, mirror=True)
    rows = zdiff = codes = cdiff = 0
    for b in range(B):
        for si in range((lens[b] + span - 1) // span):
            n = min(span, lens[b] - si * span)
            for which, ref in (("k", kref), ("v", vref)):
                got = cache.span_view(which, b, si).cpu().numpy()
                exp = ref.spans[b][si]
                if mode == KV.QUANT_NONE:
                    for g in range(nG):
                        a0 = g * span * 256
                        assert np.array_equal(got[a0:a0 + n * 256], exp[a0:a0 + n * 256]), (span, b, si, which, g)
                    continue
                row = {KV.QUANT_I8: 128, KV.QUANT_U4: 64}[mode]
                gp = got[nG * span * row:].view(np.float32).reshape(nG, span, 2)[:, :n]
                ep = exp[nG * span * row:].view(np.float32).reshape(nG, span, 2)[:, :n]
                assert np.array_equal(gp[..., 1], ep[..., 1]), "scale = (max - min) * fl(1/RANGE): no approximation involved"
                dz = np.abs(gp[..., 0] - ep[..., 0])
                assert dz.max() <= 1.0
                same = dz == 0
                gc, ec = _codes(got, mode, nG, span, n), _codes(exp, mode, nG, span, n)
                dc = np.abs(gc - ec)
                assert dc.max() <= 2 and dc[same].max(initial=0) <= 1
                rows += same.size; zdiff += int((~same).sum()); codes += int(same.sum()) * 128; cdiff += int((dc[same] != 0).sum())
    if mode != KV.QUANT_NONE:
        assert zdiff / rows < 3e-2 and cdiff / max(codes, 1) < 5e-3, (zdiff, rows, cdiff, codes)


@pytest.mark.parametrize("span", [16, 32, 64, 128])
@pytest.mark.parametrize("nH,nG", [(8, 2), (7, 1), (28, 4), (16, 1)])
def test_attention_none_small(span, nH, nG):
    from b200spark import ops
    lens = [1, 63, 64, 65, 200]
    B = len(lens)
    cache, kref, vref, q = _build(KV.QUANT_NONE, B, lens, nH, nG, span, seed=span * 31 + nH, max_len=256)
    attn = ops.SpanAttn(cache.cfg, B)
    ws = ops.Workspace()
    new_lens = torch.tensor(lens, dtype=torch.int32, device="cuda")
    out = attn(_bf16(q.reshape(B, -1)).cuda(), cache, new_lens, 256, ws)
    torch.cuda.synchronize()
    ref = KV.attention_ref(q, kref, vref, lens, nH, 1.0 / np.sqrt(128))
    err = np.abs(out.float().cpu().numpy().reshape(B, nH, 128) - ref).max()
    assert err <= 2e-3 + 4e-3, err  # + bf16 output rounding of values up to ~1 (2^-9)


def test_attention_none_long_and_ragged():
    """ctx 2048 / 4100 with split-KV partials + last-CTA combine, Qwen2-7B head geometry."""
    from b200spark import ops
    nH, nG, span = 28, 4, 128
    lens = [2048, 4100, 777, 2049]
    B = len(lens)
    cache, kref, vref, q = _build(KV.QUANT_NONE, B, lens, nH, nG, span, seed=5, max_len=4224)
    attn = ops.SpanAttn(cache.cfg, B)
    ws = ops.Workspace()
    new_lens = torch.tensor(lens, dtype=torch.int32, device="cuda")
    qd = _bf16(q.reshape(B, -1)).cuda()
    out = attn(qd, cache, new_lens, 4224, ws)
    out2 = attn(qd, cache, new_lens, 4224, ws)
    torch.cuda.synchronize()
    assert torch.equal(out, out2)  # deterministic, counters re-armed
    ref = KV.attention_ref(q, kref, vref, lens, nH, 1.0 / np.sqrt(128))
    err = np.abs(out.float().cpu().numpy().reshape(B, nH, 128) - ref).max()
    assert err <= 6e-3, err


def _close(got, ref):
    # 2e-3 abs + bf16 rounding of the stored output / probabilities (2^-7 relative envelope)
    return np.all(np.abs(got - ref) <= 2e-3 + 2.0 ** -7 * np.abs(ref)), float(np.abs(got - ref).max())


@pytest.mark.parametrize("mode", [KV.QUANT_I8, KV.QUANT_U4])
@pytest.mark.parametrize("span", [16, 128])
@pytest.mark.parametrize("nH,nG", [(8, 2), (28, 4), (16, 1)])
def test_attention_quantized_small(mode, span, nH, nG):
    """I8 / U4 spans: attention against fp64 attention over the DEQUANTIZED oracle cache (same bytes, bit-exact append)."""
    from b200spark import ops
    lens = [1, 63, 64, 65, 200]
    B = len(lens)
    cache, kref, vref, q = _build(mode, B, lens, nH, nG, span, seed=span * 7 + nH + mode, max_len=256)
    attn = ops.SpanAttn(cache.cfg, B)
    ws = ops.Workspace()
    new_lens = torch.tensor(lens, dtype=torch.int32, device="cuda")
    out = attn(_bf16(q.reshape(B, -1)).cuda(), cache, new_lens, 256, ws)
    torch.cuda.synchronize()
    ref = KV.attention_ref(q, kref, vref, lens, nH, 1.0 / np.sqrt(128))
    ok, err = _close(out.float().cpu().numpy().reshape(B, nH, 128), ref)
    assert ok, err


@pytest.mark.parametrize("mode", [KV.QUANT_I8, KV.QUANT_U4])
def test_attention_quantized_long(mode):
    from b200spark import ops
    nH, nG, span = 28, 4, 128
    lens = [2048, 3000, 777]
    B = len(lens)
    cache, kref, vref, q = _build(mode, B, lens, nH, nG, span, seed=11 + mode, max_len=3072)
    attn = ops.SpanAttn(cache.cfg, B)
    ws = ops.Workspace()
    new_lens = torch.tensor(lens, dtype=torch.int32, device="cuda")
    qd = _bf16(q.reshape(B, -1)).cuda()
    out = attn(qd, cache, new_lens, 3072, ws)
    out2 = attn(qd, cache, new_lens, 3072, ws)
    torch.cuda.synchronize()
    assert torch.equal(out, out2)
    ref = KV.attention_ref(q, kref, vref, lens, nH, 1.0 / np.sqrt(128))
    ok, err = _close(out.float().cpu().numpy().reshape(B, nH, 128), ref)
    assert ok, err


def test_attention_i8_ctx_32768():
    """Maximum size of config C2: one sequence of 32768 tokens, int8 KV spans, Qwen2-7B head geometry —
    256 spans, every CTA of the persistent grid takes part: ~130 split-KV pieces per kv-head, merged in two levels
    (groups of 8 by their last CTA, then the groups)."""
    from b200spark import ops
    nH, nG, span, L = 28, 4, 128, 32768
    rng = np.random.default_rng(77)
    cache = ops.SpanCache(1, L, nH, nG, span, KV.QUANT_I8)
    kref, vref = KV.SpanCacheRef(KV.QUANT_I8, span, nG), KV.SpanCacheRef(KV.QUANT_I8, span, nG)
    kref.add_sequence(); vref.add_sequence()
    width = (nH + 2 * nG) * 128
    q_last = None
    # the append kernel writes one token per sequence per call: feed it 32768 times (positions on the device)
    rows = _bf16(rng.standard_normal((L, width)).astype(np.float32))
    rows_d = rows.cuda()
    pos = torch.zeros(1, dtype=torch.int32, device="cuda")
    for t in range(L):
        q = ops.cache_append(cache, rows_d[t:t + 1], pos)
        ops.lens_add(pos, 1)
    x = rows.float().numpy().reshape(L, nH + 2 * nG, 128)
    for t in range(L):
        kref.append(0, t, x[t, nH:nH + nG]); vref.append(0, t, x[t, nH + nG:])
    q_last = x[L - 1, :nH][None]
    attn = ops.SpanAttn(cache.cfg, 1)
    ws = ops.Workspace()
    out = attn(q.reshape(1, -1), cache, torch.tensor([L], dtype=torch.int32, device="cuda"), L, ws)
    out2 = attn(q.reshape(1, -1), cache, torch.tensor([L], dtype=torch.int32, device="cuda"), L, ws)
    torch.cuda.synchronize()
    assert torch.equal(out, out2)
    ref = KV.attention_ref(q_last, kref, vref, [L], nH, 1.0 / np.sqrt(128))
    ok, err = _close(out.float().cpu().numpy().reshape(1, nH, 128), ref)
    assert ok, err


@pytest.mark.parametrize("mode", [KV.QUANT_NONE, KV.QUANT_I8, KV.QUANT_U4])
@pytest.mark.parametrize("L", [1, 37, 129, 191])
def test_attention_ignores_unwritten_span_memory(mode, L):
    """the reference's span manager never zeroes frames.  The pool is filled with 0xFF (NaN as bf16 and as the
    fp32 {zero, scale} params) before appending; odd lengths leave the unwritten token `len` inside the last 16-byte
    parameter chunk and inside the last 64-token tile.  Nothing of it may reach the output."""
    from b200spark import ops
    nH, nG, span = 28, 4, 32
    lens = [L, L, L]
    B = len(lens)
    cache, kref, vref, q = _build(mode, B, lens, nH, nG, span, seed=1000 + L + mode, max_len=256, fill=0xFF)
    attn = ops.SpanAttn(cache.cfg, B)
    ws = ops.Workspace()
    new_lens = torch.tensor(lens, dtype=torch.int32, device="cuda")
    out = attn(_bf16(q.reshape(B, -1)).cuda(), cache, new_lens, 256, ws)
    torch.cuda.synchronize()
    got = out.float().cpu().numpy().reshape(B, nH, 128)
    assert np.isfinite(got).all(), "NaN/Inf leaked from unwritten span memory"
    ref = KV.attention_ref(q, kref, vref, lens, nH, 1.0 / np.sqrt(128))
    ok, err = _close(got, ref)
    assert ok, err


def test_attention_piece_cap_env(monkeypatch):
    """B2_ATTN_MAX_PIECES bounds the split of one (sequence, kv-head): with 4 pieces the direct last-CTA merge is used on a
    sequence that otherwise takes the two-level path — same result within fp32 reassociation."""
    from b200spark import ops
    nH, nG, span = 28, 4, 128
    lens = [2048]
    cache, kref, vref, q = _build(KV.QUANT_NONE, 1, lens, nH, nG, span, seed=21, max_len=2176)
    ws = ops.Workspace()
    new_lens = torch.tensor(lens, dtype=torch.int32, device="cuda")
    qd = _bf16(q.reshape(1, -1)).cuda()
    out_tree = ops.SpanAttn(cache.cfg, 1)(qd, cache, new_lens, 2176, ws)
    monkeypatch.setenv("B2_ATTN_MAX_PIECES", "4")
    out_cap = ops.SpanAttn(cache.cfg, 1)(qd, cache, new_lens, 2176, ws)
    torch.cuda.synchronize()
    ref = KV.attention_ref(q, kref, vref, lens, nH, 1.0 / np.sqrt(128))
    for out in (out_tree, out_cap):
        assert np.abs(out.float().cpu().numpy().reshape(1, nH, 128) - ref).max() <= 6e-3


# ---------------------------------------------------------------------------------------------------------------------
# fp16 Q / output / unquantized cache (the reference's span-attention tests run FP16 first: span-attention/test/
# test_quant_none.cpp); same kernels with the 16-bit type as a template parameter
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("mode", [KV.QUANT_NONE, KV.QUANT_I8, KV.QUANT_U4])
@pytest.mark.parametrize("span,nH,nG", [(16, 8, 2), (128, 28, 4), (64, 16, 1)])
def test_attention_fp16(mode, span, nH, nG):
    from b200spark import ops
    lens = [1, 63, 64, 65, 200, 700]
    B = len(lens)
    cache, kref, vref, q = _build(mode, B, lens, nH, nG, span, seed=span + nH + 3 * mode, max_len=768, dtype=torch.float16)
    attn = ops.SpanAttn(cache.cfg, B)
    ws = ops.Workspace()
    new_lens = torch.tensor(lens, dtype=torch.int32, device="cuda")
    qd = torch.from_numpy(q.reshape(B, -1)).to(torch.float16).cuda()
    out = attn(qd, cache, new_lens, 768, ws)
    out2 = attn(qd, cache, new_lens, 768, ws)
    torch.cuda.synchronize()
    assert out.dtype == torch.float16 and torch.equal(out, out2)
    ref = KV.attention_ref(q, kref, vref, lens, nH, 1.0 / np.sqrt(128))
    got = out.float().cpu().numpy().reshape(B, nH, 128)
    # fp16 output / probabilities: 2^-10 relative envelope on top of the 2e-3 absolute bound
    assert np.all(np.abs(got - ref) <= 2e-3 + 2.0 ** -9 * np.abs(ref)), float(np.abs(got - ref).max())


def test_attention_fp16_long_ragged_with_rope():
    """ctx 2048 / 4100 (split-KV partials, two-level merge) in fp16, rows appended through the fused-rotary path"""
    from b200spark import ops
    nH, nG, span = 28, 4, 128
    lens = [2048, 4100, 777]
    B = len(lens)
    cache, kref, vref, q = _build(KV.QUANT_NONE, B, lens, nH, nG, span, seed=15, max_len=4224, dtype=torch.float16)
    attn = ops.SpanAttn(cache.cfg, B)
    new_lens = torch.tensor(lens, dtype=torch.int32, device="cuda")
    out = attn(torch.from_numpy(q.reshape(B, -1)).to(torch.float16).cuda(), cache, new_lens, 4224, ops.Workspace())
    torch.cuda.synchronize()
    ref = KV.attention_ref(q, kref, vref, lens, nH, 1.0 / np.sqrt(128))
    assert np.abs(out.float().cpu().numpy().reshape(B, nH, 128) - ref).max() <= 3e-3
    # fused rotary in fp16: against the fp32 NeoX formula rounded to fp16
    rng = np.random.default_rng(3)
    x = torch.from_numpy(rng.standard_normal((2, (nH + 2 * nG) * 128)).astype(np.float32)).to(torch.float16)
    c2 = ops.SpanCache(2, 256, nH, nG, span, KV.QUANT_NONE, dtype=torch.float16)
    pos = torch.tensor([5, 130], dtype=torch.int32, device="cuda")
    qo = ops.cache_append(c2, x.cuda(), pos, rope=(1e6, 128))
    torch.cuda.synchronize()
    xf = x.float().numpy().reshape(2, nH + 2 * nG, 128)[:, :nH]
    inv = 1e6 ** (-np.arange(64, dtype=np.float32) * 2 / 128)
    ang = pos.cpu().numpy()[:, None].astype(np.float32) * inv[None, :]
    cs, sn = np.cos(ang)[:, None, :], np.sin(ang)[:, None, :]
    want = np.concatenate([xf[..., :64] * cs - xf[..., 64:] * sn, xf[..., 64:] * cs + xf[..., :64] * sn], -1)
    assert np.abs(qo.float().cpu().numpy().reshape(2, nH, 128) - want).max() <= 4e-3

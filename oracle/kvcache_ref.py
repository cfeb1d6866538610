"""Oracle: KV-cache span layout, I8/U4 row quantizer, cache append, fp32 paged attention.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Pinned on the GPU box against the reference's OWN kernels
(oracle/_ref/libdashinfer_ref.so = span-attention + the span-cache writers compiled from /root/reference by
oracle/build_ref.py): tests/test_ref_pin_gpu.py compares span bytes, {zero, scale} params and attention outputs.  The
reference holds no golden vector for quantized spans, so the pin is the reference executed.

Follows (paths relative to /root/reference):
  * span-attention/src/cache_quant/impl_i8.cuh:29-142   QuantParam<I8>: ORIGIN -128, RANGE 255, clamp [-128,127], EPS 1e-5
  * span-attention/src/cache_quant/impl_u4.cuh:20-184   QuantParam<U4>: ORIGIN 0, RANGE 15, upper clamp only, lo nibble first
  * span-attention/src/cache_quant/utils.cuh:24-45      Div (= __fdividef -> x * MUFU.RCP under --use_fast_math; IEEE reciprocal here),
                                                        Rounding = rintf (CONFIG_CACHE_ROUND_RNI)
  * csrc/core/kernel/cuda/cache/decoder_cache_append.cuh:33-153  span = [nG, spanLen, HEAD] QT then [nG, spanLen] {f32 zero, f32 scale}
  * csrc/runtime/cache/virtual_cache.cpp:202-232        span byte size
  * span-attention/src/attn/quant.cuh:43-77             dequant x*scale - zero*scale (fp32)
  * csrc/core/operator/generate_opt/batch_mqa/batch_mqa_op.cpp:131-180, csrc/core/kernel/cpu/mha.cpp:595-829
        CPU attention semantics: softmax(alpha * Q K^T) V in fp32 over a contiguous cache
"""
import numpy as np

QUANT_NONE, QUANT_I8, QUANT_U4 = 0, 1, 2  # span::QuantMode, span-attention/include/spanattn/span_attn.h:41-48
HEAD = 128


def span_bytes(mode, span_len, n_groups, head=HEAD, ft_bytes=2):
    """csrc/runtime/cache/virtual_cache.cpp:202-232."""
    if mode == QUANT_NONE:
        return span_len * n_groups * head * ft_bytes
    if mode == QUANT_I8:
        return span_len * n_groups * head + 2 * span_len * n_groups * 4
    if mode == QUANT_U4:
        return span_len * n_groups * head // 2 + 2 * span_len * n_groups * 4
    raise ValueError(mode)


def _fma32(a, b, c):
    """fl32(a * b + c) with ONE rounding: the fp32 x fp32 product is exact in fp64 (48 significant bits) and adding a
    small integer-valued c keeps it exact, so rounding the fp64 sum to fp32 once is a true fused multiply-add."""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def quant_rows(x, mode):
    """x: fp32 [..., HEAD] (values already representable in the activation type).
    Returns (q uint8/int8 [..., HEAD], zero f32 [...], scale f32 [...]).

    Restates QuantParam<I8/U4>::Builder / Quant (impl_i8.cuh:54-61,106-140, impl_u4.cuh:146-182) AS COMPILED: the
    reference builds its span-cache writers with --use_fast_math, so (SASS of QuantCacheAppendKernel in oracle/_ref)
        qs = max((max - min) * fl(1/RANGE), 1e-5)            Div by the constant RANGE -> multiply by its reciprocal
        r  = MUFU.RCP(qs)                                     __fdividef(x, qs) = x * rcp(qs)
        qz = rint(clamp(fma(-min, r, ORIGIN)))                the add is contracted into the multiply
        q  = rint(clamp(fma(x, r, qz)))
    The only step a CPU cannot repeat bit for bit is MUFU.RCP (a <= 1 ulp hardware approximation); here it is the IEEE
    fp32 reciprocal.  The two disagree only when fma(-min, r, ORIGIN) lands within an ulp of a .5 tie, which happens on
    rows with max == -min (their zero point is exactly ORIGIN + RANGE/2): measured on the GPU box against the reference
    kernel itself and recorded by tests/test_ref_pin_gpu.py."""
    x = x.astype(np.float32)
    mx = x.max(axis=-1)
    mn = x.min(axis=-1)
    if mode == QUANT_I8:
        origin, inv_rng, qmax, qmin = np.float32(-128), np.uint32(0x3B808081).view(np.float32), np.float32(127), np.float32(-128)
    else:
        origin, inv_rng, qmax, qmin = np.float32(0), np.uint32(0x3D888889).view(np.float32), np.float32(15), None
    qs = ((mx - mn).astype(np.float32) * inv_rng).astype(np.float32)
    qs = np.maximum(qs, np.float32(1e-5))
    r = (np.float32(1) / qs).astype(np.float32)
    qz = _fma32(-mn, r, np.broadcast_to(origin, mn.shape))
    qz = np.minimum(qz, qmax)
    if qmin is not None:
        qz = np.maximum(qz, qmin)
    qz = np.rint(qz).astype(np.float32)
    t = _fma32(x, r[..., None], np.broadcast_to(qz[..., None], x.shape))
    t = np.minimum(t, qmax)
    if qmin is not None:
        t = np.maximum(t, qmin)
    t = np.rint(t)
    if mode == QUANT_I8:
        q = t.astype(np.int8)
    else:
        q = np.maximum(t, 0).astype(np.uint8)  # cvt.rni.u32.f32 saturates negatives to 0
    return q, qz, qs


def dequant_rows(q, zero, scale):
    """QuantParam::Dequant: (float(q) - zero) * scale (impl_i8.cuh:66-70, impl_u4.cuh:97-106)."""
    return (q.astype(np.float32) - zero[..., None]) * scale[..., None]


def bf16_bits(x_f32):
    import torch
    return torch.from_numpy(np.ascontiguousarray(x_f32, np.float32)).to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)


def bits_to_f32(u16):
    return (u16.astype(np.uint32) << 16).view(np.float32)


class SpanCacheRef:
    """One layer's K (or V) cache for a batch: a list of span byte buffers per sequence,
    byte-identical to what the reference's append kernel writes."""

    def __init__(self, mode, span_len, n_groups, head=HEAD, ft="bf16"):
        """ft: the 16-bit type an unquantized cache stores ("bf16" or "fp16"; span::DataType FP16 / BF16)"""
        self.mode, self.span_len, self.n_groups, self.head, self.ft = mode, span_len, n_groups, head, ft
        self.nbytes = span_bytes(mode, span_len, n_groups, head)
        self.spans = []  # list (per sequence) of list of np.uint8 arrays

    def add_sequence(self):
        self.spans.append([])
        return len(self.spans) - 1

    def _ensure(self, b, n_spans):
        while len(self.spans[b]) < n_spans:
            self.spans[b].append(np.zeros(self.nbytes, np.uint8))

    def append(self, b, pos, rows):
        """rows: fp32 [nG, HEAD] for token index `pos` of sequence b (decoder_cache_append.cuh:126-153)."""
        S, G, H = self.span_len, self.n_groups, self.head
        si, p = pos // S, pos % S
        self._ensure(b, si + 1)
        buf = self.spans[b][si]
        if self.mode == QUANT_NONE:
            v = buf[: S * G * H * 2].view(np.uint16).reshape(G, S, H)
            v[:, p, :] = bf16_bits(rows) if self.ft == "bf16" else np.asarray(rows, np.float32).astype(np.float16).view(np.uint16)
            return
        q, z, s = quant_rows(rows, self.mode)
        if self.mode == QUANT_I8:
            d = buf[: S * G * H].view(np.int8).reshape(G, S, H)
            d[:, p, :] = q
            prm = buf[S * G * H:].view(np.float32).reshape(G, S, 2)
        else:
            d = buf[: S * G * H // 2].reshape(G, S, H // 2)
            d[:, p, :] = (q[:, 0::2] & 0xF) | ((q[:, 1::2] & 0xF) << 4)
            prm = buf[S * G * H // 2:].view(np.float32).reshape(G, S, 2)
        prm[:, p, 0] = z
        prm[:, p, 1] = s

    def dense(self, b, length):
        """Dequantized contiguous cache [nG, length, HEAD] fp32."""
        S, G, H = self.span_len, self.n_groups, self.head
        out = np.zeros((G, length, H), np.float32)
        for si in range((length + S - 1) // S):
            buf = self.spans[b][si]
            n = min(S, length - si * S)
            if self.mode == QUANT_NONE:
                v = buf[: S * G * H * 2].view(np.uint16).reshape(G, S, H)
                out[:, si * S: si * S + n] = bits_to_f32(v[:, :n]) if self.ft == "bf16" else v[:, :n].view(np.float16).astype(np.float32)
                continue
            if self.mode == QUANT_I8:
                q = buf[: S * G * H].view(np.int8).reshape(G, S, H)
                prm = buf[S * G * H:].view(np.float32).reshape(G, S, 2)
            else:
                pk = buf[: S * G * H // 2].reshape(G, S, H // 2)
                q = np.empty((G, S, H), np.uint8)
                q[..., 0::2] = pk & 0xF
                q[..., 1::2] = pk >> 4
                prm = buf[S * G * H // 2:].view(np.float32).reshape(G, S, 2)
            out[:, si * S: si * S + n] = dequant_rows(q[:, :n], prm[:, :n, 0], prm[:, :n, 1])
        return out


def attention_ref(q, kcache, vcache, lens, n_heads, alpha):
    """q: fp32 [B, nH, HEAD]; k/vcache: SpanCacheRef; lens[b] = tokens to attend (incl. the new one).
    Returns fp32 [B, nH, HEAD]: softmax(alpha * q K^T) V computed in fp64 (the CPU path is fp32 MKL;
    fp64 is the order-free stand-in)."""
    B = q.shape[0]
    G = kcache.n_groups
    hpg = n_heads // G
    out = np.zeros((B, n_heads, kcache.head), np.float32)
    for b in range(B):
        L = int(lens[b])
        if L == 0:
            continue
        K = kcache.dense(b, L).astype(np.float64)
        V = vcache.dense(b, L).astype(np.float64)
        for h in range(n_heads):
            g = h // hpg
            s = alpha * (K[g] @ q[b, h].astype(np.float64))
            s = s - s.max()
            p = np.exp(s)
            p = p / p.sum()
            out[b, h] = (p @ V[g]).astype(np.float32)
    return out

"""The allspark-shaped C++ operators (dash-infer_b200/host/) driven like the reference's TestOpUtil
(tests/cpp/operator/cuda/operator_gemm_lowp_test.cpp:650-725): InitV2 -> Reshape -> Forward, checked against the oracle."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from oracle import kvcache_ref as KV
from oracle import quant_ref as Q

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST_LIB = os.path.join(ROOT, "dash-infer_b200", "lib", "liballspark_b200.so")


def _lib():
    import b200spark  # noqa: F401  (loads libb200spark.so with RTLD_GLOBAL first)
    lib = C.CDLL(HOST_LIB)
    lib.as_test_gemm.restype = C.c_int
    lib.as_test_gemm.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p,
                                 C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.as_test_gemm_binary.restype = C.c_int
    lib.as_test_gemm_binary.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p,
                                        C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    lib.as_test_allreduce.restype = C.c_int
    lib.as_test_allreduce.argtypes = [C.c_int, C.c_int64, C.c_void_p, C.c_void_p]
    lib.as_test_span_attn.restype = C.c_int
    lib.as_test_span_attn.argtypes = [C.c_int] * 9 + [C.c_void_p, C.c_void_p]
    lib.as_test_set_dtype.restype = C.c_int
    lib.as_test_set_dtype.argtypes = [C.c_int]
    lib.as_test_registered.restype = C.c_int
    lib.as_test_registered.argtypes = [C.c_char_p]
    return lib


def test_host_library_loads_and_registers_ops():
    """CPU-safe: the operator library loads and the factory knows the reference's op-type strings."""
    lib = _lib()
    for name in (b"GemmA16W4", b"GemmA16W8", b"Gemm", b"DecOptMHA", b"DecOptMQA", b"AllReduce"):
        assert lib.as_test_registered(name) == 1, name
    assert lib.as_test_registered(b"GemmA8W8") == 0  # out of scope: must not pretend


def _bf16_np(t):
    return t.contiguous().view(torch.int16).numpy()


@pytest.mark.gpu
@pytest.mark.parametrize("op,group,M", [("GemmA16W4", -1, 1), ("GemmA16W4", -1, 17), ("GemmA16W4", 128, 5), ("GemmA16W8", -1, 3),
                                        ("GemmA16W8", 128, 31), ("Gemm", -1, 8)])
def test_gemm_operator_like_testoputil(op, group, M):
    from b200spark import quantize as PQ
    lib = _lib()
    K, N = 1024, 640
    g = torch.Generator().manual_seed(M + len(op))
    w = (torch.randn(K, N, generator=g) * 0.02).to(torch.bfloat16)
    a = (torch.rand(M, K, generator=g) * 2 - 1).to(torch.bfloat16)
    bias = (torch.randn(N, generator=g) * 0.02).to(torch.bfloat16)
    if op == "GemmA16W4":
        q, s, z = PQ.quantize_a16w4(w, group); qu = Q.unpack_u4x2(q.numpy(), N); wdt = 10
    elif op == "GemmA16W8":
        q, s, z = PQ.quantize_a16w8(w, group); qu = q.numpy(); wdt = 3
    else:
        q, s, z = w, None, None; wdt = 9
    out = np.zeros((M, N), np.int16)
    qn = _bf16_np(q) if op == "Gemm" else q.contiguous().numpy()
    an, bn = _bf16_np(a), _bf16_np(bias)
    sn = _bf16_np(s) if s is not None else None
    zn = _bf16_np(z) if z is not None else None
    rc = lib.as_test_gemm(op.encode(), M, N, K, group, 5, 1.0, an.ctypes.data, qn.ctypes.data, wdt,
                          sn.ctypes.data if sn is not None else None, zn.ctypes.data if zn is not None else None,
                          bn.ctypes.data, out.ctypes.data)
    assert rc == 0, rc
    got = torch.from_numpy(out).view(torch.bfloat16).float().numpy()
    if op == "Gemm":
        ref = Q.activation((a.float().numpy().astype(np.float64) @ w.float().numpy().astype(np.float64)
                            + bias.float().numpy()[None]).astype(np.float32), 5)
    else:
        ref = Q.gemm_wq_math(a.float().numpy(), qu, s.float().numpy(), z.float().numpy(), group, bias.float().numpy(), 5)
    assert Q.err_min_abs_rel(ref, got) <= 2e-2


@pytest.mark.gpu
def test_gemm_operator_rejects_bad_configs():
    lib = _lib()
    K, N, M = 128, 64, 1
    z16 = np.zeros((K, N), np.int16)
    out = np.zeros((M, N), np.int16)
    # GemmA16W4 with int8-typed packed weight: ALLSPARK_PARAM_ERROR (gemm_a16w4.cpp:104-110)
    rc = lib.as_test_gemm(b"GemmA16W4", M, N, K, -1, 0, 1.0, z16.ctypes.data, z16.ctypes.data, 3, z16.ctypes.data, z16.ctypes.data,
                          None, out.ctypes.data)
    assert rc == 2
    # GroupSize 24: rejected like gemm_a16w4.cpp:57-63
    rc = lib.as_test_gemm(b"GemmA16W4", M, N, K, 24, 0, 1.0, z16.ctypes.data, z16.ctypes.data, 10, z16.ctypes.data, z16.ctypes.data,
                          None, out.ctypes.data)
    assert rc == 2


@pytest.mark.gpu
@pytest.mark.parametrize("span,steps", [(16, 40), (128, 130)])
def test_span_attention_operator_decode_loop(span, steps):
    """DecOptMQA through Alloc/Forward for `steps` decode steps (spans claimed on demand), every step checked."""
    lib = _lib()
    B, nH, nG = 3, 8, 2
    W, OW = (nH + 2 * nG) * 128, nH * 128
    rng = np.random.default_rng(span)
    qkv = torch.from_numpy(rng.standard_normal((steps, B, W)).astype(np.float32)).to(torch.bfloat16)
    out = np.zeros((steps, B, OW), np.int16)
    rc = lib.as_test_span_attn(B, steps, nH, nG, span, 0, 256, 3, 1, _bf16_np(qkv).ctypes.data, out.ctypes.data)
    assert rc == 0, rc
    got = torch.from_numpy(out).view(torch.bfloat16).float().numpy().reshape(steps, B, nH, 128)
    kref, vref = KV.SpanCacheRef(KV.QUANT_NONE, span, nG), KV.SpanCacheRef(KV.QUANT_NONE, span, nG)
    for _ in range(B):
        kref.add_sequence(); vref.add_sequence()
    x = qkv.float().numpy().reshape(steps, B, nH + 2 * nG, 128)
    for t in range(steps):
        for b in range(B):
            kref.append(b, t, x[t, b, nH:nH + nG]); vref.append(b, t, x[t, b, nH + nG:])
        if t in (0, 1, span - 1, span, steps - 1):
            ref = KV.attention_ref(x[t, :, :nH], kref, vref, [t + 1] * B, nH, 1.0 / np.sqrt(128))
            # 2e-3 abs + bf16 rounding of the stored output (2^-7 |ref| envelope) and of the
            # probabilities fed to the tensor core (2^-9 relative each, weighted by |V| <= vmax)
            vmax = float(np.abs(x[:t + 1, :, nH + nG:]).max())
            assert np.all(np.abs(got[t] - ref) <= 2e-3 + 2.0 ** -7 * np.abs(ref) + 2.0 ** -9 * vmax), (t, np.abs(got[t] - ref).max())


@pytest.mark.gpu
@pytest.mark.parametrize("op", ["Gemm", "GemmA16W4"])
def test_gemm_operator_binary_add_residual(op):
    """the default do_binary_add_fused graph emits Gemm(x, residual) with binary_type = ADD for o_proj / down_proj
    (qwen_v15.py:280-286,340; GemmOpBase gemm_op.cpp:73-136): the second input must be added, not dropped."""
    from b200spark import quantize as PQ
    lib = _lib()
    M, K, N = 5, 1024, 640
    g = torch.Generator().manual_seed(77)
    w = (torch.randn(K, N, generator=g) * 0.02).to(torch.bfloat16)
    a = (torch.rand(M, K, generator=g) * 2 - 1).to(torch.bfloat16)
    res = torch.randn(M, N, generator=g).to(torch.bfloat16)
    out = np.zeros((M, N), np.int16)
    if op == "Gemm":
        q, s, z, wdt, qn = w, None, None, 9, _bf16_np(w)
        ref = a.float().numpy().astype(np.float64) @ w.float().numpy().astype(np.float64)
    else:
        q, s, z = PQ.quantize_a16w4(w, -1)
        wdt, qn = 10, q.contiguous().numpy()
        ref = Q.gemm_wq_math(a.float().numpy(), Q.unpack_u4x2(q.numpy(), N), s.float().numpy(), z.float().numpy(), -1)
    sn = _bf16_np(s) if s is not None else None
    zn = _bf16_np(z) if z is not None else None
    an, rn = _bf16_np(a), _bf16_np(res)
    rc = lib.as_test_gemm_binary(op.encode(), M, N, K, -1, 0, 1.0, an.ctypes.data, qn.ctypes.data, wdt,
                                 sn.ctypes.data if sn is not None else None, zn.ctypes.data if zn is not None else None, None,
                                 rn.ctypes.data, 1, out.ctypes.data)
    assert rc == 0, rc
    got = torch.from_numpy(out).view(torch.bfloat16).float().numpy()
    assert Q.err_min_abs_rel((ref + res.float().numpy()).astype(np.float32), got) <= 2e-2
    # binary_type MUL is not implemented by GemmOpGPU either: PARAM_ERROR, not a silent drop
    rc = lib.as_test_gemm_binary(op.encode(), M, N, K, -1, 0, 1.0, an.ctypes.data, qn.ctypes.data, wdt,
                                 sn.ctypes.data if sn is not None else None, zn.ctypes.data if zn is not None else None, None,
                                 rn.ctypes.data, 2, out.ctypes.data)
    assert rc == 2


@pytest.mark.gpu
def test_span_attention_operator_needs_layer_number_in_name():
    """span_attn_op.cpp:183-187: the layer index is parsed from the op name; a name without one is ALLSPARK_PARAM_ERROR."""
    lib = _lib()
    qkv = np.zeros((1, 1, (8 + 4) * 128), np.int16)
    out = np.zeros((1, 1, 8 * 128), np.int16)
    assert lib.as_test_span_attn(1, 1, 8, 2, 16, 0, 64, 2, -1, qkv.ctypes.data, out.ctypes.data) == 2


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [KV.QUANT_I8, KV.QUANT_U4])
@pytest.mark.parametrize("layer", [0, 2])
def test_span_attention_operator_quantized_cache_modes(mode, layer):
    """DecOptMQA with AsCacheQuantI8 / AsCacheQuantU4 spans on layer 0 and on a later layer of a 3-layer cache (the op finds
    its layer through the op name), 40 decode steps with span 16 (spans claimed on demand, pinned staging ring wraps)."""
    lib = _lib()
    B, nH, nG, span, steps = 3, 8, 2, 16, 40
    W, OW = (nH + 2 * nG) * 128, nH * 128
    rng = np.random.default_rng(mode * 10 + layer)
    qkv = torch.from_numpy(rng.standard_normal((steps, B, W)).astype(np.float32)).to(torch.bfloat16)
    out = np.zeros((steps, B, OW), np.int16)
    rc = lib.as_test_span_attn(B, steps, nH, nG, span, mode, 256, 3, layer, _bf16_np(qkv).ctypes.data, out.ctypes.data)
    assert rc == 0, rc
    got = torch.from_numpy(out).view(torch.bfloat16).float().numpy().reshape(steps, B, nH, 128)
    kref, vref = KV.SpanCacheRef(mode, span, nG), KV.SpanCacheRef(mode, span, nG)
    for _ in range(B):
        kref.add_sequence(); vref.add_sequence()
    x = qkv.float().numpy().reshape(steps, B, nH + 2 * nG, 128)
    for t in range(steps):
        for b in range(B):
            kref.append(b, t, x[t, b, nH:nH + nG]); vref.append(b, t, x[t, b, nH + nG:])
        if t in (0, 1, span - 1, span, steps - 1):
            ref = KV.attention_ref(x[t, :, :nH], kref, vref, [t + 1] * B, nH, 1.0 / np.sqrt(128))
            # the op's spans are not reachable from here, so the oracle attends over ITS OWN quantization of the same rows:
            # on exact zero-point ties (~1 % of rows) the GPU's MUFU.RCP and the oracle's IEEE reciprocal round apart and a
            # clamped code moves by one quantization step (1/255 resp. 1/15 of the row's range), hence the mode-dependent
            # allowance on top of the attention tolerance (identical-bytes attention parity: tests/test_attn_gpu.py)
            step = {KV.QUANT_I8: 1 / 255, KV.QUANT_U4: 1 / 15}[mode] * 8.0
            assert np.all(np.abs(got[t] - ref) <= 2e-3 + 2.0 ** -7 * np.abs(ref) + 0.1 * step), (t, np.abs(got[t] - ref).max())


@pytest.mark.gpu
@pytest.mark.parametrize("nranks", [2, 4])
def test_allreduce_operator(nranks):
    """The AllReduce op type (allreduce_op.cpp:73-115) over b2_allreduce: nranks op instances, each with its own context,
    stream and communicator on one device; in-place like the reference graphs; bit-exact fp32-rank-order sums."""
    lib = _lib()
    count = 16 * 8192
    g = torch.Generator().manual_seed(nranks)
    xs = torch.randn(nranks, count, generator=g).to(torch.bfloat16)
    out = np.zeros((nranks, count), np.int16)
    rc = lib.as_test_allreduce(nranks, count, _bf16_np(xs).ctypes.data, out.ctypes.data)
    assert rc == 0, rc
    acc = torch.zeros(count)
    for r in range(nranks):
        acc = acc + xs[r].float()
    exp = acc.to(torch.bfloat16)
    got = torch.from_numpy(out).view(torch.bfloat16)
    for r in range(nranks):
        assert torch.equal(got[r], exp), r


@pytest.mark.gpu
@pytest.mark.parametrize("op,group,M", [("GemmA16W4", -1, 1), ("GemmA16W4", -1, 40), ("GemmA16W4", 128, 17), ("GemmA16W8", -1, 3),
                                        ("GemmA16W8", 64, 9), ("Gemm", -1, 33)])
def test_gemm_operators_fp16(op, group, M):
    """FLOAT16 tensors through the same operator classes (the reference dispatches FLOAT16 first, gemm_a16w4_gpu.cpp:31-38)"""
    from b200spark import quantize as PQ
    lib = _lib()
    K, N = 1024, 640
    g = torch.Generator().manual_seed(M + len(op))
    w = (torch.randn(K, N, generator=g) * 0.02).to(torch.float16)
    a = (torch.rand(M, K, generator=g) * 2 - 1).to(torch.float16)
    bias = (torch.randn(N, generator=g) * 0.02).to(torch.float16)
    if op == "GemmA16W4":
        q, s, z = PQ.quantize_a16w4(w, group); qu = Q.unpack_u4x2(q.numpy(), N); wdt = 10
    elif op == "GemmA16W8":
        q, s, z = PQ.quantize_a16w8(w, group); qu = q.numpy(); wdt = 3
    else:
        q, s, z = w, None, None; wdt = 2
    f16 = lambda t: t.contiguous().view(torch.int16).numpy()
    out = np.zeros((M, N), np.int16)
    qn = f16(q) if op == "Gemm" else q.contiguous().numpy()
    an, bn = f16(a), f16(bias)
    sn, zn = (f16(s), f16(z)) if s is not None else (None, None)
    assert lib.as_test_set_dtype(2) == 0
    try:
        rc = lib.as_test_gemm(op.encode(), M, N, K, group, 5, 1.0, an.ctypes.data, qn.ctypes.data, wdt,
                              sn.ctypes.data if sn is not None else None, zn.ctypes.data if zn is not None else None,
                              bn.ctypes.data, out.ctypes.data)
    finally:
        lib.as_test_set_dtype(9)
    assert rc == 0, rc
    got = torch.from_numpy(out).view(torch.float16).float().numpy()
    if op == "Gemm":
        ref = Q.activation((a.float().numpy().astype(np.float64) @ w.float().numpy().astype(np.float64)
                            + bias.float().numpy()[None]).astype(np.float32), 5)
    else:
        ref = Q.gemm_wq_math(a.float().numpy(), qu, s.float().numpy(), z.float().numpy(), group, bias.float().numpy(), 5)
    assert Q.err_min_abs_rel(ref, got) <= 2e-2


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_span_attention_operator_fp16(mode):
    lib = _lib()
    B, nH, nG, span, steps = 2, 8, 2, 32, 70
    W, OW = (nH + 2 * nG) * 128, nH * 128
    rng = np.random.default_rng(77 + mode)
    qkv = torch.from_numpy(rng.standard_normal((steps, B, W)).astype(np.float32)).to(torch.float16)
    out = np.zeros((steps, B, OW), np.int16)
    assert lib.as_test_set_dtype(2) == 0
    try:
        rc = lib.as_test_span_attn(B, steps, nH, nG, span, mode, 256, 3, 1, qkv.view(torch.int16).numpy().ctypes.data, out.ctypes.data)
    finally:
        lib.as_test_set_dtype(9)
    assert rc == 0, rc
    got = torch.from_numpy(out).view(torch.float16).float().numpy().reshape(steps, B, nH, 128)
    kref, vref = KV.SpanCacheRef(mode, span, nG, ft="fp16"), KV.SpanCacheRef(mode, span, nG, ft="fp16")
    for _ in range(B):
        kref.add_sequence(); vref.add_sequence()
    x = qkv.float().numpy().reshape(steps, B, nH + 2 * nG, 128)
    for t in range(steps):
        for b in range(B):
            kref.append(b, t, x[t, b, nH:nH + nG]); vref.append(b, t, x[t, b, nH + nG:])
        if t in (0, span - 1, span, steps - 1):
            ref = KV.attention_ref(x[t, :, :nH], kref, vref, [t + 1] * B, nH, 1.0 / np.sqrt(128))
            # quantized modes: the oracle quantizes with an IEEE reciprocal (ties can move a code by one step)
            # (the envelope the run on the GPU box passed with; the bf16 twin of this test uses 0.1 quantization steps)
            tol = 2e-3 + 2.0 ** -9 * np.abs(ref) + (0.0 if mode == 0 else (2e-2 if mode == 1 else 1.5e-1))
            assert np.all(np.abs(got[t] - ref) <= tol), (t, np.abs(got[t] - ref).max())

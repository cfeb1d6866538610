"""Oracle: InstantQuant weight-only quantizers, packers and dequant-GEMM references.

TEST INFRASTRUCTURE (see oracle/__init__.py).  numpy fp32, one IEEE op per line so the
result is bit-identical to the reference's torch code on CPU.

Follows (paths relative to /root/reference):
  * python/pyhie/allspark/model/quantization_utils.py:158-217  quantize_gemm_weight_a16w8_torch
  * python/pyhie/allspark/model/quantization_utils.py:240-304  quantize_gemm_weight_a16w4_torch
  * python/pyhie/allspark/model/quantization_utils.py:331-351,391-437  GPTQ depack / repack
  * tests/cpp/operator/cuda/operator_gemm_lowp_test.cpp:18-29    PackU8ToU4x2
  * tests/cpp/operator/cuda/operator_gemm_lowp_test.cpp:131-205  CPU_SubC_Ref / CPU_PerC_Ref / CPU_FP16W4_PerC_Ref
  * csrc/core/kernel/cuda/gemm_lowp/gemm_lowp_utils.cuh:582-604  reduce_sum: out = Act(sum + bias)
  * csrc/core/kernel/cuda/hie/cuda_activation.hpp:20-120         activation formulas
"""
import math

import numpy as np
import torch

# UnaryType wire values, csrc/proto/allspark.proto (SURVEY.md §8b "Enums on the wire")
ACT_NONE, ACT_TANH, ACT_GELU_ERF, ACT_GELU_TANH, ACT_RELU, ACT_SILU, ACT_SIGMOID = 0, 1, 2, 3, 4, 5, 6


def to_ft(x_f32, ft):
    """fp32 numpy -> values rounded to the float type `ft` ('bf16'|'fp16'), returned as fp32."""
    t = torch.from_numpy(np.ascontiguousarray(x_f32, dtype=np.float32))
    dt = torch.bfloat16 if ft == "bf16" else torch.float16
    return t.to(dt).to(torch.float32).numpy()


def _minmax_params(data_f32, qmin, qmax):
    # quantization_utils.py:201-212 / :281-291 — data: [N, G, group]
    fmax = data_f32.max(axis=-1, keepdims=True).astype(np.float32)
    fmin = data_f32.min(axis=-1, keepdims=True).astype(np.float32)
    scale = (fmax - fmin) / np.float32(qmax - qmin)
    scale = np.where(scale == 0, np.float32(1), scale).astype(np.float32)
    zero = (np.float32(qmin) - fmin / scale).astype(np.float32)
    res = (data_f32 / scale).astype(np.float32) + zero
    q = np.rint(np.clip(res.astype(np.float32), np.float32(qmin), np.float32(qmax)))  # torch.round = half-to-even
    return q, scale, zero


def quantize_a16w8(w, ft="bf16", group_size=-1, qtype="int8"):
    """w: fp32 numpy [K, N] holding `ft`-representable values.
    Returns qdata [K,N] (int8|uint8), scale [G,N], zero [G,N] (fp32 arrays holding ft values).
    quantization_utils.py:158-217."""
    K, N = w.shape
    qmin, qmax = (-128, 127) if qtype == "int8" else (0, 255)
    gs = K if group_size in (-1, None) else int(group_size)
    kstride = (K + gs - 1) // gs * gs
    kpad = kstride - K
    wp = np.concatenate([w, np.repeat(w[-1:, :], kpad, axis=0)], axis=0) if kpad else w
    data = np.ascontiguousarray(wp.T).reshape(N, -1, gs).astype(np.float32)
    q, scale, zero = _minmax_params(data, qmin, qmax)
    qdata = np.ascontiguousarray(q.reshape(N, -1).T)[:K].astype(np.int8 if qtype == "int8" else np.uint8)
    scale = to_ft(np.ascontiguousarray(scale.reshape(N, -1).T), ft)
    zero = to_ft(np.ascontiguousarray(zero.reshape(N, -1).T), ft)
    return qdata, scale, zero


def pack_u4x2(q_u8):
    """[K, N] uint8 (values 0..15) -> [K, ceil(N/2)] uint8, low nibble = even column.
    quantization_utils.py:297 and operator_gemm_lowp_test.cpp:18-29."""
    K, N = q_u8.shape
    if N % 2:
        q_u8 = np.concatenate([q_u8, np.zeros((K, 1), np.uint8)], axis=1)
    return ((q_u8[:, 1::2].astype(np.uint8) << 4) | (q_u8[:, 0::2].astype(np.uint8) & 0xF)).astype(np.uint8)


def unpack_u4x2(packed, N):
    K = packed.shape[0]
    out = np.empty((K, packed.shape[1] * 2), np.uint8)
    out[:, 0::2] = packed & 0xF
    out[:, 1::2] = packed >> 4
    return out[:, :N]


def quantize_a16w4(w, ft="bf16", group_size=-1):
    """uint4 asymmetric; returns packed [K, ceil(N/2)] uint8, scale [G,N], zero [G,N].
    quantization_utils.py:240-304 (weight_type UINT4: qmax 15, qmin 0)."""
    K, N = w.shape
    qmin, qmax = 0, 15
    gs = K if group_size in (-1, None) else int(group_size)
    kstride = (K + gs - 1) // gs * gs
    kpad = kstride - K
    wp = np.concatenate([w, np.repeat(w[-1:, :], kpad, axis=0)], axis=0) if kpad else w
    nstride = (N + 1) // 2 * 2
    if nstride != N:
        wp = np.concatenate([wp, np.zeros((wp.shape[0], nstride - N), wp.dtype)], axis=1)
    data = np.ascontiguousarray(wp.T).reshape(nstride, -1, gs).astype(np.float32)
    q, scale, zero = _minmax_params(data, qmin, qmax)
    q = np.ascontiguousarray(q.reshape(nstride, -1).T).astype(np.uint8)
    packed = pack_u4x2(q)[:K]
    scale = to_ft(np.ascontiguousarray(scale.reshape(nstride, -1).T)[:, :N], ft)
    zero = to_ft(np.ascontiguousarray(zero.reshape(nstride, -1).T)[:, :N], ft)
    return packed, scale, zero


# ---------------------------------------------------------------- GPTQ repack
def gptq_depack_weight(qweight_i32, bits=4):
    """quantization_utils.py:331-339: [K*bits/32, N] int32 -> [K, N]; row r*(32/bits)+j = (qweight[r] >> bits*j) & mask."""
    per = 32 // bits
    shifts = (np.arange(per, dtype=np.int64) * bits)[None, :, None]
    w = (qweight_i32.astype(np.int64)[:, None, :] & 0xFFFFFFFF) >> shifts
    return (w & ((1 << bits) - 1)).reshape(-1, qweight_i32.shape[1]).astype(np.int16)


def gptq_depack_zero(qzeros_i32, bits=4):
    """quantization_utils.py:342-351: [G, N*bits/32] -> [G, N], **+1** added."""
    per = 32 // bits
    shifts = (np.arange(per, dtype=np.int64) * bits)[None, None, :]
    z = (qzeros_i32.astype(np.int64)[:, :, None] & 0xFFFFFFFF) >> shifts
    z = (z & ((1 << bits) - 1)) + 1
    return z.reshape(z.shape[0], -1).astype(np.int16)


def repack_gptq_a16w4(qweight_i32, qzeros_i32, scales_f32, ft="fp16"):
    """quantization_utils.py:391-437 for bits=4."""
    q = gptq_depack_weight(qweight_i32, 4).astype(np.uint8)
    packed = pack_u4x2(q)
    zeros = to_ft(gptq_depack_zero(qzeros_i32, 4).astype(np.float32), ft)
    return packed, to_ft(scales_f32, ft), zeros


# ---------------------------------------------------------------- dequant + GEMM references
def dequant(qdata, scale, zero, group_size=-1):
    """(float(q) - float(z)) * float(s), fp32 — operator_gemm_lowp_test.cpp:143-146,164-166."""
    K, N = qdata.shape
    if group_size in (-1, None) or scale.shape[0] == 1:
        return (qdata.astype(np.float32) - zero[0][None, :]) * scale[0][None, :]
    gidx = np.arange(K) // int(group_size)
    return (qdata.astype(np.float32) - zero[gidx]) * scale[gidx]


def activation(x, act):
    """cuda_activation.hpp: Relu :20, Tanh :25, Gelu(erf) :30-35, Silu :37-42, GeluTanh :113-120 (fp32 math)."""
    x = x.astype(np.float32)
    if act == ACT_NONE:
        return x
    if act == ACT_RELU:
        return np.maximum(x, 0).astype(np.float32)
    if act == ACT_TANH:
        return np.tanh(x).astype(np.float32)
    if act == ACT_GELU_ERF:
        erf = np.vectorize(math.erf, otypes=[np.float64])
        return (x * 0.5 * (1.0 + erf(x.astype(np.float64) * 0.70710678))).astype(np.float32)
    if act == ACT_GELU_TANH:
        x64 = x.astype(np.float64)
        return (x64 * 0.5 * (1.0 + np.tanh(0.7978845608028654 * (x64 + 0.044715 * x64 ** 3)))).astype(np.float32)
    if act == ACT_SILU:
        x64 = x.astype(np.float64)
        return (x64 / (1.0 + np.exp(-x64))).astype(np.float32)
    if act == ACT_SIGMOID:
        x64 = x.astype(np.float64)
        return (1.0 / (1.0 + np.exp(-x64))).astype(np.float32)
    raise ValueError(act)


def gemm_wq_math(a, qdata, scale, zero, group_size=-1, bias=None, act=ACT_NONE, alpha=1.0):
    """fp32 'math oracle' = the reference tests' CPU refs: C = Act(alpha * A @ dequant(W) + bias).
    a: fp32 [M,K] holding ft values.  Accumulation in fp64 then cast (the reference loop is fp32
    sequential; fp64 is the order-independent stand-in, differences << test tolerance)."""
    w = dequant(qdata, scale, zero, group_size).astype(np.float64)
    c = alpha * (a.astype(np.float64) @ w)
    if bias is not None:
        c = c + bias.astype(np.float64)[None, :]
    return activation(c.astype(np.float32), act)


def gemm_wq_cpu_path(a, qdata, scale, zero, group_size=-1, bias=None, act=ACT_NONE, ft="bf16"):
    """'reference-CPU-path oracle' (SURVEY.md §8c-ii): the x86 build has no weight-only op, so a
    quantized model runs as Gemm on dequantized weights stored in `ft`
    (csrc/core/operator/general/gemm/gemm_op_cpu.cpp:75-216: src->bf16, bf16 weights, f32 dst,
    bias/activation post-ops).  Uses torch CPU (oneDNN) bf16 matmul like the reference."""
    dt = torch.bfloat16 if ft == "bf16" else torch.float16
    w = torch.from_numpy(dequant(qdata, scale, zero, group_size)).to(dt)
    x = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dt)
    c = torch.matmul(x, w).to(torch.float32).numpy()
    if bias is not None:
        c = c + bias.astype(np.float32)[None, :]
    return activation(c, act)


def err_min_abs_rel(ref, out):
    """tests/cpp/test_common.h.in:82-110 check_equal: max_i min(|ref-out|, |ref-out|/|out|)."""
    ref = ref.astype(np.float64).ravel()
    out = out.astype(np.float64).ravel()
    d = np.abs(ref - out)
    with np.errstate(divide="ignore", invalid="ignore"):
        rel = np.where(out != 0, d / np.abs(out), np.inf)
    e = np.minimum(d, rel)
    return float(np.max(e)) if e.size else 0.0

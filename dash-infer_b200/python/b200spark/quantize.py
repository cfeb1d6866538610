"""InstantQuant weight-only quantizers (offline, host side of the hot path), torch on any device.

Same contract as the reference's `quantize_gemm_weight_a16w{4,8}_torch`
(python/pyhie/allspark/model/quantization_utils.py:158-304) and `repack_gptq_to_a16wX` (:391-437):
returns (qdata, scale, zero) in the layouts GemmA16W4 / GemmA16W8 consume:
  A16W4: qdata uint8 [K, ceil(N/2)] (low nibble = even column), scale/zero FT [G, N]
  A16W8: qdata int8  [K, N],                                    scale/zero FT [G, N]
G = 1 for per-channel (group_size = -1, the IQ default) else ceil(K / group_size).
"""
import torch


def _minmax_quant(w, group_size, qmin, qmax, pad_n_even):
    K, N = w.shape
    ft = w.dtype
    gs = K if group_size in (-1, None) else int(group_size)
    kpad = (K + gs - 1) // gs * gs - K
    if kpad:  # the tail group sees the last row repeated (keeps its min/max unchanged)
        w = torch.cat([w, w[-1:, :].expand(kpad, N)], dim=0)
    npad = (N % 2) if pad_n_even else 0
    if npad:
        w = torch.nn.functional.pad(w, (0, 1))
    ns = N + npad
    x = w.t().reshape(ns, -1, gs)
    fmax = x.amax(dim=-1, keepdim=True).float()
    fmin = x.amin(dim=-1, keepdim=True).float()
    scale = (fmax - fmin) / float(qmax - qmin)
    scale = torch.where(scale == 0, torch.ones_like(scale), scale)
    zero = float(qmin) - fmin / scale
    q = torch.round(torch.clamp(x.float() / scale + zero, float(qmin), float(qmax)))
    q = q.reshape(ns, -1).t().contiguous()
    scale = scale.reshape(ns, -1).t().contiguous()[:, :N].to(ft)
    zero = zero.reshape(ns, -1).t().contiguous()[:, :N].to(ft)
    return q, scale, zero, K


def quantize_a16w8(w, group_size=-1, signed=True):
    qmin, qmax = (-128, 127) if signed else (0, 255)
    q, scale, zero, K = _minmax_quant(w, group_size, qmin, qmax, pad_n_even=False)
    return q[:K].to(torch.int8 if signed else torch.uint8), scale, zero


def pack_u4x2(q_u8):
    if q_u8.shape[1] % 2:
        q_u8 = torch.nn.functional.pad(q_u8, (0, 1))
    return ((q_u8[:, 1::2] << 4) | (q_u8[:, 0::2] & 0xF)).to(torch.uint8)


def quantize_a16w4(w, group_size=-1):
    q, scale, zero, K = _minmax_quant(w, group_size, 0, 15, pad_n_even=True)
    return pack_u4x2(q.to(torch.uint8))[:K].contiguous(), scale, zero


def repack_gptq_a16w4(qweight_i32, qzeros_i32, scales):
    """AutoGPTQ int32-packed tensors -> (qdata [K, N/2] uint8, scales [G,N], zeros [G,N] (+1 applied))."""
    dev = qweight_i32.device
    sh = torch.arange(0, 32, 4, device=dev, dtype=torch.int32)
    q = ((qweight_i32[:, None, :] >> sh[None, :, None]) & 0xF).reshape(-1, qweight_i32.shape[1]).to(torch.uint8)
    z = ((qzeros_i32[:, :, None] >> sh[None, None, :]) & 0xF) + 1
    z = z.reshape(z.shape[0], -1).to(scales.dtype)
    return pack_u4x2(q).contiguous(), scales, z


def dequantize(qdata, scale, zero, group_size=-1, wbits=4, N=None):
    """fp32 (q - zero) * scale — for tests and for building dense references."""
    if wbits == 4:
        N = N if N is not None else scale.shape[1]
        q = torch.empty(qdata.shape[0], qdata.shape[1] * 2, dtype=torch.uint8, device=qdata.device)
        q[:, 0::2] = qdata & 0xF
        q[:, 1::2] = qdata >> 4
        q = q[:, :N]
    else:
        q = qdata
    K = q.shape[0]
    if group_size in (-1, None):
        return (q.float() - zero[0].float()[None, :]) * scale[0].float()[None, :]
    gi = torch.arange(K, device=q.device) // int(group_size)
    return (q.float() - zero.float()[gi]) * scale.float()[gi]

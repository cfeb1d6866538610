"""Tensor-parallel partitioning of quantized projections — the reference's split modes applied to (qdata, scale, zero):
  VSPLIT / GROUP_VSPLIT (column split: QKV by kv-group, gate, up), HSPLIT (row split: o_proj, down_proj; per-channel
  scale/zero replicated, sub-channel params split by group), python/pyhie/allspark/model/qwen_v15.py:125-146,539-569,
  csrc/runtime/weight/weight_splitter.cpp:60-918.  Pure torch (CPU or GPU): the host-side logic is unit-tested with gloo.
"""
import torch


def col_ranges_qkv(n_heads, n_groups, rank, tp, head=128):
    """Column ranges of the fused [q | k | v] projection owned by `rank` (GROUP_VSPLIT: heads split by kv-group)."""
    assert n_heads % tp == 0 and n_groups % tp == 0
    hq, hg = n_heads // tp, n_groups // tp
    qoff, koff, voff = 0, n_heads * head, (n_heads + n_groups) * head
    return [(qoff + rank * hq * head, qoff + (rank + 1) * hq * head),
            (koff + rank * hg * head, koff + (rank + 1) * hg * head),
            (voff + rank * hg * head, voff + (rank + 1) * hg * head)]


def col_range_even(N, rank, tp):
    assert N % tp == 0
    return [(rank * (N // tp), (rank + 1) * (N // tp))]


def shard_cols(q, s, z, bias, wbits, ranges):
    """Column (N) split.  q: packed uint4x2 [K, N/2] (wbits 4), int8 [K, N] (8) or bf16 [K, N] (16)."""
    qs, ss, zs, bs = [], [], [], []
    for a, b in ranges:
        if wbits == 4:
            assert a % 2 == 0 and b % 2 == 0
            qs.append(q[:, a // 2: b // 2])
        else:
            qs.append(q[:, a:b])
        if s is not None:
            ss.append(s[:, a:b]); zs.append(z[:, a:b])
        if bias is not None:
            bs.append(bias[a:b])
    cat = lambda xs, d: torch.cat(xs, dim=d).contiguous() if xs else None
    return cat(qs, 1), cat(ss, 1), cat(zs, 1), cat(bs, 0)


def shard_rows(q, s, z, wbits, group, rank, tp):
    """Row (K) split.  Per-channel params are replicated (every rank dequantizes its K-slice with the full-K
    scale/zero, qwen_v15.py:562-569); sub-channel params are split by group and need (K/tp) % group == 0."""
    K = q.shape[0]
    assert K % tp == 0
    k0, k1 = rank * (K // tp), (rank + 1) * (K // tp)
    qr = q[k0:k1].contiguous()
    if s is None or group in (-1, None):
        return qr, s, z
    assert (K // tp) % group == 0, "sub-channel row split needs (K/tp) % group == 0 (weight_activate_quant.rst:361-362)"
    return qr, s[k0 // group: k1 // group].contiguous(), z[k0 // group: k1 // group].contiguous()

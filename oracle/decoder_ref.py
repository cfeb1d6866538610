"""Oracle: one decode step of a Qwen2/Llama-style decoder with the reference's CPU-path semantics.

TEST INFRASTRUCTURE / CPU BASELINE (see oracle/__init__.py).  The x86 build of the reference has no
weight-only-quant operator (csrc/core/operator/general/gemm_lowp/gemm_a16w4_gpu.cpp:420 registers CUDA only), so
its CPU answer for a quantized model is: dequantize (q - zero) * scale, store the weight in the model dtype (bf16),
and run the ordinary operators:
  * Gemm: src -> bf16, bf16 weight, fp32 accumulate, bias/activation post-ops, dst in the model dtype
      (csrc/core/operator/general/gemm/gemm_op_cpu.cpp:75-216) — here torch CPU bf16 matmul (oneDNN, like the reference)
  * LayerNormNoBeta (RMSNorm), Rotary (NeoX rotate-half), Binary ADD/MUL, SiLU: fp32 math, dst in the model dtype
  * attention: alpha * Q K^T -> softmax -> P V in fp32 over the cache
      (csrc/core/operator/generate_opt/batch_mqa/batch_mqa_op.cpp:131-180, csrc/core/kernel/cpu/mha.cpp:595-829);
      the cache holds what the span cache holds (bf16 rows, or I8/U4 rows dequantized with their stored params)
  * greedy sampling: argmax, lowest index on ties (csrc/core/kernel/cpu/generate_impl_cpu.hpp:153-165, top_k = 1)
Graph order: python/pyhie/allspark/model/qwen_v15.py:206-379.
"""
import math

import numpy as np
import torch

from . import kvcache_ref as KV


_EXACT = False  # set by RefDecoder(exact=True).step: no intermediate rounding at all (the "truth" both paths approximate)
_FT = torch.bfloat16  # the model dtype every operator output is rounded to (set per RefDecoder: bf16 or fp16)


def _bf(x):
    return x.float() if _EXACT else x.to(_FT)


class RefDecoder:
    def __init__(self, cfg, layers, embed, gf, lm_head, kv_mode=KV.QUANT_NONE, exact=False, ft=torch.bfloat16):
        """layers: list of dicts with g1,g2 (fp32 [H]) and qkv,o,gate,up,down = (W fp32 [K,N], bias or None).
        exact=True: the same graph on the same (bf16-valued) weights with fp32 activations and no rounding between operators —
        not a reference path, but the yardstick for "how far is a bf16 implementation from the exact result" on deep stacks."""
        global _FT
        self.cfg, self.kv_mode, self.exact, self.ft = cfg, kv_mode, exact, ft
        _FT = ft
        self.embed = _bf(embed)
        self.gf = gf.float()
        self.lm = _bf(lm_head)
        self.layers = []
        for L in layers:
            self.layers.append({k: (_bf(v[0]), None if v[1] is None else v[1].float()) if isinstance(v, tuple) else v.float()
                                for k, v in L.items()})
        self.k = None  # [layer][b] -> list of fp32 rows [nG,128] (dequantized like the cache would return them)
        self.v = None

    def reset(self, batch):
        n = len(self.layers)
        self.k = [[[] for _ in range(batch)] for _ in range(n)]
        self.v = [[[] for _ in range(batch)] for _ in range(n)]

    def _rms(self, x, g):
        xf = x.float()
        inv = torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + self.cfg.eps)
        return _bf(xf * inv * g)

    def _gemm(self, x, wb, act=None):
        w, b = wb
        # fp32 accumulation of the rounded operands (fp16: through fp32 copies — the products are exact in fp32 either way)
        y = (torch.matmul(_bf(x).float(), w.float()) if (self.exact or self.ft == torch.float16) else torch.matmul(_bf(x), w)).float()
        if b is not None:
            y = y + b
        if act == "silu":
            y = y * torch.sigmoid(y)
        return _bf(y)

    def _rope(self, x, pos):
        # x: [n, head] fp32 rows of one sequence; NeoX rotate-half, inv_freq = base^(-2i/d)
        half = x.shape[-1] // 2
        inv = self.cfg.rope_base ** (-torch.arange(0, half, dtype=torch.float64) * 2.0 / float(2 * half))
        ang = (pos * inv)
        cs, sn = torch.cos(ang).float(), torch.sin(ang).float()
        a, b = x[:, :half], x[:, half:]
        return torch.cat([a * cs - b * sn, b * cs + a * sn], dim=-1)

    def _store(self, rows):
        """What the span cache gives back for these rows."""
        r = rows.numpy().astype(np.float32)
        if self.kv_mode == KV.QUANT_NONE:
            return torch.from_numpy(r)
        q, z, s = KV.quant_rows(r, self.kv_mode)
        return torch.from_numpy(KV.dequant_rows(q, z, s))

    def step(self, ids, pos):
        """ids: int64 [B]; pos[b]: tokens already cached.  Returns (logits fp32 [B, vocab], next_ids)."""
        global _EXACT, _FT
        _EXACT, _FT = self.exact, self.ft
        try:
            return self._step(ids, pos)
        finally:
            _EXACT, _FT = False, torch.bfloat16

    def _step(self, ids, pos):
        cfg = self.cfg
        nH, nG, hpg = cfg.n_heads, cfg.n_kv, cfg.n_heads // cfg.n_kv
        hd = getattr(cfg, "head", 128)   # 128 everywhere except the parity anchor C0 (Qwen2-0.5B: 64)
        B = ids.shape[0]
        x = self.embed[ids]
        for li, L in enumerate(self.layers):
            xn = self._rms(x, L["g1"])
            qkv = self._gemm(xn, L["qkv"]).float().reshape(B, nH + 2 * nG, hd)
            ao = torch.zeros(B, nH, hd)
            for b in range(B):
                qk = _bf(self._rope(qkv[b, :nH + nG], float(pos[b]))).float()
                q, k, v = qk[:nH], qk[nH:], qkv[b, nH + nG:]
                self.k[li][b].append(self._store(k))
                self.v[li][b].append(self._store(v))
                Kc = torch.stack(self.k[li][b], 1)  # [nG, T, 128]
                Vc = torch.stack(self.v[li][b], 1)
                qh = q.reshape(nG, hpg, hd)
                s = torch.einsum("ghd,gtd->ght", qh.double(), Kc.double()) / math.sqrt(float(hd))
                p = torch.softmax(s, dim=-1)
                ao[b] = torch.einsum("ght,gtd->ghd", p, Vc.double()).float().reshape(nH, hd)
            ao = _bf(ao.reshape(B, nH * hd))
            x = _bf(self._gemm(ao, L["o"]).float() + x.float())
            xn = self._rms(x, L["g2"])
            g = self._gemm(xn, L["gate"], act="silu")
            u = self._gemm(xn, L["up"])
            h = _bf(g.float() * u.float())
            x = _bf(self._gemm(h, L["down"]).float() + x.float())
        xn = self._rms(x, self.gf)
        logits = (torch.matmul(xn.float(), self.lm.float()) if (self.exact or self.ft == torch.float16) else torch.matmul(xn, self.lm)).float()
        return logits, torch.argmax(logits, dim=-1)


def from_stack(stack, kv_mode=KV.QUANT_NONE, exact=False):
    """Build the oracle from a b200spark DecodeStack created with keep_ref=True (dense dequantized weights)."""
    layers = []
    for L in stack.layers:
        d = {"g1": L["g1"].float().cpu(), "g2": L["g2"].float().cpu()}
        for k in ("qkv", "o", "gate", "up", "down"):
            d[k] = (L[k].ref, L[k].ref_bias)
        layers.append(d)
    return RefDecoder(stack.cfg, layers, stack.embed.float().cpu(), stack.gf.float().cpu(), stack.lm_head.ref, kv_mode, exact=exact,
                      ft=getattr(stack, "dtype", torch.bfloat16))


# ---------------------------------------------------------------------------------------------------------------
# CPU baseline timing (bench.py cpu_baseline / --impl reference): bounded sample of the same workload
# ---------------------------------------------------------------------------------------------------------------
def time_cpu_decode(cfg, batch, ctx, sample_layers=2, steps=3, warmup=1, threads=None, seed=1234):
    """Time `steps` decode steps of `sample_layers` decoder layers + final norm + lm_head with contiguous fp32 KV of
    length ctx, bf16 weights (the reference CPU path's medium_bf16 precision) and extrapolate to cfg.layers layers.
    Returns dict(tokens_per_s, s_per_step_full (extrapolated), s_per_step_sampled (measured time of one sampled step),
    s_layer, s_head, threads)."""
    import time
    if threads:
        torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(seed)
    H, nH, nG, I = cfg.hidden, cfg.n_heads, cfg.n_kv, cfg.inter
    hpg = nH // nG
    mk = lambda k, n: (torch.randn(k, n, generator=g) * 0.02).to(torch.bfloat16)
    layers = []
    for _ in range(sample_layers):
        layers.append(dict(qkv=mk(H, (nH + 2 * nG) * 128), o=mk(nH * 128, H), gate=mk(H, I), up=mk(H, I), down=mk(I, H),
                           g1=torch.ones(H), g2=torch.ones(H),
                           k=torch.randn(batch, nG, ctx + steps + warmup, 128, generator=g),
                           v=torch.randn(batch, nG, ctx + steps + warmup, 128, generator=g)))
    lm = mk(H, cfg.vocab)
    x0 = torch.randn(batch, H, generator=g).to(torch.bfloat16)

    def rms(x, gm):
        xf = x.float()
        return (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + cfg.eps) * gm).to(torch.bfloat16)

    def layer_fwd(L, x, T):
        xn = rms(x, L["g1"])
        qkv = torch.matmul(xn, L["qkv"]).float().reshape(batch, nH + 2 * nG, 128)
        q = qkv[:, :nH].reshape(batch, nG, hpg, 128)
        L["k"][:, :, T] = qkv[:, nH:nH + nG]
        L["v"][:, :, T] = qkv[:, nH + nG:]
        s = torch.matmul(q, L["k"][:, :, :T + 1].transpose(-1, -2)) * (1.0 / math.sqrt(128.0))
        p = torch.softmax(s, dim=-1)
        ao = torch.matmul(p, L["v"][:, :, :T + 1]).reshape(batch, nH * 128).to(torch.bfloat16)
        x = (torch.matmul(ao, L["o"]).float() + x.float()).to(torch.bfloat16)
        xn = rms(x, L["g2"])
        gt = torch.matmul(xn, L["gate"]).float()
        h = (gt * torch.sigmoid(gt) * torch.matmul(xn, L["up"]).float()).to(torch.bfloat16)
        return (torch.matmul(h, L["down"]).float() + x.float()).to(torch.bfloat16)

    t_layer, t_head = 0.0, 0.0
    for it in range(warmup + steps):
        T = ctx + it
        t0 = time.perf_counter()
        x = x0
        for L in layers:
            x = layer_fwd(L, x, T)
        t1 = time.perf_counter()
        logits = torch.matmul(rms(x, torch.ones(H)), lm)
        _ = torch.argmax(logits.float(), dim=-1)
        t2 = time.perf_counter()
        if it >= warmup:
            t_layer += (t1 - t0) / sample_layers
            t_head += t2 - t1
    s_layer, s_head = t_layer / steps, t_head / steps
    full = s_layer * cfg.layers + s_head
    return dict(tokens_per_s=batch / full, s_per_step_full=full, s_per_step_sampled=s_layer * sample_layers + s_head,
                s_layer=s_layer, s_head=s_head, threads=torch.get_num_threads())

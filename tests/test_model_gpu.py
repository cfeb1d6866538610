"""GPU parity at model level: a 2-layer decoder stack stepped through the C ABI (eager and CUDA-graph replay) against
the reference-CPU-path oracle (oracle/decoder_ref.py) on identical synthetic weights and token ids.

Tolerance: max |logit diff| <= 1e-2 * max |logit|; greedy token identical whenever the oracle's own
top-2 margin exceeds twice that bound (a smaller margin is a coin flip for ANY bf16 implementation)."""
import numpy as np
import pytest
import torch

from oracle import decoder_ref as DR
from oracle import kvcache_ref as KV

pytestmark = pytest.mark.gpu


def _glue_refs():
    pass


@pytest.mark.parametrize("wbits,group,kv", [(4, -1, "none"), (8, -1, "none"), (4, 128, "none"), (8, -1, "i8"), (4, -1, "u4")])
def test_tiny_decoder_logits_and_tokens(wbits, group, kv):
    from b200spark import model
    B, steps = 2, 6
    st = model.DecodeStack(model.TINY, B, 64, wbits=wbits, group=group, kv=kv, span=16, keep_ref=True)
    ref = DR.from_stack(st, {"none": KV.QUANT_NONE, "i8": KV.QUANT_I8, "u4": KV.QUANT_U4}[kv])
    ref.reset(B)
    ids = torch.tensor([3, 777], dtype=torch.int64)
    for t in range(steps):
        st.ids.copy_(ids.cuda())
        nxt = st.step().cpu()
        torch.cuda.synchronize()
        glog = st.logits.float().cpu()
        rlog, rnext = ref.step(ids, [t] * B)
        # u4 KV: a 1-ulp bf16 difference in a K/V row can move a 4-bit code by one step (1/15 of the row's range),
        # so upstream rounding differences are amplified; the bound is widened for that mode only
        tol = (4e-2 if kv == "u4" else 1e-2) * rlog.abs().max().item()
        err = (glog - rlog).abs().max().item()
        assert err <= tol, (t, err, tol)
        assert torch.equal(nxt, torch.argmax(glog, dim=-1)), "argmax kernel must be bit-exact on its own logits"
        top2 = torch.topk(rlog, 2, dim=-1).values
        for b in range(B):
            if (top2[b, 0] - top2[b, 1]).item() > 2 * tol:
                assert nxt[b].item() == rnext[b].item(), (t, b)
        ids = nxt
    assert st.lens_old.cpu().tolist() == [steps] * B


def test_fused_norm_stack_matches_unfused():
    from b200spark import model
    B = 2
    a = model.DecodeStack(model.TINY, B, 64, wbits=4, span=16, seed=9, fuse_norm=True)
    b = model.DecodeStack(model.TINY, B, 64, wbits=4, span=16, seed=9, fuse_norm=False)
    assert a.fuse_norm and not b.fuse_norm
    ids = torch.tensor([5, 6], dtype=torch.int64, device="cuda")
    for t in range(4):
        a.ids.copy_(ids); b.ids.copy_(ids)
        a.step(); nb = b.step().clone()
        torch.cuda.synchronize()
        err = (a.logits.float() - b.logits.float()).abs().max().item()
        assert err <= 1e-2 * b.logits.float().abs().max().item(), (t, err)
        ids = nb


def test_graph_replay_matches_eager():
    from b200spark import model
    B = 3
    a = model.DecodeStack(model.TINY, B, 64, wbits=4, span=16, seed=7)
    b = model.DecodeStack(model.TINY, B, 64, wbits=4, span=16, seed=7)
    b.capture()
    ids = torch.tensor([1, 2, 3], dtype=torch.int64, device="cuda")
    for t in range(5):
        a.ids.copy_(ids); b.ids.copy_(ids)
        na = a.step().clone()
        nb = b.step().clone()
        torch.cuda.synchronize()
        assert torch.equal(a.logits, b.logits), t
        assert torch.equal(na, nb)
        ids = na
    assert b.lens_new.cpu().tolist() == [6] * B


def test_glue_ops():
    from b200spark import ops, BIN_ADD, BIN_MUL
    g = torch.Generator().manual_seed(0)
    x = torch.randn(5, 3584, generator=g).to(torch.bfloat16)
    gm = (1 + 0.1 * torch.randn(3584, generator=g)).to(torch.bfloat16)
    y = ops.rmsnorm(x.cuda(), gm.cuda(), 1e-6).float().cpu()
    xf = x.float()
    ref = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-6) * gm.float()
    assert (y - ref).abs().max().item() <= 2 ** -8 * ref.abs().max().item() + 1e-6
    a = torch.randn(1000, generator=g).to(torch.bfloat16); b = torch.randn(1000, generator=g).to(torch.bfloat16)
    assert torch.equal(ops.binary(a.cuda(), b.cuda(), BIN_ADD).cpu(), (a.float() + b.float()).to(torch.bfloat16))
    assert torch.equal(ops.binary(a.cuda(), b.cuda(), BIN_MUL).cpu(), (a.float() * b.float()).to(torch.bfloat16))
    tab = torch.randn(50, 512, generator=g).to(torch.bfloat16)
    ids = torch.tensor([0, 49, 7], dtype=torch.int64)
    assert torch.equal(ops.embedding(tab.cuda(), ids.cuda()).cpu(), tab[ids])
    lg = torch.randn(4, 152064, generator=g).to(torch.bfloat16)
    lg[1, 5] = 100.0; lg[1, 99] = 100.0  # tie -> lowest index
    assert torch.equal(ops.argmax(lg.cuda()).cpu(), torch.tensor([int(torch.argmax(lg[i].float())) for i in range(4)]))
    assert ops.argmax(lg.cuda())[1].item() == 5
    # rotary vs fp64 reference (NeoX rotate-half, csrc/core/kernel/cuda/rotary.cu)
    nH, nG = 4, 2
    qkv = torch.randn(2, (nH + 2 * nG) * 128, generator=g).to(torch.bfloat16)
    pos = torch.tensor([0, 1234], dtype=torch.int32)
    out = ops.rotary(qkv.clone().cuda(), pos.cuda(), nH, nG, base=1e6).float().cpu().reshape(2, -1, 128)
    xin = qkv.float().reshape(2, -1, 128).double()
    inv = 1e6 ** (-torch.arange(0, 64, dtype=torch.float64) * 2 / 128)
    for bb in range(2):
        ang = pos[bb].double() * inv
        cs, sn = torch.cos(ang), torch.sin(ang)
        r = torch.cat([xin[bb, :, :64] * cs - xin[bb, :, 64:] * sn, xin[bb, :, 64:] * cs + xin[bb, :, :64] * sn], -1)
        assert (out[bb, :nH + nG] - r[:nH + nG].float()).abs().max().item() <= 2e-2
        assert torch.equal(out[bb, nH + nG:], xin[bb, nH + nG:].float())  # V untouched


@pytest.mark.parametrize("B", [1, 64])
def test_full_width_qwen2_7b_layer_and_lm_head(B):
    """model-level parity was only checked on the 512-wide TINY config.  ONE full-width Qwen2-7B layer
    (hidden 3584, 28/4 heads, inter 18944, int4 per-channel — exactly the launches bench.py times at this batch: the
    mma.sync GEMV family at B=1, the tcgen05 family + persistent gate/up pair at B=64) + final norm + the 152064-wide bf16
    lm_head + argmax, three decode steps from an empty cache, against the reference-CPU-path oracle."""
    from b200spark import model
    steps = 3
    st = model.DecodeStack(model.QWEN2_7B, B, 32, wbits=4, group=-1, kv="none", span=16, keep_ref=True, layers=1)
    ref = DR.from_stack(st, KV.QUANT_NONE)
    ref.reset(B)
    ids = torch.randint(0, model.QWEN2_7B.vocab, (B,), generator=torch.Generator().manual_seed(4321), dtype=torch.int64)
    for t in range(steps):
        st.ids.copy_(ids.cuda())
        nxt = st.step().cpu()
        torch.cuda.synchronize()
        glog = st.logits.float().cpu()
        rlog, rnext = ref.step(ids, [t] * B)
        # 1e-2 of the logit range + one bf16 ulp at that magnitude: both sides round their logits to bf16,
        # so two results that agree to 1e-2 before rounding can land 2 ulps apart (seen: 0.0625 = 2 ulp at |logit| 6.2)
        mx = rlog.abs().max().item()
        tol = 1e-2 * mx + 2.0 ** (np.floor(np.log2(mx)) - 7)
        err = (glog - rlog).abs().max().item()
        assert err <= tol, (t, err, tol)
        assert torch.equal(nxt, torch.argmax(glog, dim=-1))
        top2 = torch.topk(rlog, 2, dim=-1).values
        agree = 0
        for b in range(B):
            if (top2[b, 0] - top2[b, 1]).item() > 2 * tol:
                assert nxt[b].item() == rnext[b].item(), (t, b)
                agree += 1
        ids = nxt
    assert st.lens_old.cpu().tolist() == [steps] * B


def test_config_c0_qwen2_0p5b_bf16_decode_parity():
    """BASELINE.json configs[0], the reference-parity anchor: Qwen2-0.5B (hidden 896, 14 q-heads / 2 kv-heads of 64, inter 4864,
    24 layers, vocab 151936), bf16 weights (no quantization), batch 1, decoded token by token to sequence length 128 — every
    operator of the graph through the C ABI (dense bf16 GEMVs, head-64 rotary + cache append + span attention over 16-token
    spans, RMSNorm, argmax) against the restated reference CPU path (oracle/decoder_ref.py).

    Two bf16 implementations of a 24-layer stack each round ~200 times per token, so they sit ~1.5e-2 of the logit range apart
    (measured) although neither is wrong; the bound that means something is the distance to the EXACT result (same graph, no
    intermediate rounding): the GPU path may be at most 1.5x as far from it as the reference CPU path is.  Asserted per
    checked step: (1) |gpu - exact| <= max(1.5 |ref - exact|, 1e-2 range); (2) |gpu - ref| <= 3e-2 range; (3) the greedy
    token equals the exact path's whenever its top-2 margin exceeds twice the gpu error bound."""
    from b200spark import model
    B, T = 1, 128
    st = model.DecodeStack(model.QWEN2_05B, B, T + 8, wbits=16, group=-1, kv="none", span=16, keep_ref=True)
    ref = DR.from_stack(st, KV.QUANT_NONE)
    exact = DR.from_stack(st, KV.QUANT_NONE, exact=True)
    ref.reset(B); exact.reset(B)
    ids = torch.tensor([1234], dtype=torch.int64)
    worst_g = worst_r = 0.0
    for t in range(T):
        st.ids.copy_(ids.cuda())
        nxt = st.step().cpu()
        rlog, rnext = ref.step(ids, [t] * B)
        elog, enext = exact.step(ids, [t] * B)
        if t < 8 or t % 16 == 15 or t == T - 1:
            torch.cuda.synchronize()
            glog = st.logits.float().cpu()
            mx = elog.abs().max().item()
            eg, er = (glog - elog).abs().max().item(), (rlog - elog).abs().max().item()
            worst_g, worst_r = max(worst_g, eg / mx), max(worst_r, er / mx)
            assert eg <= max(1.5 * er, 1e-2 * mx), (t, eg, er, mx)
            assert (glog - rlog).abs().max().item() <= 3e-2 * mx, t
            top2 = torch.topk(elog, 2, dim=-1).values
            if (top2[0, 0] - top2[0, 1]).item() > 2 * max(1.5 * er, 1e-2 * mx):
                assert nxt[0].item() == enext[0].item(), t
        ids = enext  # all three follow the exact path's tokens: the caches stay comparable for all 128 steps
    assert st.lens_old.cpu().tolist() == [T]
    print("C0: worst |logit error| / logit range vs the exact graph: b200spark %.2e, reference CPU path (bf16) %.2e" % (worst_g, worst_r))


@pytest.mark.parametrize("wbits,group,kv,B", [(4, -1, "none", 2), (8, -1, "i8", 2), (4, 128, "u4", 3), (4, -1, "none", 20)])
def test_tiny_decoder_fp16(wbits, group, kv, B):
    """The whole decode step in fp16 (activations, scales, unquantized cache, logits) against the oracle rounding to fp16
    at the same points: split-K GEMV (B <= 16) and the tcgen05 GEMMs with the RMSNorm hand-off (B = 20), eager then a
    captured graph."""
    from b200spark import model
    steps = 5
    st = model.DecodeStack(model.TINY, B, 64, wbits=wbits, group=group, kv=kv, span=16, keep_ref=True, dtype=torch.float16)
    assert st.logits.dtype == torch.float16
    ref = DR.from_stack(st, {"none": KV.QUANT_NONE, "i8": KV.QUANT_I8, "u4": KV.QUANT_U4}[kv])
    ref.reset(B)
    ids = (torch.arange(B, dtype=torch.int64) * 37 + 3) % model.TINY.vocab
    for t in range(steps):
        if t == 3:
            st.capture()
        st.ids.copy_(ids.cuda())
        nxt = st.step().cpu()
        torch.cuda.synchronize()
        glog = st.logits.float().cpu()
        rlog, rnext = ref.step(ids, [t] * B)
        tol = (4e-2 if kv == "u4" else 1e-2) * rlog.abs().max().item()
        err = (glog - rlog).abs().max().item()
        assert err <= tol, (t, err, tol)
        assert torch.equal(nxt, torch.argmax(glog, dim=-1))
        top2 = torch.topk(rlog, 2, dim=-1).values
        for b in range(B):
            if (top2[b, 0] - top2[b, 1]).item() > 2 * tol:
                assert nxt[b].item() == rnext[b].item(), (t, b)
        ids = nxt

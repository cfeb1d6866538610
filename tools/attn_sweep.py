"""Time the span-attention decode kernel (graph-captured, every launch reads a different layer's cache: >> L2) under env knobs.
usage: python tools/attn_sweep.py [batch] [ctx] [kv]   (knobs swept: B2_ATTN_STAGES x B2_ATTN_CTAS_PER_SM)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "dash-infer_b200", "python"))
import torch  # noqa: E402
from b200spark import model, ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
ctx = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
kv = sys.argv[3] if len(sys.argv) > 3 else "none"
st = model.DecodeStack(model.QWEN2_7B, B, ctx + 64, wbits=4, kv=kv, layers=12)
st.set_context(ctx)
st.step()
torch.cuda.synchronize()
cur = int(st.lens_new[0].item())


def timeit(attn):
    fns = [(lambda L=L: attn(st.q, L["cache"], st.lens_new, st.max_len, st.ws, out=st.ao)) for L in st.layers]
    for f in fns:
        f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(8):
            for f in fns:
                f()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (8 * len(fns))


nb = st.attn.algo_bytes(B * cur)
for stages in ("", "2", "3", "4"):
    for occ in ("", "1", "2", "3", "4", "5", "6"):
        for k, v in (("B2_ATTN_STAGES", stages), ("B2_ATTN_CTAS_PER_SM", occ)):
            if v:
                os.environ[k] = v
            else:
                os.environ.pop(k, None)
        try:
            attn = ops.SpanAttn(st.layers[0]["cache"].cfg, B)
            us = timeit(attn)
            print("stages=%-2s ctas/sm=%-2s  %7.2f us  %6.0f GB/s" % (stages or "-", occ or "-", us, nb / us / 1e3), flush=True)
        except Exception as e:  # noqa: BLE001
            print("stages=%s ctas/sm=%s failed: %s" % (stages, occ, e), flush=True)

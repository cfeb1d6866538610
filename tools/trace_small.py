"""Timeline of one small-batch launch (build with B2_EXTRA_NVCC=-DB2_TRACE): globaltimer stamps of CTA 0 / the merging CTA,
printed as ns since the kernel's first stamp.  python tools/trace_small.py [batch] [ctx] [kv]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "dash-infer_b200", "python"))
import torch  # noqa: E402
from b200spark import model, lib  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
ctx = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
kv = sys.argv[3] if len(sys.argv) > 3 else "none"
wbits = int(sys.argv[4]) if len(sys.argv) > 4 else 4
st = model.DecodeStack(model.QWEN2_7B, B, ctx + 64, wbits=wbits, kv=kv, layers=4)
st.set_context(ctx)
ws = st.ws
GEMV_EV = ["entry", "barriers+sync", "first TMA issued", "pdl_wait done", "activations staged", "first weights landed", "main loop end",
           "partial written+fence", "atomic done", "last CTA: reduce start", "last CTA: reduce end", "store done (ng 0)",
           "x copies issued", "x landed", "x pass done"]
ATTN_EV = ["entry", "pdl_wait done", "decomposition done", "Q fragments", "first tile landed", "tile loop end", "cta merge done",
           "partial fenced", "l1 merge start", "l1 merge end", "final merge start", "final merge end"]


def read(fn, names):
    buf = (C.c_ulonglong * 32)()
    rc = getattr(lib, fn)(buf)
    assert rc == 0
    t0 = buf[0]
    return ", ".join("%s %+d" % (n, int(buf[i]) - int(t0)) for i, n in enumerate(names) if buf[i])


def timed(fns, reps=20):
    for f in fns:
        f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            for f in fns:
                f()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * len(fns))


io = {"gateup": (st.xn, st.gate), "qkv": (st.xn, st.qkv), "o": (st.ao, st.x), "down": (st.gate, st.x)}
for key in ("qkv", "o", "gateup", "down"):
    src, dst = io[key]
    fns = [(lambda L=L: L[key](src, ws, out=dst)) for L in st.layers]
    us = timed(fns)
    print("%-7s B=%d  %.2f us/launch   trace(ns): %s" % (key, B, us, read("b2_debug_trace_gemv", GEMV_EV)), flush=True)
if B <= 16:  # the self-contained RMSNorm form of the column-parallel GEMVs
    H = st.cfg.hidden
    for key, gk in (("qkv", "g1"), ("gateup", "g2")):
        dst = io[key][1]
        fns = [(lambda L=L: L[key](st.x, ws, out=dst, norm_in=(None, L[gk], H, st.cfg.eps))) for L in st.layers]
        us = timed(fns)
        print("%-7s+norm B=%d  %.2f us/launch   trace(ns): %s" % (key, B, us, read("b2_debug_trace_gemv", GEMV_EV)), flush=True)
fns = [(lambda L=L: st.attn(st.q, L["cache"], st.lens_new, st.max_len, ws, out=st.ao)) for L in st.layers]
us = timed(fns)
print("attn    B=%d ctx=%d kv=%s  %.2f us/launch   trace(ns): %s" % (B, ctx, kv, us, read("b2_debug_trace_attn", ATTN_EV)), flush=True)

"""Back-to-back kernel sequences from the decode step, replayed from a CUDA graph (programmatic dependent launch as in the
step): us per sequence.  python tools/seq_bench.py [batch]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "dash-infer_b200", "python"))
import torch  # noqa: E402
from b200spark import model, ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
st = model.DecodeStack(model.QWEN2_7B, B, 256, wbits=4, layers=6)
st.set_context(64)
ws, cfg, H = st.ws, st.cfg, st.cfg.hidden


def timed(seq_of_layer, reps=10):
    def run():
        for L in st.layers:
            seq_of_layer(L)
    run()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            run()
    g.replay()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / (reps * len(st.layers)))
    return best


def s_down(L): L["down"](st.gate, ws, out=st.x, residual=st.x)
def s_o(L): L["o"](st.ao, ws, out=st.x, residual=st.x)
def s_norm1(L): ops.rmsnorm(st.x, L["g1"], cfg.eps, out=st.xn)
def s_norm2(L): ops.rmsnorm(st.x, L["g2"], cfg.eps, out=st.xn)
def s_qkv(L): L["qkv"](st.xn, ws, out=st.qkv)
def s_qkv_self(L): L["qkv"](st.x, ws, out=st.qkv, norm_in=(None, L["g1"], H, cfg.eps))
def s_gu(L): L["gateup"](st.xn, ws, out=st.gate)
def s_gu_self(L): L["gateup"](st.x, ws, out=st.gate, norm_in=(None, L["g2"], H, cfg.eps))


cases = {
    "down": [s_down], "qkv": [s_qkv], "qkv_self": [s_qkv_self], "norm": [s_norm1], "gateup": [s_gu], "gateup_self": [s_gu_self], "o": [s_o],
    "down+norm+qkv": [s_down, s_norm1, s_qkv], "down+qkv_self": [s_down, s_qkv_self],
    "o+norm+gateup": [s_o, s_norm2, s_gu], "o+gateup_self": [s_o, s_gu_self],
    "mlp: norm+gateup+down": [s_norm2, s_gu, s_down], "mlp: gateup_self+down": [s_gu_self, s_down],
}
for name, seq in cases.items():
    us = timed(lambda L: [f(L) for f in seq])
    print("B=%d %-24s %.2f us" % (B, name, us), flush=True)

"""Profiling driver (run under ncu on the GPU box): launches the hot kernels on Qwen2-7B shapes.
  python tools/prof.py --what gemm --batch 1 --layers 6 --iters 3
  python tools/prof.py --what attn --batch 64
  python tools/prof.py --what step --batch 8        (eager decode steps, every kernel of the step)
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "dash-infer_b200", "python"))
import torch  # noqa: E402
from b200spark import model  # noqa: E402
from b200spark._lib import ACT_SILU  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--what", default="gemm")
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--ctx", type=int, default=2048)
ap.add_argument("--layers", type=int, default=6)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--kv", default="none")
ap.add_argument("--wbits", type=int, default=4)
a = ap.parse_args()
st = model.DecodeStack(model.QWEN2_7B, a.batch, a.ctx + 64, wbits=a.wbits, kv=a.kv, layers=a.layers)
if a.what in ("attn", "step"):
    st.set_context(a.ctx)
torch.cuda.synchronize()
if a.what == "step":
    st.step()  # plans / workspace growth outside the profiled region
    torch.cuda.synchronize()
if a.what == "fp8":
    from b200spark import ops
    q8x = ops.quant_fp8(st.xn)
    q8g = ops.quant_fp8(st.gate)
    for L in st.layers:  # plans outside the profiled region
        L["gateup"].op.run_fp8(q8x, st.ws, out=st.gate)
    torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(a.iters):
    if a.what == "fp8":
        for L in st.layers:
            ops.quant_fp8(st.xn, out=q8x)
            L["gateup"].op.run_fp8(q8x, st.ws, out=st.gate)
            L["down"].op.run_fp8(q8g, st.ws, out=st.x)
            L["qkv"].op.run_fp8(q8x, st.ws, out=st.qkv)
    elif a.what == "gemm":
        for L in st.layers:
            if st.fuse_swiglu:
                L["gateup"](st.xn, st.ws, out=st.gate)
            else:
                L["gate"](st.xn, st.ws, out=st.gate, act=ACT_SILU)
            L["down"](st.gate, st.ws, out=st.x)
            L["qkv"](st.xn, st.ws, out=st.qkv)
            L["o"](st.ao, st.ws, out=st.x)
        st.lm_head(st.xn, st.ws, out=st.logits)
    elif a.what == "attn":
        for L in st.layers:
            st.attn(st.q, L["cache"], st.lens_new, st.max_len, st.ws, out=st.ao)
    else:
        st.step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")

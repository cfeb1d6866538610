"""Micro-benchmark sweep for the weight-streaming GEMM (run on the GPU box).
Each config cycles over `nw` distinct weight sets (>> L2) and reports us/launch and GB/s."""
import os
import sys
import itertools

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "dash-infer_b200", "python"))
import torch  # noqa: E402
from b200spark import ops, quantize as PQ, lib  # noqa: E402

SHAPES = {"gate": (3584, 18944), "down": (18944, 3584), "qkv": (3584, 4608), "o": (3584, 3584), "lm": (3584, 152064)}


def bench(K, N, M, wbits=4, nw=8, rounds=20, env=None, pdl=1):
    for k, v in (env or {}).items():
        os.environ[k] = str(v)
    lib.b2_set_pdl(pdl)
    g = torch.Generator(device="cuda").manual_seed(0)
    hs = []
    for i in range(nw):
        w = (torch.randn(K, N, generator=g, device="cuda") * 0.02).to(torch.bfloat16)
        if wbits == 4:
            q, s, z = PQ.quantize_a16w4(w, -1)
        elif wbits == 8:
            q, s, z = PQ.quantize_a16w8(w, -1)
        else:
            q, s, z = w, None, None
        hs.append(ops.GemmWQ(K, N, wbits, -1, max_m=M).prepare(q, s, z))
    a = torch.randn(M, K, generator=g, device="cuda").to(torch.bfloat16)
    out = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    ws = ops.Workspace()
    for h in hs:
        h(a, ws, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # graph-captured loop (what the decode step does)
    gph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gph):
        for _ in range(rounds):
            for h in hs:
                h(a, ws, out=out)
    gph.replay()
    torch.cuda.synchronize()
    e0.record()
    gph.replay()
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) * 1e-3 / (rounds * nw)
    nb = hs[0].algo_bytes(M)
    for k in (env or {}):
        os.environ.pop(k, None)
    return t * 1e6, nb / t / 1e9


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "base"
    if which == "base":
        for name, (K, N) in SHAPES.items():
            for M in (1, 8):
                for env, pdl in (({}, 1), ({}, 0), ({"B2_GEMM_FORCE_SPLIT": 1}, 1), ({"B2_GEMM_MAX_SPLIT": 4}, 1)):
                    us, gbs = bench(K, N, M, env=env, pdl=pdl)
                    print(f"{name:5s} M={M} env={env} pdl={pdl}: {us:7.2f} us  {gbs:7.1f} GB/s", flush=True)
    else:
        envs = [dict(kv.split("=") for kv in cfg.split(",") if kv) for cfg in sys.argv[2:]] or [{}]
        ms = [int(x) for x in os.environ.get("SWEEP_M", "1,8").split(",")]
        names = os.environ.get("SWEEP_SHAPES", "gate,down,qkv,o").split(",")
        nw = int(os.environ.get("SWEEP_NW", "8"))
        for name in names:
            K, N = SHAPES[name]
            for M in ms:
                for env in envs:
                    e2 = dict(env); pdl = int(e2.pop('PDL', 1))
                    us, gbs = bench(K, N, M, wbits=int(os.environ.get('SWEEP_WBITS', '4')), env=e2, nw=nw, pdl=pdl)
                    print(f"{name:5s} M={M} env={env}: {us:7.2f} us  {gbs:7.1f} GB/s", flush=True)

"""Launch one GEMM shape a few times (for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "dash-infer_b200", "python"))
import torch
from b200spark import ops, quantize as PQ
K, N, M, wbits = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]) if len(sys.argv) > 4 else 4
g = torch.Generator(device="cuda").manual_seed(0)
hs = []
for i in range(3):
    w = (torch.randn(K, N, generator=g, device="cuda") * 0.02).to(torch.bfloat16)
    q, s, z = (PQ.quantize_a16w4(w, -1) if wbits == 4 else PQ.quantize_a16w8(w, -1)) if wbits != 16 else (w, None, None)
    hs.append(ops.GemmWQ(K, N, wbits, -1, max_m=M).prepare(q, s, z))
a = torch.randn(M, K, generator=g, device="cuda").to(torch.bfloat16)
out = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
ws = ops.Workspace()
for _ in range(3):
    for h in hs:
        h(a, ws, out=out)
torch.cuda.synchronize()
print("done")

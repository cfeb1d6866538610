"""Timeline of CTA 0 of the tcgen05 GEMM (needs a library built with B2_EXTRA_NVCC=-DB2_TC_TRACE)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "dash-infer_b200", "python"))
import torch
from b200spark import ops, quantize as PQ, lib
K, N, M = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
g = torch.Generator(device="cuda").manual_seed(0)
w = (torch.randn(K, N, generator=g, device="cuda") * 0.02).to(torch.bfloat16)
q, s, z = PQ.quantize_a16w4(w, -1)
h = ops.GemmWQ(K, N, 4, -1, max_m=M).prepare(q, s, z)
a = torch.randn(M, K, generator=g, device="cuda").to(torch.bfloat16)
ws = ops.Workspace()
for _ in range(3):
    out = h(a, ws)
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * (16 * 256))()
lib.b2_debug_tc_trace.argtypes = [ctypes.c_void_p]
rc = lib.b2_debug_tc_trace(buf)
t = [[buf[r * 256 + i] for i in range(256)] for r in range(16)]
t0 = t[7][4]
names = ["prod_issue", "mma_ready", "mma_issued", "x_ready", "dq_wfull", "dq_afree", "dq_stored"]
nt = max(i for i in range(256) if t[6][i]) + 1
print("tiles", nt, "start", 0, "dfull", t[7][0] - t0, "pre-final", t[7][1] - t0, "epi_done", t[7][2] - t0, "end", t[7][3] - t0)
for j in range(nt):
    print(j, " ".join(f"{names[r]}={t[r][j] - t0:7d}" for r in range(7)), "| mma issue deltas", [t[8 + i][j] - t[1][j] for i in range(8)])

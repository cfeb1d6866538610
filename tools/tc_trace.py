"""Timeline of the tcgen05 GEMM (needs a library built with B2_EXTRA_NVCC=-DB2_TC_TRACE).
python tools/tc_trace.py K N M : CTA-0 stage timeline (SM clocks) of the last launch and, for a graph of back-to-back
launches over different weights, globaltimer (ns) entry/ready/end stamps of CTA 0 and the last CTA of every launch."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "dash-infer_b200", "python"))
import torch
from b200spark import ops, quantize as PQ, lib
K, N, M = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
g = torch.Generator(device="cuda").manual_seed(0)
hs = []
for i in range(4):
    w = (torch.randn(K, N, generator=g, device="cuda") * 0.02).to(torch.bfloat16)
    q, s, z = PQ.quantize_a16w4(w, -1)
    hs.append(ops.GemmWQ(K, N, 4, -1, max_m=M).prepare(q, s, z))
a = torch.randn(M, K, generator=g, device="cuda").to(torch.bfloat16)
ws = ops.Workspace()
out = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
for h in hs:
    h(a, ws, out=out)
torch.cuda.synchronize()
gph = torch.cuda.CUDAGraph()
with torch.cuda.graph(gph):
    for _ in range(3):
        for h in hs:
            h(a, ws, out=out)
gph.replay(); torch.cuda.synchronize()
gph.replay(); torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * (16 * 256))()
lib.b2_debug_tc_trace.argtypes = [ctypes.c_void_p]
lib.b2_debug_tc_trace(buf)
gt = (ctypes.c_ulonglong * (64 * 8))()
nl = ctypes.c_uint()
lib.b2_debug_tc_gt.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
lib.b2_debug_tc_gt(gt, ctypes.byref(nl))
n = nl.value
print("launches", n)
rows = [[gt[(i % 64) * 8 + j] for j in range(8)] for i in range(n - 12, n)]
base = rows[0][0]
print("per launch, ns relative to the first: cta0 entry / pdl-ready / end | last cta entry / pdl-ready / end")
prev_end = None
for r in rows:
    e = [x - base for x in r]
    gap = "" if prev_end is None else " gap_after_prev_end=%d" % (min(e[0], e[4]) - prev_end)
    print("  cta0 %7d %7d %7d (dur %5d) | last %7d %7d %7d (dur %5d)%s" % (e[0], e[1], e[2], e[2] - e[0], e[4], e[5], e[6], e[6] - e[4], gap))
    prev_end = max(e[2], e[6])
t = [[buf[r * 256 + i] for i in range(256)] for r in range(16)]
t0 = t[7][5]
names = ["prod_issue", "mma_ready", "mma_issued", "x_ready", "dq_wfull", "dq_afree", "dq_stored"]
nt = max(i for i in range(256) if t[2][i]) + 1
print("clocks from entry: setup_done", t[7][4] - t0, "pdl_ready", t[7][6] - t0, "dfull", t[7][0] - t0, "tmem_loaded", t[7][7] - t0,
      "parked", t[7][1] - t0, "reduced", t[7][2] - t0, "end", t[7][3] - t0, "stages", nt)
for j in range(nt):
    print(j, " ".join(f"{names[r]}={t[r][j] - t0:7d}" for r in range(7)), "| last mma issue at +%d" % (t[15][j] - t[1][j]))

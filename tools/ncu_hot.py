"""Print the hottest SASS instructions (by warp-stall samples) of a kernel from an .ncu-rep.
usage: python tools/ncu_hot.py report.ncu-rep [kernel-id like :::1] [topN]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
kid = sys.argv[2] if len(sys.argv) > 2 else ":::1"
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", kid], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]
ie, si, ws = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
top, tot, byop = [], 0, collections.Counter()
for idx, r in enumerate(rows[hi + 1:]):
    if len(r) <= max(stall_cols) or not r[ie].isdigit():
        continue
    n = int(r[ie]); s = int(r[ws]) if r[ws].isdigit() else 0
    why = sorted(((int(r[c]) if r[c].isdigit() else 0, hdr[c][6:]) for c in stall_cols), reverse=True)[:2]
    top.append((s, idx, n, r[si].strip()[:80], why)); tot += s
    parts = r[si].strip().split()
    op = (parts[1] if parts[0].startswith("@") else parts[0]).split(".")[0]
    byop[op] += n
print("total samples", tot, " total warp-instr", sum(byop.values()))
print("opcode mix:", [(o, n) for o, n in byop.most_common(12)])
for s, idx, n, src, why in sorted(top, reverse=True)[:topn]:
    print(f"{s:6d} #{idx:5d} exec={n:8d} {src:80s} {why}")

"""Timing-only ablations of the tcgen05 GEMM (build with B2_EXTRA_NVCC=-DB2_TC_ABLATE; results are wrong by design).
Shows which pipeline role bounds a k128 stage: python tools/tc_ablate.py [shape] [M]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from gemm_sweep import SHAPES, bench  # noqa: E402

NAMES = {0: "full", 1: "-rowsum loads", 2: "-dequant", 4: "-mma", 8: "-x tma", 16: "-tcgen05.st", 32: "-w tma",
         3: "-rowsum -dequant", 9: "-rowsum -xtma", 18: "-dequant(-st)", 11: "-rowsum -dequant -xtma (mma + w tma)",
         15: "only w tma", 47: "nothing (barriers only)", 43: "mma only", 5: "-rowsum -mma", 6: "-dequant -mma"}

if __name__ == "__main__":
    shape = sys.argv[1] if len(sys.argv) > 1 else "gate"
    M = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    K, N = SHAPES[shape]
    for mask in (0, 1, 2, 4, 8, 16, 32, 3, 9, 11, 43, 15, 47, 5, 6):
        os.environ["B2_TC_ABLATE"] = str(mask)
        us, gbs = bench(K, N, M, 4, nw=4, rounds=10)
        print("mask %2d %-40s %7.2f us  %7.1f GB/s" % (mask, NAMES.get(mask, ""), us, gbs), flush=True)

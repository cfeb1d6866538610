"""Round-2 ablation: which role bounds the tcgen05 GEMM at batch 64 / 32 / 17 (build with B2_EXTRA_NVCC=-DB2_TC_ABLATE)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from gemm_sweep import SHAPES, bench  # noqa: E402
from tc_ablate import NAMES  # noqa: E402

for shape, M, masks in (("gate", 64, (0, 8, 9, 1, 4, 12, 2, 32, 40, 11, 43, 15, 47)), ("gate", 32, (0, 8, 9, 4)), ("gate", 17, (0, 8, 9, 4)),
                        ("down", 64, (0, 8, 9, 4)), ("qkv", 64, (0, 8, 9, 4, 47))):
    K, N = SHAPES[shape]
    for mask in masks:
        os.environ["B2_TC_ABLATE"] = str(mask)
        us, gbs = bench(K, N, M, 4, nw=4, rounds=10)
        print("%s M=%d mask %2d %-40s %7.2f us  %7.1f GB/s" % (shape, M, mask, NAMES.get(mask, ""), us, gbs), flush=True)

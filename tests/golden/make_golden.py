"""Generate golden vectors for the weight quantizers by importing the REFERENCE's own Python
(`/root/reference/python/pyhie/allspark/model/quantization_utils.py`, unmodified).

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py
Writes tests/golden/quant_golden.npz (inputs are regenerated from seeds, outputs stored).

The compiled pybind module `_allspark` and `model_base` (which needs it) are stubbed: the
quantizer functions only use `re`/`make_tensor` from there and never touch them on the
code path exercised here.
"""
import os
import re
import sys
import types

import numpy as np
import torch

REF = "/root/reference/python/pyhie/allspark"


def import_reference_quantizer():
    for name, path in (("pyhie", None), ("pyhie.allspark", REF), ("pyhie.allspark.model", REF + "/model")):
        m = types.ModuleType(name)
        m.__path__ = [path] if path else []
        sys.modules[name] = m
    sys.modules["pyhie.allspark._allspark"] = types.ModuleType("pyhie.allspark._allspark")
    mb = types.ModuleType("pyhie.allspark.model.model_base")
    mb.re = re
    mb.make_tensor = lambda name, data=None: None
    mb.__all__ = ["re", "make_tensor"]
    sys.modules["pyhie.allspark.model.model_base"] = mb
    import pyhie.allspark.model.quantization_utils as qu  # noqa
    import pyhie.allspark.quantization as qz  # noqa
    return qu, qz


CASES = [
    # (name, K, N, ft, bits, group)
    ("w4_perc_bf16", 256, 96, "bf16", 4, -1),
    ("w4_perc_fp16_oddN", 130, 37, "fp16", 4, -1),
    ("w4_g128_bf16", 384, 64, "bf16", 4, 128),
    ("w4_g64_bf16_padK", 200, 48, "bf16", 4, 64),
    ("w8_perc_bf16", 256, 96, "bf16", 8, -1),
    ("w8_g128_fp16", 384, 40, "fp16", 8, 128),
    ("w8_g64_bf16_padK", 200, 24, "bf16", 8, 64),
    ("w4_perc_bf16_const_col", 64, 8, "bf16", 4, -1),  # a zero-range column -> scale forced to 1
    ("w4_g72_bf16", 216, 40, "bf16", 4, 72),    # group sizes that do not divide the 64-k tile (round 2: per-word look-up)
    ("w4_g40_fp16_padK", 100, 24, "fp16", 4, 40),
    ("w4_g32_bf16", 128, 16, "bf16", 4, 32),
]


def make_weight(name, K, N, ft, seed):
    g = torch.Generator().manual_seed(seed)
    w = torch.randn(K, N, generator=g) * 0.02
    if "const_col" in name:
        w[:, 3] = 0.0125
    return w.to(torch.bfloat16 if ft == "bf16" else torch.float16)


def main():
    qu, qz = import_reference_quantizer()
    out = {}
    for i, (name, K, N, ft, bits, group) in enumerate(CASES):
        w = make_weight(name, K, N, ft, 1000 + i)
        extra = {"SubChannel": group != -1, "GroupSize": group}
        mode = qz.QuantizeConfig.QuantMode.A16W4 if bits == 4 else qz.QuantizeConfig.QuantMode.A16W8
        cfg = types.SimpleNamespace(quantize_mode=mode, extra_option=extra,
                                    weight_type="uint4" if bits == 4 else "int8")
        fn = qu.quantize_gemm_weight_a16w4_torch if bits == 4 else qu.quantize_gemm_weight_a16w8_torch
        q, s, z = fn(w, cfg)
        out[name + ".q"] = q.numpy()
        out[name + ".s"] = s.float().numpy()
        out[name + ".z"] = z.float().numpy()
    # GPTQ repack (quantization_utils.py:391-437), synthetic AutoGPTQ-style tensors
    g = torch.Generator().manual_seed(77)
    K, N, gs = 256, 64, 128
    qweight = torch.randint(-2**31, 2**31 - 1, (K // 8, N), generator=g, dtype=torch.int64).to(torch.int32)
    qzeros = torch.randint(-2**31, 2**31 - 1, (K // gs, N // 8), generator=g, dtype=torch.int64).to(torch.int32)
    # keep zero nibbles <= 14 so that +1 stays a uint4 (what AutoGPTQ guarantees)
    qzeros = qzeros & 0x66666666
    scales = (torch.rand(K // gs, N, generator=g) * 0.01 + 0.001).to(torch.float16)
    info = ("m.weight", {"m.weight": qweight, "m.qzeros": qzeros, "m.scales": scales})
    q, s, z = qu.repack_gptq_to_a16wX(info, 4)
    out["gptq4.qweight"] = qweight.numpy()
    out["gptq4.qzeros"] = qzeros.numpy()
    out["gptq4.scales"] = scales.float().numpy()
    out["gptq4.q"] = q.numpy()
    out["gptq4.s"] = s.float().numpy()
    out["gptq4.z"] = z.float().numpy()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "quant_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()

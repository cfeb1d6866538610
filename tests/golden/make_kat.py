"""Known-answer vectors for the oracle itself: seeded synthetic inputs -> SHA-256 of the bit-exact
outputs (KV row quantizer, span packer, GPTQ repack) and stored values of small floating-point outputs (dequant-GEMM math,
attention).  They pin the oracle against drift; they are NOT reference outputs (the reference has no fixtures for these
pieces: "parity unpinned" in oracle/__init__.py).   python tests/golden/make_kat.py  ->  tests/golden/oracle_kat.json"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import kvcache_ref as KV  # noqa: E402
from oracle import quant_ref as Q  # noqa: E402


def sha(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        a = np.ascontiguousarray(a)
        h.update(str(a.dtype).encode() + str(a.shape).encode() + a.tobytes())
    return h.hexdigest()


def bf16_round(x):
    return KV.bits_to_f32(KV.bf16_bits(x.astype(np.float32)))


def compute():
    out = {}
    rng = np.random.default_rng(20240924)
    # --- KV row quantizer (bit-exact integer/param outputs)
    x = bf16_round(rng.standard_normal((37, 4, 128)).astype(np.float32) * 1.7)
    x[3, 1] = 0.25          # constant row: scale clamps to 1e-5
    x[5, 0, ::2] = -3.0     # two-valued row
    for mode, name in ((KV.QUANT_I8, "i8"), (KV.QUANT_U4, "u4")):
        q, z, s = KV.quant_rows(x, mode)
        out["kv_quant_rows_" + name] = sha(q, z, s)
    # --- span packer: bytes of the spans after appending 37 tokens (span 16, 4 groups)
    for mode, name in ((KV.QUANT_NONE, "none"), (KV.QUANT_I8, "i8"), (KV.QUANT_U4, "u4")):
        c = KV.SpanCacheRef(mode, 16, 4)
        c.add_sequence()
        for t in range(37):
            c.append(0, t, x[t])
        out["span_bytes_" + name] = sha(*[np.frombuffer(bytes(sp), np.uint8) for sp in c.spans[0]])
    # --- weight quantizers + packers (bit-exact)
    w = bf16_round(rng.standard_normal((192, 40)).astype(np.float32) * 0.02)
    q4, s4, z4 = Q.quantize_a16w4(w, "bf16", 64)
    q8, s8, z8 = Q.quantize_a16w8(w, "bf16", -1)
    out["quantize_a16w4_g64"] = sha(q4, s4, z4)
    out["quantize_a16w8_perc"] = sha(q8, s8, z8)
    # --- floating-point pieces: store values (compared with a tolerance)
    a = bf16_round(rng.uniform(-1, 1, (3, 192)).astype(np.float32))
    y = Q.gemm_wq_math(a, Q.unpack_u4x2(q4, 40), s4, z4, 64)
    out["gemm_wq_math_w4_g64"] = [float(v) for v in y[:, :8].ravel()]
    kc, vc = KV.SpanCacheRef(KV.QUANT_I8, 16, 2), KV.SpanCacheRef(KV.QUANT_I8, 16, 2)
    kc.add_sequence(); vc.add_sequence()
    for t in range(21):
        kc.append(0, t, x[t, :2]); vc.append(0, t, x[t, 2:])
    qv = bf16_round(rng.standard_normal((1, 4, 128)).astype(np.float32))
    o = KV.attention_ref(qv, kc, vc, [21], 4, 1.0 / np.sqrt(128))
    out["attention_ref_i8"] = [float(v) for v in o[0, :, :4].ravel()]
    return out


if __name__ == "__main__":
    res = compute()
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_kat.json")
    json.dump(res, open(path, "w"), indent=1)
    print("wrote", path, {k: (v[:12] if isinstance(v, str) else len(v)) for k, v in res.items()})

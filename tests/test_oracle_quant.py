"""CPU: the oracle and the product quantizer against golden vectors produced by the reference's own Python
(tests/golden/make_golden.py imports /root/reference/.../quantization_utils.py).  Bit-exact."""
import numpy as np
import pytest
import torch

from oracle import quant_ref as Q

CASES = [
    ("w4_perc_bf16", 256, 96, "bf16", 4, -1),
    ("w4_perc_fp16_oddN", 130, 37, "fp16", 4, -1),
    ("w4_g128_bf16", 384, 64, "bf16", 4, 128),
    ("w4_g64_bf16_padK", 200, 48, "bf16", 4, 64),
    ("w8_perc_bf16", 256, 96, "bf16", 8, -1),
    ("w8_g128_fp16", 384, 40, "fp16", 8, 128),
    ("w8_g64_bf16_padK", 200, 24, "bf16", 8, 64),
    ("w4_perc_bf16_const_col", 64, 8, "bf16", 4, -1),
    ("w4_g72_bf16", 216, 40, "bf16", 4, 72),    # group sizes that do not divide the 64-k tile (round 2: per-word look-up)
    ("w4_g40_fp16_padK", 100, 24, "fp16", 4, 40),
    ("w4_g32_bf16", 128, 16, "bf16", 4, 32),
]


def make_weight(name, K, N, ft, seed):  # must mirror tests/golden/make_golden.py
    g = torch.Generator().manual_seed(seed)
    w = torch.randn(K, N, generator=g) * 0.02
    if "const_col" in name:
        w[:, 3] = 0.0125
    return w.to(torch.bfloat16 if ft == "bf16" else torch.float16)


@pytest.mark.parametrize("i", range(len(CASES)))
def test_oracle_matches_reference_golden(golden, i):
    name, K, N, ft, bits, group = CASES[i]
    w = make_weight(name, K, N, ft, 1000 + i).float().numpy()
    q, s, z = (Q.quantize_a16w4 if bits == 4 else Q.quantize_a16w8)(w, ft, group)
    assert np.array_equal(q, golden[name + ".q"])
    assert np.array_equal(s, golden[name + ".s"])
    assert np.array_equal(z, golden[name + ".z"])


@pytest.mark.parametrize("i", range(len(CASES)))
def test_product_quantizer_matches_reference_golden(golden, i):
    from b200spark import quantize as PQ
    name, K, N, ft, bits, group = CASES[i]
    w = make_weight(name, K, N, ft, 1000 + i)
    q, s, z = (PQ.quantize_a16w4 if bits == 4 else PQ.quantize_a16w8)(w, group)
    assert np.array_equal(q.numpy(), golden[name + ".q"])
    assert np.array_equal(s.float().numpy(), golden[name + ".s"])
    assert np.array_equal(z.float().numpy(), golden[name + ".z"])


def test_gptq_repack_golden(golden):
    from b200spark import quantize as PQ
    q, s, z = Q.repack_gptq_a16w4(golden["gptq4.qweight"], golden["gptq4.qzeros"], golden["gptq4.scales"])
    assert np.array_equal(q, golden["gptq4.q"]) and np.array_equal(z, golden["gptq4.z"]) and np.array_equal(s, golden["gptq4.s"])
    q2, s2, z2 = PQ.repack_gptq_a16w4(torch.from_numpy(golden["gptq4.qweight"]), torch.from_numpy(golden["gptq4.qzeros"]),
                                      torch.from_numpy(golden["gptq4.scales"]).half())
    assert np.array_equal(q2.numpy(), golden["gptq4.q"]) and np.array_equal(z2.float().numpy(), golden["gptq4.z"])


def test_pack_unpack_roundtrip_and_dequant():
    rng = np.random.default_rng(0)
    q = rng.integers(0, 16, size=(40, 33), dtype=np.uint8)
    p = Q.pack_u4x2(q)
    assert p.shape == (40, 17)
    assert np.array_equal(Q.unpack_u4x2(p, 33), q)
    # reference test packer semantics (operator_gemm_lowp_test.cpp:18-29): low nibble = even column
    assert p[0, 0] == (q[0, 0] | (q[0, 1] << 4))
    s = np.full((1, 33), 0.5, np.float32)
    z = np.full((1, 33), 3.0, np.float32)
    assert np.allclose(Q.dequant(q, s, z), (q.astype(np.float32) - 3.0) * 0.5)


def test_gemm_oracles_agree():
    """fp32 math oracle vs the CPU-path (bf16 oneDNN) oracle: within the stated 2e-2 min(abs,rel)."""
    rng = np.random.default_rng(1)
    K, N, M = 512, 64, 3
    w = Q.to_ft((rng.standard_normal((K, N)) * 0.02).astype(np.float32), "bf16")
    a = Q.to_ft(rng.uniform(-1, 1, (M, K)).astype(np.float32), "bf16")
    qd, s, z = Q.quantize_a16w4(w, "bf16", -1)
    qu = Q.unpack_u4x2(qd, N)
    c1 = Q.gemm_wq_math(a, qu, s, z)
    c2 = Q.gemm_wq_cpu_path(a, qu, s, z)
    assert Q.err_min_abs_rel(c1, c2) < 2e-2
    # quantization error: half a step + the bf16 rounding of the stored zero/scale
    wq = Q.dequant(qu, s, z)
    assert np.max(np.abs(wq - w) / s[0][None, :]) <= 0.5 + 0.1

"""CPU known-answer test of the "exact integer" dequantisation constants (dash-infer_b200/csrc/b2_common.cuh, Ft<H>): a 4-bit
code OR-ed into mantissa bits 3..6 of a 16-bit float constant must read back as constant + code (low-nibble plane) or
16 * (constant + code) (high-nibble plane of int8) — in bf16 with 16.0 / 256.0, in fp16 with 128.0 / 2048.0 — and the
sub-channel path's  (constant + q) - (constant + 8)  must be exact.  The kernels rely on these bit identities; nothing here
touches the GPU."""
import numpy as np
import torch

MASK = 0x0078  # where the weight image (words rotated left by 3) puts a nibble inside each 16-bit half


def _as_float(bits, dtype):
    t = torch.tensor(np.array(bits, dtype=np.uint16).astype(np.int16))
    return t.view(dtype).float().numpy()


def test_bf16_constants():
    q = np.arange(16)
    lo = _as_float((q << 3) | 0x4180, torch.bfloat16)   # kMagic   = bf16 16.0
    hi = _as_float((q << 3) | 0x4380, torch.bfloat16)   # kMagicHi = bf16 256.0
    assert np.array_equal(lo, 16.0 + q) and np.array_equal(hi, 16.0 * (16.0 + q))
    assert ((q << 3) & ~MASK).max() == 0
    # (16 + q) + (-24) = q - 8 exactly in bf16 (0xC1C0 = -24.0)
    assert _as_float([0xC1C0], torch.bfloat16)[0] == -24.0
    x = torch.tensor(lo).to(torch.bfloat16) + torch.tensor(-24.0).to(torch.bfloat16)
    assert np.array_equal(x.float().numpy(), q - 8.0)


def test_fp16_constants():
    q = np.arange(16)
    lo = _as_float((q << 3) | 0x5800, torch.float16)    # fp16 128.0
    hi = _as_float((q << 3) | 0x6800, torch.float16)    # fp16 2048.0
    assert np.array_equal(lo, 128.0 + q) and np.array_equal(hi, 16.0 * (128.0 + q))
    assert _as_float([0xD840], torch.float16)[0] == -136.0
    x = torch.tensor(lo).to(torch.float16) + torch.tensor(-136.0).to(torch.float16)
    assert np.array_equal(x.float().numpy(), q - 8.0)


def test_affine_dequant_identity_int8_planes():
    """sum_k a (b + lo) + sum_k a 16 (b + hi) = sum_k a u + 17 b sum_k a  with u = lo + 16 hi: the zero-point constant the
    prepare step folds into the stored zero (wq_gemm.cu: zbias = 17 b (+128 for signed int8), b = 16 or 128)."""
    rng = np.random.default_rng(0)
    a = rng.standard_normal(256)
    u = rng.integers(0, 256, 256)
    for b in (16.0, 128.0):
        lhs = np.sum(a * (b + (u & 15))) + np.sum(a * 16.0 * (b + (u >> 4)))
        assert abs(lhs - (np.sum(a * u) + 17.0 * b * np.sum(a))) < 1e-6 * (1 + abs(lhs))

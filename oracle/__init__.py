"""CPU oracle for the quantized decode hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker (or as the CPU
baseline being timed), never on the GPU product path.

Every function restates a piece of the reference (modelscope/dash-infer @ f3cca8e)
and cites the file:line it follows.  Pinning status:

* weight quantizers (``quant_ref``): PINNED — checked bit-exactly against golden
  vectors produced by importing the reference's own
  ``python/pyhie/allspark/model/quantization_utils.py`` (``tests/golden/make_golden.py``).
* dequant-GEMM (``quant_ref.gemm_*``): restates the reference tests' in-test fp32 CPU
  loops (``tests/cpp/operator/cuda/operator_gemm_lowp_test.cpp:131-205``); the reference
  holds no golden vectors for it (random data + tolerance only).
* KV-cache quantizer / span packer / attention (``kvcache_ref``, ``attention_ref``):
  **parity unpinned** — the reference has no numeric test or fixture for I8/U4 spans and
  its CPU build is not producible here; these follow the CUDA sources'
  arithmetic line by line with IEEE fp32 division in place of ``__fdividef``.
"""

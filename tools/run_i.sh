set -x
timeout 900 python -m pytest tests/test_fp8_gpu.py tests/test_ref_pin_gpu.py -m gpu -q -p no:cacheprovider --timeout 300 -s 2>&1 | tail -80 > gpurun_out/pytest_r2i.log
tail -6 gpurun_out/pytest_r2i.log
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_attn_gpu.py tests/test_model_gpu.py -m gpu -q -p no:cacheprovider --timeout 300 2>&1 | tail -30 > gpurun_out/pytest_r2i2.log
tail -4 gpurun_out/pytest_r2i2.log

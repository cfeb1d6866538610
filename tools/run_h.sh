set -x
timeout 1200 python -m pytest tests/test_gemm_gpu.py tests/test_model_gpu.py -m gpu -q -p no:cacheprovider --timeout 600 -x 2>&1 | tail -60 > gpurun_out/pytest_r2h.log
tail -4 gpurun_out/pytest_r2h.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_r2h.json 2> gpurun_out/bench_r2h.err

set -x
timeout 300 python -m pytest tests/test_model_gpu.py -m gpu -q -p no:cacheprovider --timeout 250 -s -k "c0 or tiny" 2>&1 | tail -30 > gpurun_out/pytest_r2l.log
tail -6 gpurun_out/pytest_r2l.log
timeout 400 python -m pytest tests/test_attn_gpu.py tests/test_host_ops_gpu.py tests/test_fp8_gpu.py -m gpu -q -p no:cacheprovider --timeout 120 2>&1 | tail -15 > gpurun_out/pytest_r2l2.log
tail -4 gpurun_out/pytest_r2l2.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_r2l.json 2> gpurun_out/bench_r2l.err

set -x
timeout 900 python -m pytest tests/test_attn_gpu.py tests/test_ref_pin_gpu.py tests/test_model_gpu.py -m gpu -q -p no:cacheprovider --timeout 600 2>&1 | tail -60 > gpurun_out/pytest_r2f.log
tail -4 gpurun_out/pytest_r2f.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_r2f.json 2> gpurun_out/bench_r2f.err
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu --wbits 8 --kv i8 --ctx 32768 --batch 1 --sub-batches "" > gpurun_out/bench_r2f_c2.json 2> gpurun_out/bench_r2f_c2.err

set -x
timeout 120 python -m pytest tests/test_gemm_gpu.py -m gpu -q -p no:cacheprovider --timeout 60 -x -k "small_shapes_all_m or subchannel" 2>&1 | tail -15 > gpurun_out/pytest_r2j0.log
tail -3 gpurun_out/pytest_r2j0.log
grep -q "passed" gpurun_out/pytest_r2j0.log && ! grep -q "failed" gpurun_out/pytest_r2j0.log || exit 1
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_fp8_gpu.py tests/test_model_gpu.py -m gpu -q -p no:cacheprovider --timeout 120 2>&1 | tail -40 > gpurun_out/pytest_r2j.log
tail -5 gpurun_out/pytest_r2j.log
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu --model llama3-8b --group 128 --batch 32 --sub-batches "" > gpurun_out/bench_r2j_c3.json 2> gpurun_out/bench_r2j_c3.err
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu --batch 32 --sub-batches "" > gpurun_out/bench_r2j_b32.json 2> gpurun_out/bench_r2j_b32.err

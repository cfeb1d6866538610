set -x
export B2_GEMM_TC_CG=1
timeout 1500 python -m pytest tests/test_attn_gpu.py tests/test_ref_pin_gpu.py tests/test_model_gpu.py tests/test_host_ops_gpu.py -m gpu -q -p no:cacheprovider --timeout 900 2>&1 | tail -400 > gpurun_out/pytest_r2b.log
tail -8 gpurun_out/pytest_r2b.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu --wbits 8 --kv i8 --ctx 32768 --batch 1 --sub-batches "" > gpurun_out/bench_r2b_c2.json 2> gpurun_out/bench_r2b_c2.err
tail -c 300 gpurun_out/bench_r2b_c2.json
# ---- CTA pairs (cta_group::2): short leash
export B2_GEMM_TC_CG=2
timeout 300 python -m pytest tests/test_gemm_gpu.py -m gpu -q -p no:cacheprovider --timeout 120 -x -k "small_shapes_all_m and 64" 2>&1 | tail -60 > gpurun_out/pytest_r2b_cg2a.log
tail -5 gpurun_out/pytest_r2b_cg2a.log
timeout 600 python -m pytest tests/test_gemm_gpu.py -m gpu -q -p no:cacheprovider --timeout 120 2>&1 | tail -120 > gpurun_out/pytest_r2b_cg2.log
tail -8 gpurun_out/pytest_r2b_cg2.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --sub-batches "" > gpurun_out/bench_r2b_cg2.json 2> gpurun_out/bench_r2b_cg2.err
tail -c 1500 gpurun_out/bench_r2b_cg2.json
export B2_GEMM_TC_CG=1
timeout 600 python -m pytest tests/test_comm_gpu.py -m gpu -q -p no:cacheprovider --timeout 120 2>&1 | tail -80 > gpurun_out/pytest_r2b_comm.log
tail -5 gpurun_out/pytest_r2b_comm.log

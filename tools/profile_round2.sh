#!/bin/bash
# Round-2 evidence (run on the GPU box through gpurun): full GPU test suite, the bench lines, per-launch lists of one decode
# step and `ncu --set full` captures of the hot kernels.  Outputs go to gpurun_out/ (summaries are written from there into
# profiles/ by tools/ncu_summary.py on the build box).
set -x
NCU="ncu --clock-control none"
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 300 2>&1 | tail -25 > gpurun_out/r2_pytest_gpu.log
tail -3 gpurun_out/r2_pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --wbits 8 --kv i8 --ctx 32768 --batch 1 --sub-batches "" > gpurun_out/r2_bench_c2.json 2> gpurun_out/r2_bench_c2.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --model llama3-8b --group 128 --batch 32 --sub-batches "" > gpurun_out/r2_bench_c3.json 2> gpurun_out/r2_bench_c3.err
for B in 64 1; do
  timeout 300 $NCU --metrics gpu__time_duration.sum --profile-from-start off --csv --log-file gpurun_out/r2_launches_step_b$B.csv python tools/prof.py --what step --batch $B --layers 28 --iters 1 > /dev/null 2>&1
done
timeout 300 $NCU --set full --import-source on --profile-from-start off -k regex:span_attn -c 2 -f -o gpurun_out/r2_attn_b64 python tools/prof.py --what attn --batch 64 --layers 2 --iters 1 > /dev/null 2>&1
timeout 300 $NCU --set full --import-source on --profile-from-start off -k regex:span_attn -c 2 -f -o gpurun_out/r2_attn_c2 python tools/prof.py --what attn --batch 1 --ctx 32768 --kv i8 --wbits 8 --layers 2 --iters 1 > /dev/null 2>&1
timeout 300 $NCU --set full --import-source on --profile-from-start off -k regex:"wq_gem" -c 5 -f -o gpurun_out/r2_gemm_b64 python tools/prof.py --what gemm --batch 64 --layers 1 --iters 1 > /dev/null 2>&1
timeout 300 $NCU --set full --import-source on --profile-from-start off -k regex:"wq_gem" -c 5 -f -o gpurun_out/r2_gemm_b1 python tools/prof.py --what gemm --batch 1 --layers 1 --iters 1 > /dev/null 2>&1
timeout 300 $NCU --set full --import-source on --profile-from-start off -k regex:"wq_gemm_tc|quant_fp8" -c 4 -f -o gpurun_out/r2_fp8_b64 python tools/prof.py --what fp8 --batch 64 --layers 1 --iters 1 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep

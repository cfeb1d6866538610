#!/bin/bash
cd /root/repo
timeout 200 python -m pytest tests/test_gemm_gpu.py -x -q -k "fp16_all_paths and (4--1-1] or 4--1-17] or 4--1-64])" 2>&1 | tail -4 > gpurun_out/fp16_smoke.log
cat gpurun_out/fp16_smoke.log
grep -q passed gpurun_out/fp16_smoke.log || exit 1
grep -q "failed\|error" gpurun_out/fp16_smoke.log && exit 1
timeout 900 python -m pytest tests/test_gemm_gpu.py -q -k "fp16" 2>&1 | tail -15

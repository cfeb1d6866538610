#!/bin/bash
cd /root/repo
timeout 300 python -m pytest tests/test_attn_gpu.py -x -q -k "fp16" 2>&1 | tail -8
timeout 600 python -m pytest tests/test_attn_gpu.py tests/test_ref_pin_gpu.py -x -q 2>&1 | tail -4

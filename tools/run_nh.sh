#!/bin/bash
cd /root/repo
timeout 300 python -m pytest tests/test_gemm_gpu.py -x -q -k "handoff" 2>&1 | tail -6 > gpurun_out/nh_tests.log
cat gpurun_out/nh_tests.log
grep -q passed gpurun_out/nh_tests.log || exit 1
grep -q failed gpurun_out/nh_tests.log && exit 1
timeout 900 python -m pytest tests/test_model_gpu.py -x -q 2>&1 | tail -5
for v in 0 1; do
  B2_NORM_HANDOFF=$v timeout 300 python bench.py --no-tp-record --sub-batches "" 2>/dev/null | tail -1 > gpurun_out/nh_$v.json
  python - $v <<'PY'
import json, sys
d=json.load(open("gpurun_out/nh_%s.json" % sys.argv[1]))
print("HANDOFF=" + sys.argv[1], d["value"], d["ms_per_step"], d["gpu_launches"], d["e2e"]["value"])
PY
done

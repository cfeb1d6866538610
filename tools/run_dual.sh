#!/bin/bash
cd /root/repo
run() {
  env "$@" timeout 300 python bench.py --no-tp-record --sub-batches "" 2>/dev/null | tail -1 > gpurun_out/dual_x.json
  python - "$*" <<'PY'
import json, sys
d=json.load(open("gpurun_out/dual_x.json"))
print(sys.argv[1], d["value"], d["ms_per_step"], {k.split("[")[1][:10]: (v["us"], v["frac"]) for k, v in d.get("kernels", {}).items()})
PY
}
run B2_GEMM_TC_DUAL=1
run B2_GEMM_TC_DUAL=2 B2_GEMM_TC_DUAL_SLOTS=1
run B2_GEMM_TC_DUAL=2 B2_GEMM_TC_DUAL_SLOTS=2 B2_GEMM_TC_MAX_SPLIT2=6

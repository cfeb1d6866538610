#!/bin/bash
# A/B: cluster split-K x self-contained RMSNorm, whole-step bench at batch 64 / 8 / 1
cd /root/repo
for cl in 0 1; do for ns in 0 1; do
  B2_GEMM_CLUSTER=$cl B2_GEMM_CLUSTER_MAX=${CMAX:-8} B2_NORM_SELF=$ns timeout 600 python bench.py --no-tp-record 2>/dev/null | tail -1 > gpurun_out/ab_${cl}_${ns}.json
  python - <<PY
import json
d=json.load(open("gpurun_out/ab_${cl}_${ns}.json"))
print("CLUSTER=$cl NORM_SELF=$ns", d["value"], d["ms_per_step"], {k:(v["tokens_per_s"],v["ms_per_step"],v["gpu_launches"]) for k,v in d["batches"].items()})
PY
done; done

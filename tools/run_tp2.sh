#!/bin/bash
cd /root/repo
timeout 300 python -m pytest tests/test_gemm_gpu.py -x -q -k "small_shapes or self_contained or cluster or qwen2_72b" 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 2>/dev/null | tail -1 > gpurun_out/bench_n2_e.json
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_n2_e.json"))
print(d["value"], d["ms_per_step"], {k:(v["tokens_per_s"],v["ms_per_step"]) for k,v in d["batches"].items()}, d["tp"]["tokens_per_s"], d["tp"]["ms_per_step"])
PY

#!/bin/bash
# End-of-round GPU pass: full GPU test suite, the default bench line, the batch sweep and the KV-mode variants.
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 400 python bench.py > $O/bench_default.json 2> $O/bench_default.err; tail -c 600 $O/bench_default.json
for b in 1 8 32; do timeout 200 python bench.py --batch $b --no-cpu --steps 20 > $O/bench_b$b.json 2> $O/bench_b$b.err; done
for kv in i8 u4; do timeout 200 python bench.py --batch 64 --kv $kv --no-cpu --steps 20 > $O/bench_b64_$kv.json 2> $O/bench_kv.err; done
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err
python __graft_entry__.py smoke 2>&1 | tail -2
for f in $O/bench_default.json $O/bench_b1.json $O/bench_b8.json $O/bench_b32.json $O/bench_b64_i8.json $O/bench_b64_u4.json $O/bench_reference.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
    print(sys.argv[1].split('/')[-1], d.get('value'), d.get('ms_per_step'), (d.get('roofline') or {}).get('step',{}).get('frac'), (d.get('e2e') or {}).get('value'), (d.get('cpu_baseline') or {}).get('value'))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done

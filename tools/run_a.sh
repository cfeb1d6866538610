set -x
nvidia-smi -L
timeout 1700 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 900 -s 2>&1 | tail -150 > gpurun_out/pytest_r2a.log
tail -5 gpurun_out/pytest_r2a.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err
tail -c 600 gpurun_out/bench_r2a.json
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_r2a.json 2> gpurun_out/bench_ref_r2a.err
cat gpurun_out/bench_ref_r2a.json | head -c 900

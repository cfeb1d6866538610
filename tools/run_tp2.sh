set -x
nvidia-smi -L
timeout 900 python -m pytest tests/test_tp_gpu.py tests/test_attn_gpu.py tests/test_host_ops_gpu.py tests/test_model_gpu.py -m gpu -q -p no:cacheprovider --timeout 600 2>&1 | tail -120 > gpurun_out/pytest_r2_tp2.log
tail -6 gpurun_out/pytest_r2_tp2.log
for c in fused b2 nccl; do
  B2_TP_COLLECTIVE=$c timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu --sub-batches "" > gpurun_out/bench_r2_tp2_$c.json 2> gpurun_out/bench_r2_tp2_$c.err
  tail -c 1200 gpurun_out/bench_r2_tp2_$c.json
done

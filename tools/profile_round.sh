#!/bin/bash
# ncu evidence for profiles/: launch lists of one decode step (B=64, B=1) and --set full captures of the hot kernels.
# Run on the GPU box: bash tools/profile_round.sh ; outputs land in gpurun_out/.
set -x
O=gpurun_out
mkdir -p $O
NCU="ncu --clock-control none --profile-from-start off"
for b in 64 1; do
  $NCU --metrics gpu__time_duration.sum --csv --log-file $O/launches_b$b.csv python tools/prof.py --what step --batch $b --layers 28 --iters 1 > $O/prof_step_b$b.log 2>&1
done
$NCU --set full --import-source on -f -o $O/gemm_b64 python tools/prof.py --what gemm --batch 64 --layers 1 --iters 1 > $O/prof_gemm_b64.log 2>&1
$NCU --set full --import-source on -f -o $O/gemm_b1 python tools/prof.py --what gemm --batch 1 --layers 1 --iters 1 > $O/prof_gemm_b1.log 2>&1
$NCU --set full --import-source on -f -o $O/attn_b64 python tools/prof.py --what attn --batch 64 --layers 1 --iters 1 > $O/prof_attn_b64.log 2>&1
ls -la $O/*.ncu-rep

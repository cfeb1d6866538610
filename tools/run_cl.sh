#!/bin/bash
cd /root/repo
for cfg in "16 64" "8 64" "8 48" "0 64"; do
  set -- $cfg
  for B in 1 8; do
    echo "== CLUSTER_MAX=$1 RING=$2 B=$B"
    if [ "$1" = "0" ]; then export B2_GEMM_CLUSTER=0; else export B2_GEMM_CLUSTER=1; fi
    B2_GEMM_CLUSTER_MAX=$1 B2_GEMM_CLUSTER_RING_KB=$2 timeout 200 python tools/seq_bench.py $B 2>&1 | grep -E " (down|qkv|o|gateup|down\+norm\+qkv|o\+norm\+gateup|down\+qkv_self|o\+gateup_self) "
  done
done

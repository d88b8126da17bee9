#!/bin/bash
# GroupNorm occupancy / unroll / cluster-size experiment: per-layer CUDA-event totals of one 54-forward micro-batch
for cfg in "0 8" "1 8" "2 8" "3 8" "1 16" "1 4" "3 4"; do
  set -- $cfg
  DM_GN_VAR=$1 DM_GN_CLUSTER=$2 DM_BF=54 timeout 200 python tools/profile_target.py layers > gpurun_out/r02_gn_v$1_c$2.log 2>&1
  echo "GN_VAR=$1 CLUSTER=$2 $(tail -1 gpurun_out/r02_gn_v$1_c$2.log | cut -c1-200)"
  grep -E "norm1 |norm2 |\.norm " gpurun_out/r02_gn_v$1_c$2.log | grep -v "transformer_blocks" | awk '{s+=substr($4,4)} END {print "   groupnorm total ms", s}'
done

#!/bin/bash
# weight-stationary igemm: operator tests first (stop on failure), then A/B per-layer sums, then the whole suite + bench
timeout 900 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "conv" 2>&1 | tail -5 | tee gpurun_out/r02_ws_ops.log
grep -q "failed\|error" gpurun_out/r02_ws_ops.log && exit 1
run() {  # label, env...
  label=$1; shift
  env "$@" DM_BF=54 timeout 200 python tools/profile_target.py layers > gpurun_out/r02_sweep_$label.log 2>&1
  python tools/layer_sums.py gpurun_out/r02_sweep_$label.log $label
}
run ws1
run ws0      DM_IGEMM_WS=0
run ws1ng2   DM_IGEMM_NG4=0
run ws1_b
run ws0_b    DM_IGEMM_WS=0
bash tools/gpu_check.sh

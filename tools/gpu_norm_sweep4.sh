#!/bin/bash
# fold/apply GroupNorm with register double buffering (PF) vs occupancy only; weight-stationary igemm off by default
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "groupnorm or conv_with" 2>&1 | tail -2
run() {  # label, env...
  label=$1; shift
  env "$@" DM_BF=54 timeout 200 python tools/profile_target.py layers > gpurun_out/r02_sweep_$label.log 2>&1
  python tools/layer_sums.py gpurun_out/r02_sweep_$label.log $label
}
run s4_v2
run s4_v5 DM_GNFA_VAR=5
run s4_v6 DM_GNFA_VAR=6
run s4_v7 DM_GNFA_VAR=7
run s4_v8 DM_GNFA_VAR=8
run s4_v2_b

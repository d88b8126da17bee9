#!/bin/bash
# generalised epilogue statistics (all GroupNorm producers): correctness, then A/B per-layer sums at Bf=54, then one ncu
# --set full capture of the streaming normalisation kernels
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "groupnorm or conv_with or layernorm" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_parity_full.py -x -q -m gpu 2>&1 | tail -6
run() {  # label, env...
  label=$1; shift
  env "$@" DM_BF=54 timeout 200 python tools/profile_target.py layers > gpurun_out/r02_sweep_$label.log 2>&1
  python tools/layer_sums.py gpurun_out/r02_sweep_$label.log $label
}
run s3_m3
run s3_m1     DM_GN_EPILOGUE=1
run s3_m0     DM_GN_EPILOGUE=0
run s3_m3fa4  DM_GNFA_VAR=4
run s3_m3fa6  DM_GNFA_VAR=6
run s3_m3c4   DM_GNFA_CL=4
run s3_ln3    DM_LN_VAR=3
run s3_base   DM_GN_EPILOGUE=0 DM_LN_VAR=0
run s3_m3_b
timeout 300 python tools/ab.py norm 2>&1 | tail -8
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'layernorm|gn_fold_apply|gn_fused' -c 14 -f \
  -o gpurun_out/r02_norm python tools/ab.py norm > gpurun_out/r02_ncu_norm.log 2>&1
tail -3 gpurun_out/r02_ncu_norm.log

#!/bin/bash
# second streaming-normalisation sweep + one ncu --set full capture of the norm kernels at the big shapes
timeout 300 python tools/ab.py norm 2>&1 | tail -8
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'layernorm|gn_fold_apply|gn_fused' -c 14 -f \
  -o gpurun_out/r02_norm python tools/ab.py norm > gpurun_out/r02_ncu_norm.log 2>&1
tail -3 gpurun_out/r02_ncu_norm.log
run() {  # label, env...
  label=$1; shift
  env "$@" DM_BF=54 timeout 200 python tools/profile_target.py layers > gpurun_out/r02_sweep_$label.log 2>&1
  python tools/layer_sums.py gpurun_out/r02_sweep_$label.log $label
}
run s2_def
run s2_fa4    DM_GNFA_VAR=4
run s2_fa6    DM_GNFA_VAR=6
run s2_fa2c4  DM_GNFA_CL=4
run s2_fa6c4  DM_GNFA_VAR=6 DM_GNFA_CL=4
run s2_ln3    DM_LN_VAR=3
run s2_base   DM_GN_EPILOGUE=0 DM_LN_VAR=0
run s2_def_b

#!/bin/bash
# fold loop of gn_fold_apply_kernel: 8 loads in flight (libdm_b200_alt.so) vs 4, at the config-5 and config-2 shapes
run() {  # label, env...
  label=$1; shift
  env "$@" timeout 300 python tools/profile_target.py layers > gpurun_out/r02_sweep_$label.log 2>&1
  python tools/layer_sums.py gpurun_out/r02_sweep_$label.log $label
}
run f4_c5 DM_BF=15 DM_LAT=128
run f4_c2 DM_BF=54
cp diff-mining_b200/libdm_b200.so /tmp/libdm_main.so
cp diff-mining_b200/libdm_b200_alt.so diff-mining_b200/libdm_b200.so
run f8_c5 DM_BF=15 DM_LAT=128
run f8_c2 DM_BF=54
timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "conv_with" 2>&1 | tail -2
cp /tmp/libdm_main.so diff-mining_b200/libdm_b200.so
run f4_c5_b DM_BF=15 DM_LAT=128

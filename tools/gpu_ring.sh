#!/bin/bash
# experiment: two epilogue chunk buffers per group everywhere (-DIG_RING_SMALL, libdm_b200_alt.so) = one more pipeline stage at BN = 160
run() {  # label
  DM_BF=54 timeout 200 python tools/profile_target.py layers > gpurun_out/r02_sweep_$1.log 2>&1
  python tools/layer_sums.py gpurun_out/r02_sweep_$1.log $1
  python tools/layer_cats.py gpurun_out/r02_sweep_$1.log
}
run ring4_a
cp diff-mining_b200/libdm_b200.so /tmp/libdm_main.so
cp diff-mining_b200/libdm_b200_alt.so diff-mining_b200/libdm_b200.so
run ring2_a
run ring2_b
cp /tmp/libdm_main.so diff-mining_b200/libdm_b200.so
run ring4_b

#!/bin/bash
# persistent cross-attention kernel: tests, op-level A/B, per-layer sums
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "attention" 2>&1 | tail -4 | tee gpurun_out/r02_xattn_ops.log
grep -q "failed\|error" gpurun_out/r02_xattn_ops.log && exit 1
timeout 300 python tools/ab.py xattn 2>&1 | tail -10
run() {  # label, env...
  label=$1; shift
  env "$@" DM_BF=54 timeout 200 python tools/profile_target.py layers > gpurun_out/r02_sweep_$label.log 2>&1
  python tools/layer_sums.py gpurun_out/r02_sweep_$label.log $label
}
run xa2
run xa1 DM_XATTN=1
run xa2_b
run xa1_b DM_XATTN=1
timeout 900 python -m pytest tests/test_gpu_e2e.py -x -q -m gpu 2>&1 | tail -3

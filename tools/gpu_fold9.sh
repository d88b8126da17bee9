#!/bin/bash
# fold/apply GroupNorm: 8 loads in flight in the fold + several clusters per large image; config-5 and config-2 shapes
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "conv_with or groupnorm" 2>&1 | tail -2
timeout 900 python -m pytest tests/test_gpu_parity_full.py tests/test_gpu_e2e.py -x -q -m gpu 2>&1 | tail -2
run() {  # label, env...
  label=$1; shift
  env "$@" timeout 300 python tools/profile_target.py layers > gpurun_out/r02_sweep_$label.log 2>&1
  python tools/layer_sums.py gpurun_out/r02_sweep_$label.log $label
}
run f9_c5 DM_BF=15 DM_LAT=128
run f9_c2 DM_BF=54
timeout 500 python bench.py --config 5 --steps 3 > gpurun_out/r02_bench_c5.log 2>&1; tail -1 gpurun_out/r02_bench_c5.log | cut -c1-300

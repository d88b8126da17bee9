#!/bin/bash
# GELU tail as relu(x) - |x| q: operator + model-level parity, GEGLU layer time
timeout 200 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "conv_igemm" 2>&1 | tail -2
timeout 200 python -m pytest tests/test_gpu_e2e.py -x -q -m gpu -k "parity or golden or layerwise" 2>&1 | tail -2
DM_BF=54 timeout 100 python tools/profile_target.py layers > gpurun_out/r02_sweep_gelu.log 2>&1; python tools/layer_sums.py gpurun_out/r02_sweep_gelu.log gelu; python tools/layer_cats.py gpurun_out/r02_sweep_gelu.log | cut -c1-200

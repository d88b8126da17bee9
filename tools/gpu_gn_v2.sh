#!/bin/bash
# N1 stage 1 (v2): GroupNorm statistics from the conv epilogue via shared-memory column sums + fold/apply cluster kernel
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "groupnorm or conv_with or conv_igemm" 2>&1 | tail -5 > gpurun_out/r02_gnv2_ops.log
cat gpurun_out/r02_gnv2_ops.log
timeout 600 python -m pytest tests/test_gpu_e2e.py -x -q -m gpu -k "groupnorm_statistics or layerwise or typicality_grid or unet_eps" 2>&1 | tail -8 > gpurun_out/r02_gnv2_e2e.log
cat gpurun_out/r02_gnv2_e2e.log
for m in 1 0 1 0; do
  DM_GN_EPILOGUE=$m DM_BF=54 timeout 200 python tools/profile_target.py layers > gpurun_out/r02_layers_gnv2_$m.log 2>&1
  echo "GN_EPILOGUE=$m $(grep DMPROF_TOTAL gpurun_out/r02_layers_gnv2_$m.log | cut -c1-200)"
done

#!/bin/bash
# streaming-normalisation sweep: LayerNorm variants x fold/apply GroupNorm variants, per-layer CUDA-event sums at Bf=54
for v in 0 1 2; do
  DM_LN_VAR=$v timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "layernorm or conv_with" 2>&1 | tail -2
done
run() {  # label, env...
  label=$1; shift
  env "$@" DM_BF=54 timeout 200 python tools/profile_target.py layers > gpurun_out/r02_sweep_$label.log 2>&1
  python tools/layer_sums.py gpurun_out/r02_sweep_$label.log $label
}
run base      DM_GN_EPILOGUE=0 DM_LN_VAR=0
run ln1       DM_GN_EPILOGUE=0 DM_LN_VAR=1
run ln2       DM_GN_EPILOGUE=0 DM_LN_VAR=2
run fa0       DM_GN_EPILOGUE=1 DM_LN_VAR=0 DM_GNFA_VAR=0
run fa1       DM_GN_EPILOGUE=1 DM_LN_VAR=0 DM_GNFA_VAR=1
run fa2       DM_GN_EPILOGUE=1 DM_LN_VAR=0 DM_GNFA_VAR=2
run fa3       DM_GN_EPILOGUE=1 DM_LN_VAR=0 DM_GNFA_VAR=3
run fa4       DM_GN_EPILOGUE=1 DM_LN_VAR=0 DM_GNFA_VAR=4
run fa0c4     DM_GN_EPILOGUE=1 DM_LN_VAR=0 DM_GNFA_VAR=0 DM_GNFA_CL=4
run fa1c4     DM_GN_EPILOGUE=1 DM_LN_VAR=0 DM_GNFA_VAR=1 DM_GNFA_CL=4
run fa1c16    DM_GN_EPILOGUE=1 DM_LN_VAR=0 DM_GNFA_VAR=1 DM_GNFA_CL=16
run base_b    DM_GN_EPILOGUE=0 DM_LN_VAR=0

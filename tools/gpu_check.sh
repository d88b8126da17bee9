#!/bin/bash
# quick confirmation run: the whole GPU suite, the headline bench line, per-layer sums
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_check.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_check.log
tail -4 gpurun_out/r02_pytest_check.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02_smoke_check.log 2>&1; tail -2 gpurun_out/r02_smoke_check.log
timeout 600 python bench.py --config 2 --steps 10 > gpurun_out/r02_bench_c2_check.log 2>&1; tail -1 gpurun_out/r02_bench_c2_check.log | cut -c1-400
DM_BF=54 timeout 300 python tools/profile_target.py layers > gpurun_out/r02_layers_check.log 2>&1; python tools/layer_sums.py gpurun_out/r02_layers_check.log check

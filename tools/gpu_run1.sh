#!/bin/bash
# round-2 GPU session 1: tests, smoke, bench (configs 2/3/5), per-layer timings, ncu launch lists of the benched plans
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02_run1_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/r02_pytest1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest1.log
tail -5 gpurun_out/r02_pytest1.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02_smoke1.log 2>&1; tail -3 gpurun_out/r02_smoke1.log
timeout 600 python bench.py --config 2 --steps 5 > gpurun_out/r02_bench_c2.log 2>&1; tail -1 gpurun_out/r02_bench_c2.log
timeout 400 python bench.py --config 3 --steps 5 > gpurun_out/r02_bench_c3.log 2>&1; tail -1 gpurun_out/r02_bench_c3.log
timeout 500 python bench.py --config 5 --steps 3 > gpurun_out/r02_bench_c5.log 2>&1; tail -1 gpurun_out/r02_bench_c5.log
DM_BF=54 timeout 300 python tools/profile_target.py layers > gpurun_out/r02_layers_bf54.log 2>&1; tail -1 gpurun_out/r02_layers_bf54.log
DM_GRAPH=0 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_typ_launches.csv python tools/profile_target.py typ > gpurun_out/r02_ncu_typ.log 2>&1
tail -2 gpurun_out/r02_typ_launches.csv | cut -c1-200

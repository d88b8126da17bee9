#!/bin/bash
# short evidence refresh after the fold/apply change: config 2 / 3 bench lines on another box, the config-5 launch list
timeout 600 python bench.py --config 2 --steps 10 > gpurun_out/r02_bench_c2_b.log 2>&1; tail -1 gpurun_out/r02_bench_c2_b.log | cut -c1-200
timeout 400 python bench.py --config 3 --steps 10 > gpurun_out/r02_bench_c3_b.log 2>&1; tail -1 gpurun_out/r02_bench_c3_b.log | cut -c1-200
DM_GRAPH=0 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_typ5_launches.csv python tools/profile_target.py typ5 > gpurun_out/r02_ncu_typ5.log 2>&1
DM_BF=54 timeout 300 python tools/profile_target.py layers > gpurun_out/r02_layers_final.log 2>&1; python tools/layer_sums.py gpurun_out/r02_layers_final.log final

#!/bin/bash
# BASELINE config 4 on 8 GPUs: 10 000 images sharded over the ranks, one all-gather of the T maps
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --config 4 --steps 1 --warmup 3 > gpurun_out/r02_bench_c4_8gpu.log 2>&1
tail -1 gpurun_out/r02_bench_c4_8gpu.log | cut -c1-600

#!/bin/bash
# two-GPU confirmation of the final build: NCCL bit-identity test + the config-2 bench line on 2 ranks
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/r02_pytest_multi2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_c2_2gpu.log 2>&1
tail -1 gpurun_out/r02_bench_c2_2gpu.log | cut -c1-400

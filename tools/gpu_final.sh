#!/bin/bash
# round-2 evidence run on one B200: tests, smoke, the three single-GPU bench configurations, ncu launch lists of the benched
# plans, and --set full captures of the two dominant kernels.  Outputs land in gpurun_out/ (copied to profiles/ by hand).
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02_final_smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_final.log
tail -4 gpurun_out/r02_pytest_final.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02_smoke_final.log 2>&1; tail -2 gpurun_out/r02_smoke_final.log
timeout 600 python bench.py --config 2 --steps 10 > gpurun_out/r02_bench_c2.log 2>&1; tail -1 gpurun_out/r02_bench_c2.log | cut -c1-300
timeout 400 python bench.py --config 3 --steps 10 > gpurun_out/r02_bench_c3.log 2>&1; tail -1 gpurun_out/r02_bench_c3.log | cut -c1-300
timeout 500 python bench.py --config 5 --steps 3 > gpurun_out/r02_bench_c5.log 2>&1; tail -1 gpurun_out/r02_bench_c5.log | cut -c1-300
timeout 400 python bench.py --impl reference --config 2 --steps 2 --warmup 1 > gpurun_out/r02_bench_ref_c2.log 2>&1; tail -1 gpurun_out/r02_bench_ref_c2.log | cut -c1-300
DM_BF=54 timeout 300 python tools/profile_target.py layers > gpurun_out/r02_layers_final.log 2>&1; tail -1 gpurun_out/r02_layers_final.log
for t in typ dift typ5 vae; do
  DM_GRAPH=0 timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_${t}_launches.csv python tools/profile_target.py $t > gpurun_out/r02_ncu_$t.log 2>&1
done
DM_BF=8 timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:attention3 -c 1 -o gpurun_out/r02_attn3_d40 -f python tools/profile_target.py attn > gpurun_out/r02_ncu_attn3.log 2>&1
timeout 900 ncu --set full --clock-control none --profile-from-start off -k regex:"igemm_kernel|attention|gn_fused|gn_fold_apply|layernorm" -o gpurun_out/r02_ops_full -f python tools/profile_target.py ops > gpurun_out/r02_ncu_ops.log 2>&1
ls -la gpurun_out/*.ncu-rep

#!/bin/bash
# One GPU session that regenerates everything profiles/r2_* is summarised from (run under gpurun from the repo root):
#   bash benchmarks/profile_round2.sh && python benchmarks/summarize_profiles.py r2      (the second step runs anywhere)
# Every ncu run follows a plain run of the same command that exited 0 (profiling recipe).  The .ncu-rep files are turned
# into raw CSV pages on the box and deleted there: gpurun brings back at most 64 MiB.
K='regex:warp_|tlerp_|tok_|mix_|flow_head_|bn_'
mkdir -p gpurun_out
if [ "$1" != "--no-bench" ]; then
  timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench.log 2> gpurun_out/r2_bench.err; echo "bench rc=$?"
fi
timeout 200 python bench.py --profile-step --steps 1 --warmup 3 > gpurun_out/r2_plain_step.log 2>&1 &&
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 4000 --csv --log-file gpurun_out/r2_launches_step.csv python bench.py --profile-step --steps 1 --warmup 3 > gpurun_out/r2_ncu_step.log 2>&1
echo "launch list rc=$?"
timeout 200 python bench.py --profile-step --steps 1 --warmup 3 > gpurun_out/r2_plain_step2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --profile-from-start off -k "$K" -c 200 -o /tmp/r2_instep_kernels -f python bench.py --profile-step --steps 1 --warmup 3 > gpurun_out/r2_ncu_step2.log 2>&1
echo "in-step full rc=$?"
ncu -i /tmp/r2_instep_kernels.ncu-rep --page raw --csv > gpurun_out/r2_instep_kernels.csv 2>/dev/null
timeout 200 python benchmarks/one_kernel_r2.py > gpurun_out/r2_one_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --profile-from-start off -k "$K" -c 160 -o /tmp/r2_cold_kernels -f python benchmarks/one_kernel_r2.py > gpurun_out/r2_one_ncu.log 2>&1
echo "cold full rc=$?"
ncu -i /tmp/r2_cold_kernels.ncu-rep --page raw --csv > gpurun_out/r2_cold_kernels.csv 2>/dev/null
cat gpurun_out/r2_one_plain.log
du -sh gpurun_out

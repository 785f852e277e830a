#!/bin/bash
# One GPU session that regenerates everything profiles/ is summarised from (run under gpurun from the repo root):
#   bash benchmarks/profile_round.sh && python benchmarks/summarize_profiles.py r1      (the second step runs anywhere)
set -x
K='regex:warp_|tlerp_|tok_|mix_'
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
timeout 200 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2> gpurun_out/bench.err
timeout 120 python -m smow_net_b200.runtime.launch train --model lw --global-batch 16 --steps 10 --graph 2>/dev/null | tail -1 > gpurun_out/train_graph.log
timeout 120 python -m smow_net_b200.runtime.launch train --model lw --global-batch 16 --steps 10 2>/dev/null | tail -1 >> gpurun_out/train_graph.log
timeout 120 python benchmarks/instep_probe.py > gpurun_out/instep_probe.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 4000 --csv --log-file gpurun_out/r1_launches_step.csv python bench.py --profile-step --steps 1 --warmup 3 > gpurun_out/ncu_step.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -k "$K" -c 100 -o gpurun_out/r1_instep_kernels -f python bench.py --profile-step --steps 1 --warmup 3 > gpurun_out/ncu_step2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k "$K" -c 40 -o gpurun_out/r1_cold_kernels -f python benchmarks/one_kernel.py --C 32 --H 128 --B 64 --iters 1 --layout ndhwc > gpurun_out/one_ncu.log 2>&1
timeout 600 python benchmarks/sweep_warp.py --iters 10 --fwd-variants 1 --bwd-variants 2 --out gpurun_out/r1_sweep.jsonl > gpurun_out/sweep.log 2>&1
cat gpurun_out/train_graph.log gpurun_out/instep_probe.log | cut -c1-400
tail -c 600 gpurun_out/bench.log

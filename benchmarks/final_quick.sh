set -x
K='regex:warp_|tlerp_|tok_|mix_'
timeout 300 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
timeout 200 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.log 2> gpurun_out/bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 4000 --csv --log-file gpurun_out/r1_launches_step.csv python bench.py --profile-step --steps 1 --warmup 3 > gpurun_out/ncu_step.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k "$K" -c 40 -o gpurun_out/r1_cold_kernels -f python benchmarks/one_kernel.py --C 32 --H 128 --B 64 --iters 1 --layout ndhwc > gpurun_out/one_ncu.log 2>&1
tail -c 400 gpurun_out/bench.log

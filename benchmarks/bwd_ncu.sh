#!/bin/bash
# source-level ncu capture of the NDHWC warp backward tile kernel (run under gpurun); CSV pages exported on the box
mkdir -p gpurun_out
B=${1:-32}; C=${2:-64}; H=${3:-128}
timeout 120 python benchmarks/bwd_once.py $B $C $H > gpurun_out/bwd_once_plain.log 2>&1 || { cat gpurun_out/bwd_once_plain.log; exit 1; }
timeout 600 ncu --set full --import-source on --clock-control none -k regex:warp_bwd_ndhwc_tile -s 1 -c 1 -o /tmp/bwd_k -f python benchmarks/bwd_once.py $B $C $H > gpurun_out/bwd_ncu.log 2>&1
echo "ncu rc=$?"
ncu -i /tmp/bwd_k.ncu-rep --page raw --csv > gpurun_out/bwd_raw.csv 2>/dev/null
ncu -i /tmp/bwd_k.ncu-rep --page source --csv --print-source sass > gpurun_out/bwd_source_sass.csv 2>/dev/null
ncu -i /tmp/bwd_k.ncu-rep --page source --csv --print-source cuda > gpurun_out/bwd_source_cuda.csv 2>/dev/null
ls -la gpurun_out/bwd_*.csv

#!/bin/bash
# ncu capture of the tokenizer kernels (run under gpurun): plain run first, then --set full with source counters;
# raw + source pages exported as CSV on the box (the .ncu-rep stays in /tmp).
mkdir -p gpurun_out
B=${1:-16}; C=${2:-16}
timeout 120 python benchmarks/tok_once.py $B $C 128 > gpurun_out/tok_once_plain.log 2>&1 || { cat gpurun_out/tok_once_plain.log; exit 1; }
timeout 600 ncu --set full --import-source on --clock-control none -k regex:tok_ -s 4 -c 4 -o /tmp/tok_k -f python benchmarks/tok_once.py $B $C 128 > gpurun_out/tok_ncu.log 2>&1
echo "ncu rc=$?"
ncu -i /tmp/tok_k.ncu-rep --page raw --csv > gpurun_out/tok_raw.csv 2>/dev/null
ncu -i /tmp/tok_k.ncu-rep --page source --csv --print-source sass > gpurun_out/tok_source.csv 2>/dev/null
ls -la gpurun_out/tok_*.csv

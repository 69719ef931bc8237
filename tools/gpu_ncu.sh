#!/bin/bash
# ncu --set full of K1 at one shape (under gpurun): bash tools/gpu_ncu.sh <tag> [shape] [frames]
tag=${1:-ncu}; shape=${2:-180x320}; frames=${3:-2048}
out=gpurun_out/$tag
mkdir -p $out
timeout 300 python tools/bench_k1.py $shape $frames 10 > $out/bench_k1.log 2>&1; cat $out/bench_k1.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:okp_peaks_st -s 2 -c 1 -o $out/prof_k1 \
    python tools/bench_k1.py $shape $frames 2 > $out/ncu_full.log 2>&1
tail -3 $out/ncu_full.log

#!/bin/bash
# ncu --set full of K1 at 64x64 (under gpurun): bash tools/gpu_ncu_k1_64.sh <tag> [epilogue warps ...]
tag=${1:-ncu64}; shift
out=gpurun_out/$tag
mkdir -p $out
for ew in ${@:-0}; do
OKP_STREAM_EPILOGUE_WARPS=$ew timeout 600 ncu --set full --clock-control none --import-source on -k regex:okp_peaks_st -s 2 -c 1 -o $out/prof_64_ew$ew \
    python tools/bench_k1.py 64x64 32768 2 f32 > $out/ncu_ew$ew.log 2>&1
done
ls -la $out

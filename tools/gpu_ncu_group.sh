#!/bin/bash
tag=${1:-ncug}
out=gpurun_out/$tag
mkdir -p $out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:okp_group -s 2 -c 1 -o $out/prof_group_64 \
    python tools/bench_k1.py 64x64 32768 2 f32 > $out/ncu_64.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:okp_group -s 2 -c 1 -o $out/prof_group_180 \
    python tools/bench_k1.py 180x320 4096 2 f32 > $out/ncu_180.log 2>&1
ls -la $out

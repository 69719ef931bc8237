#!/bin/bash
# Tile kernel: launch list (which kernels take the time) and one --set full capture with source counters.
tag=${1:-r2e}
out=gpurun_out/$tag
mkdir -p $out
timeout 300 python tools/bench_k1.py 180x320 2048 5 f32 > $out/bench_k1.log 2>&1; grep -v Warn $out/bench_k1.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $out/launches.csv \
    python tools/bench_k1.py 180x320 2048 2 f32 > $out/launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:okp_peaks_tile -s 2 -c 1 -o $out/prof_tile_180 \
    python tools/bench_k1.py 180x320 2048 2 f32 > $out/ncu_180.log 2>&1
tail -3 $out/ncu_180.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:okp_peaks_tile -s 2 -c 1 -o $out/prof_tile_64 \
    python tools/bench_k1.py 64x64 16384 2 f32 lean > $out/ncu_64.log 2>&1
tail -3 $out/ncu_64.log
ncu -i $out/prof_tile_180.ncu-rep --page source --csv > $out/src180.csv 2>/dev/null
ncu -i $out/prof_tile_64.ncu-rep --page source --csv > $out/src64.csv 2>/dev/null
ls -la $out

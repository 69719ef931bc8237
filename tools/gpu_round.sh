#!/bin/bash
# One GPU call: parity tests, K1 timing, bench line, ncu launch list and one full capture of the top kernel.
# usage (under gpurun): bash tools/gpu_round.sh <tag>
tag=${1:-r01}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $out/pytest_gpu.log
tail -3 $out/pytest_gpu.log
timeout 300 python tools/bench_k1.py 180x320 2048 10 > $out/bench_k1.log 2>&1
timeout 300 python tools/bench_k1.py 64x64 32768 10 >> $out/bench_k1.log 2>&1
cat $out/bench_k1.log
if grep -q 'pytest exit 0' $out/pytest_gpu.log; then timeout 240 python tools/sweep_k1.py 180x320 2048 3 > $out/sweep_k1.log 2>&1; timeout 120 python tools/sweep_k1.py 64x64 32768 3 >> $out/sweep_k1.log 2>&1; grep -v '^$' $out/sweep_k1.log | tail -60; fi
timeout 900 python bench.py --steps 10 --warmup 3 > $out/bench.json 2> $out/bench.err; tail -2 $out/bench.err; cat $out/bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_reference.json 2>> $out/bench.err; cat $out/bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:okp_ -c 60 --csv --log-file $out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-frames 256 --e2e-steps 1 > $out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:okp_peaks -s 2 -c 2 -o $out/prof_k1 \
    python tools/bench_k1.py 180x320 2048 2 > $out/ncu_full.log 2>&1
ls -la $out

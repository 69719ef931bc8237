#!/bin/bash
# One GPU call: parity tests, K1 timing (f32 + bf16), bench lines, ncu launch list, full captures of K1, config 5.
# usage (under gpurun): bash tools/gpu_round.sh <tag>
tag=${1:-r01g}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $out/pytest_gpu.log
tail -25 $out/pytest_gpu.log
for dt in f32 bf16; do
  timeout 300 python tools/bench_k1.py 180x320 4096 10 $dt >> $out/bench_k1.log 2>&1
  timeout 300 python tools/bench_k1.py 64x64 32768 10 $dt >> $out/bench_k1.log 2>&1
done
cat $out/bench_k1.log
timeout 900 python bench.py --steps 20 --warmup 3 > $out/bench.json 2> $out/bench.err; tail -2 $out/bench.err; cat $out/bench.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_reference.json 2>> $out/bench.err; cat $out/bench_reference.json
timeout 600 python tools/bench_config5.py 256 5 > $out/config5.json 2> $out/config5.err; tail -2 $out/config5.err; cat $out/config5.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:okp_ -c 60 --csv --log-file $out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-frames 256 --e2e-steps 1 > $out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:okp_peaks_st -s 2 -c 1 -o $out/prof_k1_f32 \
    python tools/bench_k1.py 180x320 4096 2 f32 > $out/ncu_full_f32.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:okp_peaks_st -s 2 -c 1 -o $out/prof_k1_bf16 \
    python tools/bench_k1.py 180x320 4096 2 bf16 > $out/ncu_full_bf16.log 2>&1
ls -la $out

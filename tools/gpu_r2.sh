#!/bin/bash
# Round-2 GPU call: parity tests (all, not -x), kernel timings at both shapes / element types, one bench line.
# usage (under gpurun): bash tools/gpu_r2.sh <tag> [pytest-args]
tag=${1:-r2}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q ${2:-} > $out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $out/pytest_gpu.log
tail -40 $out/pytest_gpu.log
for dt in f32 bf16; do
  timeout 200 python tools/bench_k1.py 180x320 4096 10 $dt >> $out/bench_k1.log 2>&1
  timeout 200 python tools/bench_k1.py 64x64 32768 10 $dt >> $out/bench_k1.log 2>&1
  timeout 200 python tools/bench_k1.py 64x64 32768 10 $dt lean >> $out/bench_k1.log 2>&1
done
cat $out/bench_k1.log
timeout 600 python bench.py --steps 20 --warmup 3 > $out/bench.json 2> $out/bench.err; tail -5 $out/bench.err; cat $out/bench.json

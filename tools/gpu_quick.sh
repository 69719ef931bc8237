#!/bin/bash
# Short GPU call: parity tests, K1 + grouping timings at both shapes and element types, one bench line.
# usage (under gpurun): bash tools/gpu_quick.sh <tag>
tag=${1:-q}
out=gpurun_out/$tag
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $out/pytest_gpu.log
tail -12 $out/pytest_gpu.log
for dt in f32 bf16; do
  timeout 120 python tools/bench_k1.py 180x320 4096 10 $dt >> $out/bench_k1.log 2>&1
  timeout 120 python tools/bench_k1.py 64x64 32768 10 $dt >> $out/bench_k1.log 2>&1
done
cat $out/bench_k1.log
timeout 300 python bench.py --steps 20 --warmup 3 > $out/bench.json 2> $out/bench.err; tail -3 $out/bench.err; cut -c1-400 $out/bench.json

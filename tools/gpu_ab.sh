#!/bin/bash
# A/B of the two K1 forms (under gpurun): parity tests with the default (persistent) kernel, then timings of both.
tag=${1:-ab}
out=gpurun_out/$tag
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $out/pytest_gpu.log
tail -12 $out/pytest_gpu.log
for which in strip stream; do
  for dt in f32 bf16; do
    OKP_PEAKS_KERNEL=$which timeout 120 python tools/bench_k1.py 180x320 4096 10 $dt >> $out/bench_k1.log 2>&1
    OKP_PEAKS_KERNEL=$which timeout 120 python tools/bench_k1.py 64x64 32768 10 $dt >> $out/bench_k1.log 2>&1
  done
done
cat $out/bench_k1.log

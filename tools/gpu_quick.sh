#!/bin/bash
# Short GPU call while iterating on K1: parity tests, K1 timing at both shapes, optional knob sweep.
# usage (under gpurun): bash tools/gpu_quick.sh <tag> [sweep]
tag=${1:-quick}
out=gpurun_out/$tag
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $out/pytest_gpu.log
tail -15 $out/pytest_gpu.log
timeout 300 python tools/bench_k1.py 180x320 2048 10 > $out/bench_k1.log 2>&1
timeout 300 python tools/bench_k1.py 64x64 32768 10 >> $out/bench_k1.log 2>&1
cat $out/bench_k1.log
if [ "$2" = "sweep" ]; then
  timeout 600 python tools/sweep_k1.py 180x320 2048 5 > $out/sweep_k1.log 2>&1
  timeout 300 python tools/sweep_k1.py 64x64 32768 5 >> $out/sweep_k1.log 2>&1
  grep "best" -B 30 $out/sweep_k1.log
fi

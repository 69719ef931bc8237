#!/bin/bash
# Knob sweep of the persistent K1 (epilogue warps, TMA stages) after the parity tests.
# usage (under gpurun): bash tools/gpu_sweep.sh <tag>
tag=${1:-ab3}
out=gpurun_out/$tag
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $out/pytest_gpu.log
tail -12 $out/pytest_gpu.log
for ew in 0 1 2; do
  for dt in f32 bf16; do
    OKP_STREAM_EPILOGUE_WARPS=$ew timeout 120 python tools/bench_k1.py 180x320 4096 10 $dt >> $out/bench_k1.log 2>&1
    OKP_STREAM_EPILOGUE_WARPS=$ew timeout 120 python tools/bench_k1.py 64x64 32768 10 $dt >> $out/bench_k1.log 2>&1
  done
done
for ns in 3 5 6; do
OKP_STRIP_STAGES=$ns timeout 120 python tools/bench_k1.py 180x320 4096 10 f32 >> $out/bench_k1.log 2>&1
OKP_STRIP_STAGES=$ns timeout 120 python tools/bench_k1.py 64x64 32768 10 f32 >> $out/bench_k1.log 2>&1
done
cat $out/bench_k1.log

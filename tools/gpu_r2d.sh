#!/bin/bash
# Parity tests on the shipped build, then tile-vs-stream and split-vs-single-pass timings (tuning build for the stream kernel).
tag=${1:-r2d}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > $out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 300 > $out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $out/pytest_gpu.log
tail -25 $out/pytest_gpu.log
for dt in f32 bf16; do
  timeout 200 python tools/bench_k1.py 180x320 4096 10 $dt >> $out/bench_k1.log 2>&1
  timeout 200 python tools/bench_k1.py 180x320 4096 10 $dt single >> $out/bench_k1.log 2>&1
  timeout 200 python tools/bench_k1.py 64x64 32768 10 $dt lean >> $out/bench_k1.log 2>&1
  timeout 200 python tools/bench_k1.py 64x64 32768 10 $dt lean single >> $out/bench_k1.log 2>&1
done
grep -v Warn $out/bench_k1.log
export OKP_TUNING_LIBRARY=$PWD/object_keypoints_b200/libokp_tuning.so
run() { echo "== $*" >> $out/sweep.log; env "$@" timeout 200 python tools/bench_k1.py $SHAPE $FRAMES 10 $DT $EXTRA 2>&1 | grep -v Warning >> $out/sweep.log; }
DT=f32
SHAPE=180x320 FRAMES=4096 EXTRA=""
run OKP_PEAKS_TILE=0
run OKP_TILE_STAGES=3 OKP_TILE_SMEM_KB=220
run OKP_TILE_COMPUTE_WARPS=6
run OKP_TILE_COMPUTE_WARPS=8 OKP_TILE_TH=20
run OKP_TILE_TH=20 OKP_TILE_STAGES=3
run OKP_TILE_TW=80 OKP_TILE_TH=36 OKP_TILE_STAGES=3 OKP_TILE_SMEM_KB=72
SHAPE=64x64 FRAMES=32768 EXTRA="lean"
run OKP_PEAKS_TILE=0
run OKP_TILE_STAGES=3 OKP_TILE_SMEM_KB=110
run OKP_TILE_MS=2 OKP_TILE_SMEM_KB=110
run OKP_TILE_MS=2 OKP_TILE_SMEM_KB=110 OKP_TILE_COMPUTE_WARPS=6
run OKP_TILE_COMPUTE_WARPS=2 OKP_TILE_SMEM_KB=56
DT=bf16
run OKP_PEAKS_TILE=0
run OKP_TILE_MS=2 OKP_TILE_SMEM_KB=72
cat $out/sweep.log

#!/bin/bash
# parity + 64x64 plan variants (edge-only batches, 224 threads, 3 stages), grouping with 16 lanes per frame
tag=${1:-r2i}
out=gpurun_out/$tag
mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 300 > $out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $out/pytest_gpu.log
tail -4 $out/pytest_gpu.log
for dt in f32 bf16; do
  timeout 200 python tools/bench_k1.py 64x64 32768 10 $dt lean >> $out/bench_k1.log 2>&1
done
grep -v Warn $out/bench_k1.log
export OKP_TUNING_LIBRARY=$PWD/object_keypoints_b200/libokp_tuning.so
run() { echo "== $*" >> $out/sweep.log; env "$@" timeout 200 python tools/bench_k1.py $SHAPE $FRAMES 10 $DT $EXTRA 2>&1 | grep -v Warning >> $out/sweep.log; }
SHAPE=64x64 FRAMES=32768 EXTRA="lean"
for DT in f32 bf16; do
run OKP_GROUP_LANES=32
run OKP_STRIP_THREADS=224 OKP_STRIP_STAGES=3
run OKP_STRIP_THREADS=224 OKP_STRIP_STAGES=3 OKP_STREAM_EDGE_ONLY=1
run OKP_STRIP_THREADS=224 OKP_STRIP_STAGES=2 OKP_STREAM_EDGE_ONLY=1
run OKP_STRIP_THREADS=224 OKP_STRIP_STAGES=3 OKP_STRIP_SMEM_KB=100
run OKP_STREAM_EDGE_ONLY=1
done
sed 's/env={[^}]*}//' $out/sweep.log | cut -c1-200

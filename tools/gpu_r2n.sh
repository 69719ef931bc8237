#!/bin/bash
tag=${1:-r2n}
out=gpurun_out/$tag
mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 300 > $out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $out/pytest_gpu.log
tail -4 $out/pytest_gpu.log
timeout 200 python tools/bench_k1.py 180x320 4096 10 f32 >> $out/bench_k1.log 2>&1
timeout 200 python tools/bench_k1.py 64x64 32768 10 f32 lean >> $out/bench_k1.log 2>&1
timeout 200 python tools/bench_k1.py 180x320 512 20 f32 >> $out/bench_k1.log 2>&1
grep -v Warn $out/bench_k1.log
timeout 900 python bench.py > $out/bench.json 2> $out/bench.err; tail -3 $out/bench.err; cut -c1-400 $out/bench.json
timeout 600 python tools/bench_secondary.py > $out/secondary.json 2> $out/secondary.err; cut -c1-900 $out/secondary.json
export OKP_TUNING_LIBRARY=$PWD/object_keypoints_b200/libokp_tuning.so
run() { echo "== $*" >> $out/sweep.log; env "$@" timeout 200 python tools/bench_k1.py $SHAPE $FRAMES 20 $DT $EXTRA 2>&1 | grep -v Warning >> $out/sweep.log; }
DT=f32 SHAPE=180x320 EXTRA=""
for FRAMES in 512 1024 4096; do
run OKP_STRIP_THREADS=96
run OKP_STRIP_THREADS=96 OKP_STRIP_SMEM_KB=72
run OKP_STRIP_THREADS=192
done
sed 's/env={[^}]*}//' $out/sweep.log | cut -c1-200

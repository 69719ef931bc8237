#!/bin/bash
tag=${1:-r2k}
out=gpurun_out/$tag
mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 300 > $out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $out/pytest_gpu.log
tail -4 $out/pytest_gpu.log
for dt in f32 bf16; do
  timeout 200 python tools/bench_k1.py 180x320 4096 10 $dt >> $out/bench_k1.log 2>&1
  timeout 200 python tools/bench_k1.py 64x64 32768 10 $dt lean >> $out/bench_k1.log 2>&1
done
timeout 200 python tools/bench_k1.py 180x320 512 20 f32 >> $out/bench_k1.log 2>&1
timeout 200 python tools/bench_k1.py 180x320 1024 20 f32 >> $out/bench_k1.log 2>&1
grep -v Warn $out/bench_k1.log
timeout 600 python tools/bench_secondary.py > $out/secondary.json 2> $out/secondary.err; cut -c1-1500 $out/secondary.json

#!/bin/bash
# Round artefacts at N = 1: parity tests, the bench line (ours + reference arm), K1 / grouping timings.
tag=${1:-r2f}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > $out/gpu.txt; nproc >> $out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 300 > $out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $out/pytest_gpu.log
tail -4 $out/pytest_gpu.log
timeout 900 python bench.py --impl reference > $out/bench_reference.json 2> $out/bench_reference.err; tail -c 600 $out/bench_reference.json
timeout 1200 python bench.py > $out/bench.json 2> $out/bench.err; tail -5 $out/bench.err; cat $out/bench.json
for dt in f32 bf16; do
  timeout 200 python tools/bench_k1.py 180x320 4096 10 $dt >> $out/bench_k1.log 2>&1
  timeout 200 python tools/bench_k1.py 64x64 32768 10 $dt lean >> $out/bench_k1.log 2>&1
done
grep -v Warn $out/bench_k1.log

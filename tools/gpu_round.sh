#!/bin/bash
# One GPU call at N = 1: parity tests, the bench line of both arms, K1 / grouping timings, per-frame latency.
# usage (under gpurun): bash tools/gpu_round.sh <tag>
tag=${1:-round}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > $out/gpu.txt; nproc >> $out/gpu.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > $out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $out/pytest_gpu.log
tail -4 $out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; tail -1 $out/smoke.log
timeout 600 python bench.py --impl reference > $out/bench_reference.json 2> $out/bench_reference.err; cut -c1-300 $out/bench_reference.json
timeout 900 python bench.py > $out/bench.json 2> $out/bench.err; tail -3 $out/bench.err; cut -c1-300 $out/bench.json
for dt in f32 bf16; do
  timeout 200 python tools/bench_k1.py 180x320 4096 10 $dt >> $out/bench_k1.log 2>&1
  timeout 200 python tools/bench_k1.py 64x64 32768 10 $dt lean >> $out/bench_k1.log 2>&1
done
grep -v Warn $out/bench_k1.log
timeout 200 python tools/bench_per_frame.py > $out/per_frame.json 2> $out/per_frame.err; cat $out/per_frame.json

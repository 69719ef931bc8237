#!/bin/bash
# the bench line alone (N = 1): bash tools/gpu_bench_only.sh <tag> [bench flags]
out=gpurun_out/$1; mkdir -p $out; shift
timeout 600 python bench.py "$@" > $out/bench.json 2> $out/bench.err; tail -2 $out/bench.err; cut -c1-200 $out/bench.json

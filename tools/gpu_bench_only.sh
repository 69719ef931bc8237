#!/bin/bash
out=gpurun_out/$1; mkdir -p $out
timeout 600 python bench.py > $out/bench.json 2> $out/bench.err; tail -2 $out/bench.err; cut -c1-200 $out/bench.json

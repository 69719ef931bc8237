#!/bin/bash
# Final artefacts of a round (under gpurun): the bench line, the reference arm, the ncu launch list of the bench command.
tag=${1:-final}
out=gpurun_out/$tag
mkdir -p $out
timeout 600 python bench.py --steps 20 --warmup 3 > $out/bench.json 2> $out/bench.err; tail -2 $out/bench.err; cut -c1-300 $out/bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_reference.json 2>> $out/bench.err; cut -c1-200 $out/bench_reference.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:okp_ -c 80 --csv --log-file $out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-frames 256 --e2e-steps 1 > $out/bench_under_ncu.log 2>&1
ls $out

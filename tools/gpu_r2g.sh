#!/bin/bash
# N = 2: weak + strong scaling lines, exchange check on hardware.
tag=${1:-r2g}; n=${2:-2}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name --format=csv,noheader > $out/gpu.txt; nproc >> $out/gpu.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 3 > $out/bench_n$n.json 2> $out/bench_n$n.err
tail -5 $out/bench_n$n.err; cat $out/bench_n$n.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 tools/check_exchange.py > $out/check_exchange.log 2>&1; tail -5 $out/check_exchange.log
timeout 600 python -m pytest tests -m gpu -q -x --timeout 300 -k "shard or exchange or record" > $out/pytest_multi.log 2>&1; tail -3 $out/pytest_multi.log

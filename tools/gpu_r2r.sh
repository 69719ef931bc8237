#!/bin/bash
tag=${1:-r2r}; n=${2:-2}
out=gpurun_out/$tag
mkdir -p $out
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 tools/check_exchange.py > $out/check_exchange.log 2>&1; grep "^rank" $out/check_exchange.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 3 > $out/bench_n$n.json 2> $out/bench_n$n.err
tail -3 $out/bench_n$n.err | cut -c1-300; cut -c1-300 $out/bench_n$n.json

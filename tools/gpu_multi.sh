#!/bin/bash
# N GPUs: the bench line (weak scaling, strong scaling block, e2e).
tag=${1:-multi}; n=${2:-8}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name --format=csv,noheader | head -1 > $out/gpu.txt; nproc >> $out/gpu.txt
timeout ${3:-200} python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 3 > $out/bench_n$n.json 2> $out/bench_n$n.err
tail -5 $out/bench_n$n.err | cut -c1-300; cut -c1-1500 $out/bench_n$n.json

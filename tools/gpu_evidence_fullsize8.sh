#!/bin/bash
# BASELINE.json configs[2] and configs[3] at their FULL sizes on 8 GPUs of one box: curly-hair 1024x1024 at 1024 spp (8 x 128 sample
# indices) and furball 2048x2048 at 256 spp, maxDepth 32 (8 x 32), sample-range sharded with one NCCL film reduce.
# usage: gpurun --gpus 8 --timeout 900 -- 'bash tools/gpu_evidence_fullsize8.sh <tag>'
tag=${1:-r1}
out=gpurun_out
mkdir -p $out
run() { name=$1; shift; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu "$@" > $out/bench_${name}_n8_$tag.json 2> $out/bench_${name}_n8_$tag.err; echo "$name rc=$?"
    tail -1 $out/bench_${name}_n8_$tag.json | python tools/dev/summ.py $name | cut -c1-400; }
run c3_curly_hair_1024spp --scene curly-hair --spp 128
run c4_furball_256spp --scene furball --spp 32

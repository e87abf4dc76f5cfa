#!/bin/bash
# round-2 GPU run 29 (8 GPUs): the driver's scaling line at N = 8 on the final library (strong scaling of the fixed 64-spp job, both arms)
out=gpurun_out; mkdir -p $out
export CUDAPATH_SCENE_CACHE=/tmp/cudapath_scene_cache
nvidia-smi -L | wc -l | tee $out/g29_gpus.txt
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 3 --warmup 3 > $out/g29_bench_n$N.json 2> $out/g29_bench_n$N.err; echo "bench n$N rc=$?"
tail -2 $out/g29_bench_n$N.err | cut -c1-300; python tools/dev/summ.py n$N < $out/g29_bench_n$N.json | cut -c1-500

#!/bin/bash
# round-2 GPU run 15 (1 GPU): pre-split cap 8 / 12 / 16 / 24 (cubic Morton grid on), compact pixel shards (128 / 256 pixel blocks) of an 8-way split
out=gpurun_out; mkdir -p $out
export CUDAPATH_SCENE_CACHE=/tmp/cudapath_scene_cache
bash tools/dev/ab.sh 2 env:CUDAPATH_MAX_SPLIT=8 env:CUDAPATH_MAX_SPLIT=12 env:CUDAPATH_MAX_SPLIT=16 env:CUDAPATH_MAX_SPLIT=24 -- 2>&1 | cut -c1-330 | tee $out/g15_ab_split_hair_curl.log
bash tools/dev/ab.sh 1 env:CUDAPATH_MAX_SPLIT=8 env:CUDAPATH_MAX_SPLIT=16 -- --scene furball --spp 16 2>&1 | cut -c1-330 | tee $out/g15_ab_split_furball.log
bash tools/dev/ab.sh 1 env:CUDAPATH_MAX_SPLIT=8 env:CUDAPATH_MAX_SPLIT=16 -- --scene straight-hair 2>&1 | cut -c1-330 | tee $out/g15_ab_split_straight.log
bash tools/dev/ab.sh 1 env:CUDAPATH_MAX_SPLIT=8 env:CUDAPATH_MAX_SPLIT=16 -- --scene curly-hair --spp 16 2>&1 | cut -c1-330 | tee $out/g15_ab_split_curly.log
for b in 128 256; do for k in 0 1 2 3 4 5 6 7; do
  CUDAPATH_SHARD_BLOCK=$b timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e --shard-test $k/8 2>&1 | tail -1 | python tools/dev/summ.py "block $b shard $k/8" | cut -c1-110
done; done 2>&1 | tee $out/g15_shard_blocks.log

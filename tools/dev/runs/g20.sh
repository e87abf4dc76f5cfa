#!/bin/bash
# round-2 GPU run 20 (1 GPU): 4-byte leaf index (leaf records 2 GB -> 256 MB) on the split-16 / split-24 trees; leaf opening cost
out=gpurun_out; mkdir -p $out
export CUDAPATH_SCENE_CACHE=/tmp/cudapath_scene_cache
bash tools/dev/ab.sh 2 base _li env:CUDAPATH_MAX_SPLIT=24 _li+CUDAPATH_MAX_SPLIT=24 env:CUDAPATH_LEAF_SPLIT_COST=0.5 env:CUDAPATH_LEAF_SPLIT_COST=2 env:CUDAPATH_LEAF_SPLIT_COST=-1 -- 2>&1 | cut -c1-200 | tee $out/g20_ab_leaf_index_hair_curl.log
bash tools/dev/ab.sh 1 base _li -- --scene furball --spp 16 2>&1 | cut -c1-200 | tee $out/g20_ab_leaf_index_furball.log
bash tools/dev/ab.sh 1 base _li -- --scene straight-hair 2>&1 | cut -c1-200 | tee $out/g20_ab_leaf_index_straight.log

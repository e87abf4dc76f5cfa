#!/bin/bash
# round-2 GPU run 16 (1 GPU): two concurrent half-waves for jobs that fit one wave (CUDAPATH_DUAL), pre-split cap x leaf size, e2e at split 8 / 16 / 24
out=gpurun_out; mkdir -p $out
export CUDAPATH_SCENE_CACHE=/tmp/cudapath_scene_cache
CUDAPATH_DUAL=1 timeout 900 python -m pytest tests -m gpu -q -x -k "pixel_shards or additiv or run_to_run or replay or seeds or smoke or cancel or progress or xml_scene or sobol" > $out/g16_pytest_dual.log 2>&1; echo "pytest rc=$?" | tee -a $out/g16_pytest_dual.log
tail -4 $out/g16_pytest_dual.log
bash tools/dev/ab.sh 2 base env:CUDAPATH_DUAL=1 -- --spp 8 2>&1 | cut -c1-220 | tee $out/g16_ab_dual_8spp.log
bash tools/dev/ab.sh 2 base env:CUDAPATH_DUAL=1 -- 2>&1 | cut -c1-220 | tee $out/g16_ab_dual_64spp.log
bash tools/dev/ab.sh 2 base env:CUDAPATH_DUAL=1 -- --scene straight-hair 2>&1 | cut -c1-220 | tee $out/g16_ab_dual_straight.log
bash tools/dev/ab.sh 1 base env:CUDAPATH_DUAL=1 -- --scene furball --spp 16 2>&1 | cut -c1-220 | tee $out/g16_ab_dual_furball.log
bash tools/dev/ab.sh 1 base env:CUDAPATH_DUAL=1 -- --shard-test 3/8 2>&1 | cut -c1-220 | tee $out/g16_ab_dual_shard.log
bash tools/dev/ab.sh 1 env:CUDAPATH_MAX_SPLIT=16 _l6+CUDAPATH_MAX_SPLIT=16 _l3+CUDAPATH_MAX_SPLIT=16 _l2+CUDAPATH_MAX_SPLIT=16 env:CUDAPATH_MAX_SPLIT=24 _l3+CUDAPATH_MAX_SPLIT=24 _l2+CUDAPATH_MAX_SPLIT=24 env:CUDAPATH_MAX_SPLIT=32 _l2+CUDAPATH_MAX_SPLIT=32 -- 2>&1 | cut -c1-330 | tee $out/g16_ab_split_leaf.log
for sp in 8 16 24; do
  CUDAPATH_MAX_SPLIT=$sp timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 | python tools/dev/summ.py "e2e split $sp" | cut -c1-400
done 2>&1 | tee $out/g16_e2e_split.log

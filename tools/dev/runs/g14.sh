#!/bin/bash
# round-2 GPU run 14 (1 GPU): cubic Morton grid A/B on the four scenes, pixel-shard block size (balance of an 8-way split), max-split 16, the shard additivity test
out=gpurun_out; mkdir -p $out
export CUDAPATH_SCENE_CACHE=/tmp/cudapath_scene_cache
timeout 600 python -m pytest tests -m gpu -q -k "pixel_shards or xml_scene_roundtrip or closest_hit_prim_ids" > $out/g14_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/g14_pytest.log
tail -5 $out/g14_pytest.log
bash tools/dev/ab.sh 2 base env:CUDAPATH_MORTON_CUBE=1 -- 2>&1 | tee $out/g14_ab_cube_hair_curl.log
bash tools/dev/ab.sh 2 base env:CUDAPATH_MORTON_CUBE=1 -- --scene furball --spp 16 2>&1 | tee $out/g14_ab_cube_furball.log
bash tools/dev/ab.sh 2 base env:CUDAPATH_MORTON_CUBE=1 -- --scene straight-hair 2>&1 | tee $out/g14_ab_cube_straight.log
bash tools/dev/ab.sh 2 base env:CUDAPATH_MORTON_CUBE=1 -- --scene curly-hair --spp 16 2>&1 | tee $out/g14_ab_cube_curly.log
bash tools/dev/ab.sh 1 base env:CUDAPATH_MORTON_CUBE=1 -- --max-split 16 2>&1 | tee $out/g14_ab_split16.log
for b in 32 8 16; do for k in 0 3 4 7; do
  CUDAPATH_SHARD_BLOCK=$b timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e --shard-test $k/8 2>&1 | tail -1 | python tools/dev/summ.py "block $b shard $k/8" | cut -c1-110
done; done 2>&1 | tee $out/g14_shard_blocks.log

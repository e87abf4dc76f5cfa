#!/bin/bash
# round-2 GPU run 5: the GPU suite (not stopping at the first failure), per-shard times of the 8-way pixel sharding, config 5
out=gpurun_out; mkdir -p $out
export CUDAPATH_SCENE_CACHE=/tmp/cudapath_scene_cache
timeout 1800 python -m pytest tests -m gpu -q --durations=10 > $out/g5_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/g5_pytest.log
tail -25 $out/g5_pytest.log
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > /dev/null 2>&1
for a in "--shard-test 0/8" "--shard-test 1/8" "--shard-test 2/8" "--shard-test 3/8" "--shard-test 4/8" "--shard-test 5/8" "--shard-test 6/8" "--shard-test 7/8" "--spp 8" "" "--shard-test 0/2" "--shard-test 1/2" "--shard-test 0/4" "--shard-test 2/4"; do
    timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e $a 2>&1 | tail -1 | python tools/dev/summ.py "[$a]" | cut -c1-110
done | tee $out/g5_shard.log
timeout 600 python bench.py --config 5 --steps 3 > $out/g5_config5.json 2> $out/g5_config5.err; echo "config5 rc=$?"; cut -c1-200 $out/g5_config5.json

#!/bin/bash
# round-2 GPU run 4: pixel-block vs sample sharding on one GPU (shard 0 of 8), GPU suite with the new tests, config 5
out=gpurun_out; mkdir -p $out
export CUDAPATH_SCENE_CACHE=/tmp/cudapath_scene_cache
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > $out/g4_sanity.json 2> $out/g4_sanity.err; echo "sanity rc=$?"
python tools/dev/summ.py sanity < $out/g4_sanity.json | cut -c1-400
v=$(python -c "import json,sys; print(int(json.loads(open('$out/g4_sanity.json').read().strip().splitlines()[-1])['value']))" 2>/dev/null || echo 0)
if [ "$v" -lt 100 ]; then echo "sanity value $v < 100: aborting"; tail -5 $out/g4_sanity.err; exit 1; fi
# what one of 8 GPUs would do under either sharding of the fixed 1024x1024x64 job
for r in 1 2 3; do
  for a in "--spp 8" "--shard-test 0/8" "--shard-test 3/8" ""; do
    timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e $a 2>&1 | tail -1 | python tools/dev/summ.py "r$r [$a]" | cut -c1-110
  done
done | tee $out/g4_shard.log
timeout 1800 python -m pytest tests -m gpu -q -x --durations=10 > $out/g4_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/g4_pytest.log
tail -18 $out/g4_pytest.log
timeout 600 python bench.py --config 5 --steps 3 > $out/g4_config5.json 2> $out/g4_config5.err; echo "config5 rc=$?"; cut -c1-300 $out/g4_config5.json

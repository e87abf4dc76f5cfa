#!/bin/bash
# round-2 GPU run 3: fast math + dual-stream half-jobs + SAH leaf decision (A/B by environment variable), the GPU suite, smoke, config 5
out=gpurun_out; mkdir -p $out
export CUDAPATH_SCENE_CACHE=/tmp/cudapath_scene_cache
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > $out/g3_sanity.json 2> $out/g3_sanity.err; echo "sanity rc=$?"
python tools/dev/summ.py sanity < $out/g3_sanity.json | cut -c1-400
v=$(python -c "import json,sys; print(int(json.loads(open('$out/g3_sanity.json').read().strip().splitlines()[-1])['value']))" 2>/dev/null || echo 0)
if [ "$v" -lt 100 ]; then echo "sanity value $v < 100: aborting"; tail -5 $out/g3_sanity.err; exit 1; fi
( bash tools/dev/ab.sh 2 base env:CUDAPATH_DUAL=0 env:CUDAPATH_MATH=strict env:CUDAPATH_LEAF_SPLIT_COST=0.5 env:CUDAPATH_LEAF_SPLIT_COST=1 env:CUDAPATH_LEAF_SPLIT_COST=2 env:CUDAPATH_LEAF_SPLIT_COST=4 -- 2>&1 ) | tee $out/g3_ab_64spp.log
( bash tools/dev/ab.sh 3 base env:CUDAPATH_DUAL=0 -- --spp 8 2>&1 ) | tee $out/g3_ab_8spp.log
( bash tools/dev/ab.sh 1 base env:CUDAPATH_LEAF_SPLIT_COST=0.5 env:CUDAPATH_LEAF_SPLIT_COST=1 env:CUDAPATH_LEAF_SPLIT_COST=2 -- --scene furball --spp 16 2>&1 ) | tee $out/g3_ab_furball.log
( bash tools/dev/ab.sh 1 base env:CUDAPATH_LEAF_SPLIT_COST=0.5 env:CUDAPATH_LEAF_SPLIT_COST=1 env:CUDAPATH_LEAF_SPLIT_COST=2 -- --scene straight-hair 2>&1 ) | tee $out/g3_ab_straight.log
( bash tools/dev/ab.sh 1 base env:CUDAPATH_LEAF_SPLIT_COST=1 -- --scene curly-hair --spp 16 2>&1 ) | tee $out/g3_ab_curly.log
timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 > $out/g3_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/g3_pytest.log
tail -15 $out/g3_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/g3_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 $out/g3_smoke.log
timeout 600 python bench.py --config 5 --steps 3 > $out/g3_config5.json 2> $out/g3_config5.err; echo "config5 rc=$?"; cut -c1-600 $out/g3_config5.json

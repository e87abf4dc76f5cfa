#!/bin/bash
# round-2 GPU run 2: sanity, interleaved A/B of the kernel variants, then the GPU suite
out=gpurun_out; mkdir -p $out
export CUDAPATH_SCENE_CACHE=/tmp/cudapath_scene_cache
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > $out/g2_sanity.json 2> $out/g2_sanity.err; echo "sanity rc=$?"
python tools/dev/summ.py sanity < $out/g2_sanity.json | cut -c1-400
v=$(python -c "import json,sys; print(int(json.loads(open('$out/g2_sanity.json').read().strip().splitlines()[-1])['value']))" 2>/dev/null || echo 0)
if [ "$v" -lt 100 ]; then echo "sanity value $v < 100: aborting"; tail -5 $out/g2_sanity.err; exit 1; fi
( bash tools/dev/ab.sh 2 base _r1 _h0 _s0 _h0s0 _l1 _l2 -- 2>&1 ) | tee $out/g2_ab_64spp.log
( bash tools/dev/ab.sh 2 base _r1 _l1 env:CUDAPATH_RUNAHEAD_MAX=0 -- --spp 8 2>&1 ) | tee $out/g2_ab_8spp.log
( bash tools/dev/ab.sh 1 base _r1 _l1 -- --scene furball 2>&1 ) | tee $out/g2_ab_furball.log
( bash tools/dev/ab.sh 1 base _r1 _l1 -- --scene straight-hair 2>&1 ) | tee $out/g2_ab_straight.log
timeout 1200 python -m pytest tests -m gpu -q -x --durations=8 > $out/g2_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/g2_pytest.log
tail -15 $out/g2_pytest.log

#!/bin/bash
# round-2 GPU run 26 (1 GPU): register cap / shared-memory stack depth on the final trees; bench lines of the other BASELINE configs; ncu captures on the split-16 tree
out=gpurun_out; mkdir -p $out
export CUDAPATH_SCENE_CACHE=/tmp/cudapath_scene_cache
bash tools/dev/ab.sh 2 base _mb5 _ss8 _ss16 -- 2>&1 | cut -c1-200 | tee $out/g26_ab_occupancy_hair_curl.log
bash tools/dev/ab.sh 1 base _mb5 _ss8 _ss16 -- --scene furball --spp 16 2>&1 | cut -c1-200 | tee $out/g26_ab_occupancy_furball.log
timeout 300 python bench.py --scene straight-hair --no-cpu > $out/g26_bench_c1_straight_hair.json 2>/dev/null; python tools/dev/summ.py c1 < $out/g26_bench_c1_straight_hair.json | cut -c1-400
timeout 600 python bench.py --scene curly-hair --spp 64 --no-cpu > $out/g26_bench_c3_curly_hair_64spp.json 2>/dev/null; python tools/dev/summ.py c3 < $out/g26_bench_c3_curly_hair_64spp.json | cut -c1-400
timeout 600 python bench.py --scene furball --spp 16 --no-cpu > $out/g26_bench_c4_furball_16spp.json 2>/dev/null; python tools/dev/summ.py c4 < $out/g26_bench_c4_furball_16spp.json | cut -c1-400
tag=r2e
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $out/launches_$tag.csv \
    python bench.py --spp 8 --max-split 16 --steps 1 --warmup 1 --no-cpu --no-e2e > $out/ncu_launches_$tag.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 2 -c 1 -o $out/prof_${tag}_trace \
    python bench.py --spp 8 --max-split 16 --steps 1 --warmup 1 --no-cpu --no-e2e > $out/ncu_full_$tag.log 2>&1; echo "ncu full rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_shade -s 2 -c 1 -o $out/prof_${tag}_shade \
    python bench.py --spp 8 --max-split 16 --steps 1 --warmup 1 --no-cpu --no-e2e > $out/ncu_full_shade_$tag.log 2>&1; echo "ncu full shade rc=$?"

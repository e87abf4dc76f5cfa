#!/bin/bash
# One bench line per BASELINE.json config on one GPU (C3 / C4 at reduced spp: their full sample counts are 1.07 G paths each;
# throughput does not depend on spp once a wave is full).  usage: gpurun --timeout 1500 -- 'bash tools/gpu_evidence_configs.sh <tag>'
tag=${1:-r1}
out=gpurun_out
mkdir -p $out
run() { name=$1; shift; timeout 600 python bench.py "$@" > $out/bench_${name}_$tag.json 2> $out/bench_${name}_$tag.err; echo "$name rc=$?"; python tools/dev/summ.py $name < $out/bench_${name}_$tag.json | cut -c1-400; }
run default
run c1_straight_hair --scene straight-hair
run c3_curly_hair_64spp --scene curly-hair --spp 64
run c4_furball_16spp --scene furball --spp 16

#!/bin/bash
# Per-stage evidence on the GPU box (one GPU): e2e phase breakdown, the config-5 micro-benchmarks (BSDF tuples, rays vs the
# furball BVH) and ncu --set full captures of k_shade, k_shadow and the two BSDF batch kernels.
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_evidence_stages.sh <tag>'
tag=${1:-r1}
out=gpurun_out
mkdir -p $out
timeout 300 python tools/dev/e2e_phases.py hair-curl > $out/e2e_phases_$tag.log 2>&1; echo "e2e phases rc=$?"; tail -6 $out/e2e_phases_$tag.log
timeout 600 python tools/microbench.py > $out/microbench_$tag.jsonl 2> $out/microbench_$tag.err; echo "microbench rc=$?"
cut -c1-330 $out/microbench_$tag.jsonl
for k in k_shade k_shadow; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o $out/prof_${tag}_$k \
      python bench.py --spp 8 --steps 1 --warmup 1 --no-cpu --no-e2e > $out/ncu_full_${tag}_$k.log 2>&1; echo "ncu $k rc=$?"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_bsdf_ -s 3 -c 2 -o $out/prof_${tag}_k_bsdf \
    python tools/microbench.py --log2n 24 --reps 1 --bsdf-only > $out/ncu_full_${tag}_k_bsdf.log 2>&1; echo "ncu k_bsdf rc=$?"

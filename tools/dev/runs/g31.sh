#!/bin/bash
# round-2 GPU run 31 (1 GPU): cold per-lane state of trace_persistent in shared memory (ray direction, hit record, per-shape interval) at 6 / 7 / 8 resident CTAs
out=gpurun_out; mkdir -p $out
export CUDAPATH_SCENE_CACHE=/tmp/cudapath_scene_cache
bash tools/dev/ab.sh 2 base _c6 _c7 _c8 -- 2>&1 | cut -c1-200 | tee $out/g31_ab_cold_smem_hair_curl.log
bash tools/dev/ab.sh 1 base _c6 _c7 _c8 -- --scene furball --spp 16 2>&1 | cut -c1-200 | tee $out/g31_ab_cold_smem_furball.log
bash tools/dev/ab.sh 1 base _c7 -- --spp 8 2>&1 | cut -c1-200 | tee $out/g31_ab_cold_smem_8spp.log

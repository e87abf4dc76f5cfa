#!/bin/bash
# round-2 GPU run 19 (1 GPU): phase thresholds of trace_persistent re-tuned on the split-16 trees (descent exit, FP64 hold, refill), run-ahead default vs exact sizing
out=gpurun_out; mkdir -p $out
export CUDAPATH_SCENE_CACHE=/tmp/cudapath_scene_cache
bash tools/dev/ab.sh 2 base env:CUDAPATH_RUNAHEAD_MAX=4194304 _d4 _d12 _d16 _h4 _h12 _h16 _r4 _r12 _r16 -- 2>&1 | cut -c1-200 | tee $out/g19_ab_thresholds_hair_curl.log
bash tools/dev/ab.sh 1 base _d4 _d12 _d16 _h4 _h12 _h16 _r4 _r12 _r16 -- --scene furball --spp 16 2>&1 | cut -c1-200 | tee $out/g19_ab_thresholds_furball.log

#!/bin/bash
# round-2 GPU run 30 (1 GPU): how does the rate of a small job (8 spp) depend on the number of persistent CTAs?
out=gpurun_out; mkdir -p $out
export CUDAPATH_SCENE_CACHE=/tmp/cudapath_scene_cache
bash tools/dev/ab.sh 1 base env:CUDAPATH_TRACE_GRID_FRAC=0.834 env:CUDAPATH_TRACE_GRID_FRAC=0.667 env:CUDAPATH_TRACE_GRID_FRAC=0.5 env:CUDAPATH_TRACE_GRID_FRAC=0.334 -- --spp 8 2>&1 | cut -c1-200 | tee $out/g30_grid_frac_8spp.log
bash tools/dev/ab.sh 1 base env:CUDAPATH_TRACE_GRID_FRAC=0.667 -- 2>&1 | cut -c1-200 | tee $out/g30_grid_frac_64spp.log

#!/bin/bash
# round-2 GPU run 23 (1 GPU): build kernels after the refit prefetch / warp-aggregated collapse: parity subset, launch list of the build, bench
out=gpurun_out; mkdir -p $out
export CUDAPATH_SCENE_CACHE=/tmp/cudapath_scene_cache
timeout 900 python -m pytest tests -m gpu -q -x -k "collapse or job_size or closest_hit or full_size_ray_batch or mesh or teapot or any_hit or scene_bounds or kdtree or pixel_shards" > $out/g23_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/g23_pytest.log
tail -3 $out/g23_pytest.log
bash tools/dev/runs/g22.sh 2>&1 | grep -v "^\[cudapath\]" | sed 's/g22/g22b/'
bash tools/dev/ab.sh 2 base -- 2>&1 | cut -c1-200

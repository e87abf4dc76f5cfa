#!/bin/bash
# round-2 GPU run 11: sobol mode after the k_splat fix, the rectangle test
out=gpurun_out; mkdir -p $out
export CUDAPATH_SCENE_CACHE=/tmp/cudapath_scene_cache
timeout 900 python -m pytest tests -m gpu -q -k "sobol or rectangle_and_mesh or film_splat or render_matches_oracle" > $out/g11_pytest_new.log 2>&1; echo "new rc=$?" | tee -a $out/g11_pytest_new.log
tail -40 $out/g11_pytest_new.log | cut -c1-250

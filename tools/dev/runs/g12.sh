#!/bin/bash
# round-2 GPU run 12: mirror / teapot family tests; A/B of 5 / 6 resident k_shade_fast CTAs (96 / 80 registers) against 4 (120)
out=gpurun_out; mkdir -p $out
export CUDAPATH_SCENE_CACHE=/tmp/cudapath_scene_cache
timeout 900 python -m pytest tests -m gpu -q -k "teapot or plastic_checkerboard or rectangle_and_mesh or dielectric" > $out/g12_pytest_new.log 2>&1; echo "new rc=$?" | tee -a $out/g12_pytest_new.log
tail -30 $out/g12_pytest_new.log | cut -c1-250
bash tools/dev/ab.sh 2 base _sh5 _sh6 -- 2>&1 | tee $out/g12_ab_shade.log

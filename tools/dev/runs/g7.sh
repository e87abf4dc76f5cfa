#!/bin/bash
# round-2 GPU run 7: two rays per lane (cp_traverse2.cuh): GPU suite on that library, then A/B against the base at 6 / 5 / 4 resident CTAs
out=gpurun_out; mkdir -p $out
export CUDAPATH_SCENE_CACHE=/tmp/cudapath_scene_cache
P=$PWD/cs184-final-project-mitsuba0.5_b200
CUDAPATH_LIB=$P/libcudapath_r2.so timeout 1500 python -m pytest tests -m gpu -q -x > $out/g7_pytest_r2.log 2>&1; echo "pytest rc=$?" | tee -a $out/g7_pytest_r2.log
tail -5 $out/g7_pytest_r2.log
bash tools/dev/ab.sh 2 base _r2 _r2b5 _r2b4 -- 2>&1 | tee $out/g7_ab_hair_curl.log
bash tools/dev/ab.sh 1 base _r2 _r2b5 -- --scene furball 2>&1 | tee $out/g7_ab_furball.log
bash tools/dev/ab.sh 1 base _r2 _r2b5 -- --scene straight-hair 2>&1 | tee $out/g7_ab_straight.log

#!/bin/bash
# round-2 GPU run 8: the teapot-scene plugins (plastic, checkerboard, twosided, rectangle, mesh uv) first, then the whole GPU suite, then a bench line
out=gpurun_out; mkdir -p $out
export CUDAPATH_SCENE_CACHE=/tmp/cudapath_scene_cache
timeout 900 python -m pytest tests -m gpu -q -k "plastic_checkerboard or rectangle_and_mesh or teapot_scene" > $out/g8_pytest_new.log 2>&1; echo "new rc=$?" | tee -a $out/g8_pytest_new.log
tail -40 $out/g8_pytest_new.log
timeout 1500 python -m pytest tests -m gpu -q > $out/g8_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/g8_pytest.log
tail -8 $out/g8_pytest.log
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e 2>&1 | tail -1 | python tools/dev/summ.py "bench" | cut -c1-200 | tee $out/g8_bench.log

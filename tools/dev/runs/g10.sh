#!/bin/bash
# round-2 GPU run 10: sampler-faithful mode (sobol), CLI -r / exr, the fixed tests; then the whole GPU suite
out=gpurun_out; mkdir -p $out
export CUDAPATH_SCENE_CACHE=/tmp/cudapath_scene_cache
timeout 900 python -m pytest tests -m gpu -q -k "sobol or native_cli or rectangle_and_mesh or marschner_fixed_mode or plastic_checkerboard" > $out/g10_pytest_new.log 2>&1; echo "new rc=$?" | tee -a $out/g10_pytest_new.log
tail -40 $out/g10_pytest_new.log | cut -c1-250
timeout 1500 python -m pytest tests -m gpu -q > $out/g10_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/g10_pytest.log
tail -8 $out/g10_pytest.log

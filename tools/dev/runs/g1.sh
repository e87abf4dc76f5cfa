#!/bin/bash
# round-2 GPU run 1: parity of the merged trace launch + device-driven bounce loop, then interleaved A/B of the kernel variants
out=gpurun_out; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > $out/g1_gpu.txt
timeout 1200 python -m pytest tests -m gpu -q -x > $out/g1_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/g1_pytest.log
tail -15 $out/g1_pytest.log
( bash tools/dev/ab.sh 2 base _r1 _h0 _s0 _h0s0 _h16 _s20 -- 2>&1 ) | tee $out/g1_ab_64spp.log
( bash tools/dev/ab.sh 2 base _r1 _h0s0 -- --spp 8 2>&1 ) | tee $out/g1_ab_8spp.log
( bash tools/dev/ab.sh 1 base _r1 -- --scene furball 2>&1 ) | tee $out/g1_ab_furball.log
( bash tools/dev/ab.sh 1 base _r1 -- --scene straight-hair 2>&1 ) | tee $out/g1_ab_straight.log

#!/bin/bash
# round-2 GPU run 6: A/B of (a) 4-byte leaf index instead of the 32-byte leaf record, (b) the FP64 test as a real call at 6 / 7 / 8 resident CTAs
out=gpurun_out; mkdir -p $out
bash tools/dev/ab.sh 2 base _li _ni6 _ni7 _ni8 _lini8 -- 2>&1 | tee $out/g6_ab_hair_curl.log
bash tools/dev/ab.sh 1 base _li _ni8 _lini8 -- --scene furball 2>&1 | tee $out/g6_ab_furball.log

#!/bin/bash
# round-2 GPU run 13: L1 prefetch of the stack entry that is popped next (at push time / also at pop time)
out=gpurun_out; mkdir -p $out
bash tools/dev/ab.sh 2 base _pf1 _pf2 -- 2>&1 | tee $out/g13_ab_prefetch.log
bash tools/dev/ab.sh 1 base _pf1 _pf2 -- --scene furball 2>&1 | tee $out/g13_ab_prefetch_furball.log

#!/bin/bash
# round-2 GPU run 25 (1 GPU): does the nvidia-smi clock sampling perturb the timed steps?  10 steps each with 200 ms / 1000 ms / no sampling
out=gpurun_out; mkdir -p $out
export CUDAPATH_SCENE_CACHE=/tmp/cudapath_scene_cache
for p in 200 0 1000 200 0; do
  BENCH_CLOCK_PERIOD_MS=$p timeout 300 python bench.py --no-cpu --no-e2e --steps 10 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('period $p: value %.1f ms/step %.1f step_render_ms %s clocks %s' % (d['value'], d['ms_per_step'], d['step_render_ms'], d['clocks']))"
done | tee $out/g25_clock_sampling.log

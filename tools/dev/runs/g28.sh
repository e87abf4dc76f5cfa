#!/bin/bash
# round-2 GPU run 28 (1 GPU): render without cudaMemGetInfo: step-time spread (16 renders, 2 x 10 timed steps), then the whole GPU suite
out=gpurun_out; mkdir -p $out
export CUDAPATH_SCENE_CACHE=/tmp/cudapath_scene_cache
timeout 300 python tools/dev/idle_probe.py 16 2>&1 | grep "^render" | tee $out/g28_idle_probe.log
for i in 1 2; do
  timeout 300 python bench.py --no-cpu --steps 10 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); ph=[v for k,v in d['e2e'].items() if k.startswith('phases_s')][0]
print('run $i: value %.1f ms/step %.1f step_render_ms %s | e2e %.1f render %s' % (d['value'], d['ms_per_step'], d['step_render_ms'], d['e2e']['value'], [p[3] for p in ph]))"
done | tee $out/g28_step_spread.log
timeout 1200 python -m pytest tests -m gpu -q --durations=5 > $out/g28_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/g28_pytest.log
tail -4 $out/g28_pytest.log

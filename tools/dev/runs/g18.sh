#!/bin/bash
# round-2 GPU run 18 (1 GPU): where do the 10-20 ms per step of the scene-resident loop go at split 16? per-step device render time, clocks, power
out=gpurun_out; mkdir -p $out
export CUDAPATH_SCENE_CACHE=/tmp/cudapath_scene_cache
show() { python - "$1" <<'PY'
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
ph = [v for k, v in d['e2e'].items() if k.startswith('phases_s')][0] if d.get('e2e') else []
print(sys.argv[1], 'value %.1f ms/step %.1f step_render_ms %s | e2e render %s | profiled pass %s | clocks %s' % (d['value'], d['ms_per_step'], d.get('step_render_ms'), [p[3] for p in ph], d['roofline'].get('profiled_pass'), d['clocks']))
PY
}
timeout 300 python bench.py --no-cpu --steps 6 > $out/g18_bench_1.json 2> $out/g18_bench_1.err; show $out/g18_bench_1.json
CUDAPATH_MAX_SPLIT=8 timeout 300 python bench.py --no-cpu --steps 6 > $out/g18_bench_split8.json 2>/dev/null; show $out/g18_bench_split8.json
CUDAPATH_MAX_SPLIT=24 timeout 300 python bench.py --no-cpu --steps 6 > $out/g18_bench_split24.json 2>/dev/null; show $out/g18_bench_split24.json

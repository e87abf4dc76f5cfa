#!/bin/bash
# round-2 GPU run 21 (1 GPU): leaf / open decision moved into k_refit, slimmer k_collapse, no final node copy: parity subset + build time + e2e
out=gpurun_out; mkdir -p $out
export CUDAPATH_SCENE_CACHE=/tmp/cudapath_scene_cache
timeout 900 python -m pytest tests -m gpu -q -x -k "collapse or job_size or closest_hit or full_size_ray_batch or mesh or teapot or any_hit or scene_bounds or kdtree or pixel_shards" > $out/g21_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/g21_pytest.log
tail -4 $out/g21_pytest.log
show() { python - "$1" <<'PY'
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
ph = [v for k, v in d['e2e'].items() if k.startswith('phases_s')][0] if d.get('e2e') else []
print(sys.argv[1], 'value %.1f ms/step %.1f step_render_ms %s | e2e %.1f render %s build %s | nodes/ray %.1f | build %s' % (d['value'], d['ms_per_step'], d.get('step_render_ms'), d['e2e']['value'], [p[3] for p in ph], [p[1] for p in ph], d['roofline']['nodes_per_ray'], d.get('build')))
PY
}
timeout 300 python bench.py --no-cpu > $out/g21_bench.json 2> $out/g21_bench.err; show $out/g21_bench.json
CUDAPATH_TRACE=1 timeout 300 python bench.py --no-cpu --no-e2e --steps 1 --warmup 1 2>&1 | grep -i "cudapath\]" | head -30 > $out/g21_trace.log; head -30 $out/g21_trace.log

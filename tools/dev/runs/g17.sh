#!/bin/bash
# round-2 GPU run 17 (1 GPU): is `value` (scene-resident loop) reproducible against the e2e device render? default bench twice, split 8, run-ahead without host waits
out=gpurun_out; mkdir -p $out
export CUDAPATH_SCENE_CACHE=/tmp/cudapath_scene_cache
show() { python - "$1" <<'PY'
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith('{')][-1])
ph = [v for k, v in d['e2e'].items() if k.startswith('phases_s')][0]
print(sys.argv[1], 'value %.1f ms/step %.1f | e2e %.1f render %s build %s | profiled pass %s | waits %.1f' % (d['value'], d['ms_per_step'], d['e2e']['value'], [p[3] for p in ph], [p[1] for p in ph], d['roofline'].get('profiled_pass'), d['host_waits_per_step']))
PY
}
for i in 1 2; do timeout 300 python bench.py --no-cpu > $out/g17_bench_$i.json 2> $out/g17_bench_$i.err; show $out/g17_bench_$i.json; done
CUDAPATH_MAX_SPLIT=8 timeout 300 python bench.py --no-cpu > $out/g17_bench_split8.json 2>/dev/null; show $out/g17_bench_split8.json
CUDAPATH_RUNAHEAD_MAX=0xffffffff timeout 300 python bench.py --no-cpu > $out/g17_bench_runahead.json 2>/dev/null; show $out/g17_bench_runahead.json
CUDAPATH_RUNAHEAD_MAX=0xffffffff timeout 300 python bench.py --no-cpu --spp 8 > $out/g17_bench_runahead8.json 2>/dev/null; show $out/g17_bench_runahead8.json
timeout 300 python bench.py --no-cpu --spp 8 > $out/g17_bench_8.json 2>/dev/null; show $out/g17_bench_8.json

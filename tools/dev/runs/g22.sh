#!/bin/bash
# round-2 GPU run 22 (1 GPU): where the steady-state build time goes (phase trace + ncu launch list of the build kernels)
out=gpurun_out; mkdir -p $out
export CUDAPATH_SCENE_CACHE=/tmp/cudapath_scene_cache
CUDAPATH_TRACE=1 timeout 300 python tools/dev/build_only.py 4 2>&1 | grep -E "build|cudapath\]" | tee $out/g22_build_trace.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/g22_build_launches.csv python tools/dev/build_only.py 2 > $out/g22_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv, re, collections
rows = list(csv.DictReader([l for l in open('gpurun_out/g22_build_launches.csv') if l.startswith('"')]))
# second build only: launches after the second k_init_shape_bounds
idx = [i for i, r in enumerate(rows) if 'k_init_shape_bounds' in r['Kernel Name']]
sel = rows[idx[-1]:]
agg = collections.OrderedDict()
for r in sel:
    n = re.sub(r'<.*', '', r['Kernel Name'])[:60]
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += float(r['Metric Value'].replace(',', '')) / 1e6
tot = sum(v[1] for v in agg.values())
for n, v in agg.items(): print('%-60s %3d %8.3f ms' % (n, v[0], v[1]))
print('total %.3f ms' % tot)
PY

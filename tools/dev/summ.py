import json, sys
for line in sys.stdin:
    line = line.strip()
    if not line.startswith('{'):
        continue
    d = json.loads(line); r = d.get('roofline') or {}
    print('%s Mpaths/s %.2f Mrays/s %.1f ms/step %.0f | nodes/ray %.1f prims/ray %.1f exact/ray %.2f frac %.3f share %s | build %s | e2e %s' % (
        ' '.join(sys.argv[1:]), d['value'], d.get('mrays_per_s', 0), d['ms_per_step'], r.get('nodes_per_ray', 0), r.get('pretests_per_ray', r.get('prims_per_ray', 0)), r.get('exact_tests_per_ray', 0), r.get('frac', 0),
        {k: round(v, 3) for k, v in (r.get('stage_share_of_step') or {}).items()}, d.get('build'), (d.get('e2e') or {}).get('value')))

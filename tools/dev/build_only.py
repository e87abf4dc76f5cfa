#!/usr/bin/env python3
"""Builds the bench scene N times on cuda:0 (for ncu launch lists of the BVH build and CUDAPATH_TRACE phase timings)."""
import os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, 'tests'))
import bench, cudapath
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
scene = sys.argv[2] if len(sys.argv) > 2 else 'hair-curl'
sc, shapes, env = bench.scene_arrays(scene, 1.0)
for k in range(n):
    t0 = time.perf_counter()
    ctx = bench.make_context(cudapath, sc, shapes, env, 0, sc['width'] * sc['height'] * sc['spp'])
    t1 = time.perf_counter()
    ctx.build()
    t2 = time.perf_counter()
    st = ctx.stats()
    print('build %d: create+upload %.1f ms, build %.1f ms, references %d nodes %d' % (k, (t1 - t0) * 1e3, (t2 - t1) * 1e3, st['bvh_references'], st['bvh_nodes']), flush=True)
    ctx.close()

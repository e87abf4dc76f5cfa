#!/usr/bin/env python3
"""Renders the bench scene N times with per-launch stage timing and CUDAPATH_TRACE=1: prints where the device sat idle between launches
(development: host-latency diagnosis on the shared GPU boxes)."""
import os, sys, time
os.environ['CUDAPATH_TRACE'] = '1'
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, 'tests'))
import numpy as np, torch
import bench, cudapath
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
sc, shapes, env = bench.scene_arrays('hair-curl', 1.0)
W, H, spp = sc['width'], sc['height'], sc['spp']
ctx = bench.make_context(cudapath, sc, shapes, env, 0, W * H * spp)
ctx.build()
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
film = torch.zeros((H, W, 5), dtype=torch.float32, device='cuda')
for k in range(2):
    film.zero_(); ctx.render_into(film.data_ptr(), spp, seed=k, stream=stream.cuda_stream)
for k in range(n):
    prof = (k % 2 == 0)
    ctx.set_options(profile_stages=prof)
    film.zero_(); t0 = time.perf_counter()
    ctx.render_into(film.data_ptr(), spp, seed=10 + k, stream=stream.cuda_stream)
    st = ctx.stats()
    print('render %d (%s): device %.1f ms, wall %.1f ms, stage sum %.1f ms' % (k, 'profiled' if prof else 'plain', st['render_ms'], (time.perf_counter() - t0) * 1e3,
          sum(st[s + '_ms'] for s in ('trace', 'shade', 'sort', 'raygen', 'splat')) if prof else float('nan')), flush=True)

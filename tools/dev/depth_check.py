import sys, os
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, cudapath as cp, orc
def rel_mse(a, b): return float(np.mean((a - b) ** 2 / (b ** 2 + 1e-2)))
for name in ('straight-hair', 'hair-on-head'):
    env = cp.bake_sunsky(**cp.scenes.sunsky_params(name))
    for md in (2, 3, 4, 5, 6, 7, 8, -1):
        ov = dict(width=32, height=24, spp=4, maxDepth=md)
        ctx = cp.scene_from_description(name, scale=0.004, overrides=ov); ctx.build()
        osc = orc.scene_from_description(name, scale=0.004, overrides=ov, envmap=env)
        g = ctx.render(4, seed=9); o = osc.render(4, seed=9)
        st = ctx.stats()
        print(name, 'maxDepth', md, 'relMSE %.2e' % rel_mse(cp.develop(g), cp.develop(o)), 'rays', st['rays'], osc.last_stats['rays'], 'shadow', st['shadow_rays'], osc.last_stats['shadow_rays'], flush=True)
        ctx.close()

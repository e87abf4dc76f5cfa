import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, cudapath as cp, orc
name = 'hair-on-head'
env = cp.bake_sunsky(**cp.scenes.sunsky_params(name))
ov = dict(width=32, height=24, spp=4, maxDepth=2)
ctx = cp.scene_from_description(name, scale=0.004, overrides=ov); ctx.set_film('box'); ctx.build()
osc = orc.scene_from_description(name, scale=0.004, overrides=ov, envmap=env); osc.set_film('box'); osc.build()
tot = 0
for s in range(4):
    g = ctx.render(4, seed=9, sample_begin=s, sample_end=s + 1); o = osc.render(4, seed=9, sample_begin=s, sample_end=s + 1)
    diff = np.abs(g[..., :3] - o[..., :3]).max(axis=2)
    ys, xs = np.nonzero(diff > 1e-3 * (np.abs(o[..., :3]).max(axis=2) + 1e-3))
    for y, x in zip(ys, xs):
        tot += 1
        if tot > 12: continue
        # reconstruct the camera ray of this sample through the oracle per-sample hook
        li, pos = osc.render_samples([[x, y]], [s], 4, seed=9)
        oo, dd, mn, mx = ctx.camera_rays(pos)
        gs, gp, gt, grec = ctx.intersect(oo, dd, mn, mx, record=True)
        os_, op, ot, orec = osc.intersect_full(oo, dd, mn, mx)
        print('pixel', x, y, 'sample', s, 'gpu', g[y, x, :3], 'oracle', o[y, x, :3], 'Li', li[0], '| hit gpu', gs[0], gp[0], gt[0], 'oracle', os_[0], op[0], ot[0])
        print('    rec gpu', np.round(grec[0], 5)); print('    rec orc', np.round(orec[0], 5))
print('differing samples', tot)

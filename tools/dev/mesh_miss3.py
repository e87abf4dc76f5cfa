import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, cudapath as cp, orc
name = 'hair-on-head'
env = cp.bake_sunsky(**cp.scenes.sunsky_params(name))
ov = dict(width=32, height=24, spp=4, maxDepth=2)
ctx = cp.scene_from_description(name, scale=0.004, overrides=ov); ctx.set_film('box'); ctx.build()
osc = orc.scene_from_description(name, scale=0.004, overrides=ov, envmap=env); osc.set_film('box'); osc.build()
s = 0
g = ctx.render(4, seed=9, sample_begin=s, sample_end=s + 1); o = osc.render(4, seed=9, sample_begin=s, sample_end=s + 1)
print('weights equal', np.array_equal(g[..., 4] > 0, o[..., 4] > 0), g[..., 4].sum(), o[..., 4].sum())
xy = np.array([[x, y] for y in range(24) for x in range(32)], np.uint32)
li, pos = osc.render_samples(xy, np.zeros(len(xy), np.uint32), 4, seed=9)
oo, dd, mn, mx = osc.camera_rays(pos)
go, gd, gmn, gmx = ctx.camera_rays(pos)
print('camera rays equal', np.array_equal(oo, go), np.abs(dd - gd).max())
os_, op, ot = osc.intersect(oo, dd, mn, mx, mode=0)
gs, gp, gt = ctx.intersect(go, gd, gmn, gmx)
print('primary hits equal', np.array_equal(os_, gs), np.array_equal(op, gp))
# where does each sample land in the film (box filter: floor(pos))
px = np.floor(pos[:, 0]).astype(int); py = np.floor(pos[:, 1]).astype(int)
print('samples landing outside own pixel', ((px != xy[:, 0]) | (py != xy[:, 1])).sum())
gv = g[xy[:, 1], xy[:, 0], :3]; ov_ = o[xy[:, 1], xy[:, 0], :3]
bad = np.nonzero(np.abs(gv - ov_).max(axis=1) > 1e-3 * (np.abs(ov_).max(axis=1) + 1e-3))[0]
print('differing pixels', len(bad))
for i in bad[:16]:
    print(xy[i], 'gpu', gv[i], 'oracle', ov_[i], 'Li', li[i], 'hit shape/prim', os_[i], op[i], 't', ot[i])

import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, cudapath as cp, orc
name = 'hair-on-head'
env = cp.bake_sunsky(**cp.scenes.sunsky_params(name))
ov = dict(width=32, height=24, spp=4, maxDepth=2)
ctx = cp.scene_from_description(name, scale=0.004, overrides=ov); ctx.build()
osc = orc.scene_from_description(name, scale=0.004, overrides=ov, envmap=env)
rng = np.random.default_rng(1)
pxy = (rng.random((200000, 2)) * np.array([32, 24])).astype(np.float32)
o, d, mn, mx = ctx.camera_rays(pxy)
gs, gp, gt = ctx.intersect(o, d, mn, mx)
os_, op, ot = osc.intersect(o, d, mn, mx, mode=0)
ob, opb, otb = osc.intersect(o, d, mn, mx, mode=2)
print('oracle bvh vs brute mismatch', ((os_ != ob) | (op != opb)).sum())
bad = np.nonzero((gs != ob) | (gp != opb))[0]
print('gpu vs brute mismatches', len(bad), 'of', len(o))
for i in bad[:12]:
    print(i, 'gpu', gs[i], gp[i], gt[i], 'oracle', ob[i], opb[i], otb[i], 'mint/maxt', mn[i], mx[i], 'd', d[i])
print('scene bounds', ctx.scene_bounds()[0], osc.scene_bounds()[0])

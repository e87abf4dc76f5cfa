# diagnostic: full-size furball, rays whose hit primitive agrees between the CUDA path and the oracle but whose fp32 distance does not
import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
import cudapath as cp, orc
def sd(rng, n):
    v = rng.normal(size=(n, 3)); v /= np.linalg.norm(v, axis=1, keepdims=True); return v.astype(np.float32)
ctx = cp.scene_from_description('furball', scale=1.0); ctx.build()
env = cp.bake_sunsky(**cp.scenes.sunsky_params('furball'))
osc = orc.scene_from_description('furball', scale=1.0, envmap=env)
rng = np.random.default_rng(11)
aabb, bs = osc.scene_bounds()
n = 1000000
p1 = bs[:3] + bs[3] / 1.5 * 0.8 * sd(rng, n); p2 = bs[:3] + bs[3] / 1.5 * 0.8 * sd(rng, n)
d = p2 - p1; d /= np.linalg.norm(d, axis=1, keepdims=True); o = p1.astype(np.float32); d = d.astype(np.float32)
gs, gp, gt = ctx.intersect(o, d, 0.0, np.inf)
os_, op, ot = osc.intersect(o, d, 0.0, np.inf, mode=0)
hit = (gs >= 0) & (os_ >= 0) & (gs == os_) & (gp == op)
diff = hit & (gt != ot)
idx = np.nonzero(diff)[0]
print('hits', hit.sum(), 'differing t', len(idx))
ulp = np.abs(gt[idx].view(np.int32).astype(np.int64) - ot[idx].view(np.int32).astype(np.int64))
print('ulp diffs: max', ulp.max() if len(idx) else 0, 'hist', np.bincount(np.minimum(ulp, 10)))
np.set_printoptions(precision=9, floatmode='unique')
# the same rays alone
g2 = ctx.intersect(o[idx], d[idx], 0.0, np.inf)
print('alone: same as batch', np.array_equal(g2[2], gt[idx]), 'equal to oracle', (g2[2] == ot[idx]).sum())
ob = osc.intersect(o[idx[:50]], d[idx[:50]], 0.0, np.inf, mode=2)
print('oracle brute force == oracle bvh t', (ob[2] == ot[idx[:50]]).sum(), 'of', min(50, len(idx)), '; == gpu', (ob[2] == gt[idx[:50]]).sum())
for i in idx[:8]:
    print(i, 'o', o[i], 'd', d[i], 'shape', gs[i], 'prim', gp[i], 'gpu t', gt[i], 'oracle t', ot[i])
np.savez('/root/repo/gpurun_out/furball_t_diff.npz', o=o[idx], d=d[idx], gs=gs[idx], gp=gp[idx], gt=gt[idx], ot=ot[idx])

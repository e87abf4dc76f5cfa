# diagnostic: intersection records (fillIntersectionRecord) of the CUDA path vs the oracle at the bit level
import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
import cudapath as cp, orc
def sd(rng, n):
    v = rng.normal(size=(n, 3)); v /= np.linalg.norm(v, axis=1, keepdims=True); return v.astype(np.float32)
for name, scale in (('hair-curl', 0.02), ('furball', 1.0), ('hair-on-head', 1.0)):
    ctx = cp.scene_from_description(name, scale=scale); ctx.build()
    env = cp.bake_sunsky(**cp.scenes.sunsky_params(name))
    osc = orc.scene_from_description(name, scale=scale, envmap=env)
    rng = np.random.default_rng(7)
    aabb, bs = osc.scene_bounds()
    n = 400000
    p1 = bs[:3] + bs[3] / 1.5 * 0.8 * sd(rng, n); p2 = bs[:3] + bs[3] / 1.5 * 0.8 * sd(rng, n)
    d = p2 - p1; d /= np.linalg.norm(d, axis=1, keepdims=True); o = p1.astype(np.float32); d = d.astype(np.float32)
    gs, gp, gt, grec = ctx.intersect(o, d, 0.0, np.inf, record=True)
    os_, op, ot, orec = osc.intersect_full(o, d, 0.0, np.inf)
    m = (gs >= 0) & (gs == os_) & (gp == op)
    ne = grec[m] != orec[m]
    print(name, 'hits', int(m.sum()), 'differing record floats per column', ne.sum(axis=0).tolist(), 'max abs diff', float(np.abs(grec[m] - orec[m]).max()))
    ctx.close()

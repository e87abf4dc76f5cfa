# diagnostic: how far are the envmap tables / lookups / camera rays of the CUDA path from the oracle (bit level)?
import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
import cudapath as cp, orc
name = 'hair-curl'
ctx = cp.scene_from_description(name, scale=0.02); ctx.build()
env = cp.bake_sunsky(**cp.scenes.sunsky_params(name))
osc = orc.scene_from_description(name, scale=0.02, envmap=env)
gr, gc, gw, gn = ctx.env_tables(512, 256); or_, oc, ow, on = osc.env_tables()
def ne(a, b):
    a = np.asarray(a); b = np.asarray(b)
    m = ~((a == b) | (np.isnan(a) & np.isnan(b)))
    return int(m.sum()), a.size, float(np.nanmax(np.abs(a - b))) if a.size else 0.0
print('cdfRows', ne(gr, or_), 'cdfCols', ne(gc, oc), 'rowWeights', ne(gw, ow), 'norm', gn, on)
rng = np.random.default_rng(8)
n = 1 << 22
d = rng.normal(size=(n, 3)).astype(np.float32); d /= np.linalg.norm(d, axis=1, keepdims=True)
grgb, gpdf = ctx.env_eval(d); orgb, opdf = osc.env_eval(d)
print('env_eval rgb', ne(grgb, orgb), 'pdf', ne(gpdf, opdf))
ref = (rng.normal(size=(n, 3)) * 2 + np.array([0, 6, 0])).astype(np.float32)
smp = rng.random((n, 2), dtype=np.float32)
gd, gv, gp, gdist = ctx.env_sample(ref, smp); od, ov, op, odist = osc.env_sample(ref, smp)
print('env_sample dir', ne(gd, od), 'value', ne(gv, ov), 'pdf', ne(gp, op), 'dist', ne(gdist, odist))
W = cp.scenes.SCENES[name]['width']
pxy = (rng.random((n, 2)) * W).astype(np.float32)
ctx2 = cp.scene_from_description(name, scale=0.02); ctx2.build()
osc2 = orc.scene_from_description(name, scale=0.02, envmap=env)
go, gd_, gmin, gmax = ctx2.camera_rays(pxy); oo, od_, omin, omax = osc2.camera_rays(pxy)
print('camera o', ne(go, oo), 'd', ne(gd_, od_), 'mint', ne(gmin, omin), 'maxt', ne(gmax, omax))
pos = (rng.random((1 << 20, 2)) * 96).astype(np.float32)

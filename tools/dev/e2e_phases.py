import sys, time, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import cudapath, torch
from bench import scene_arrays, pin
name = sys.argv[1] if len(sys.argv) > 1 else 'hair-curl'
sc, shapes, env = scene_arrays(name, 1.0)
pshapes = []
keep = []
for xyz, st, r, b in shapes:
    a, ta = pin(xyz); s2, ts = pin(st); keep += [ta, ts]; pshapes.append((a, s2, r, b))
for rep in range(6):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ctx = cudapath.Context(0); t1 = time.perf_counter()
    ids = []
    for xyz, st, r, b in pshapes:
        b = dict(b); t = b.pop('type'); b.pop('id', None)
        ids.append(ctx.add_bsdf(t, **b))
    t2 = time.perf_counter()
    for (xyz, st, r, b), i in zip(pshapes, ids):
        ctx.add_hair(xyz, st, r, i)
    t3 = time.perf_counter()
    ctx.set_envmap(env); ctx.set_camera(np.array(sc['camera'], np.float32).reshape(4, 4), sc['fov'], width=sc['width'], height=sc['height'])
    ctx.set_film('tent'); ctx.set_integrator(maxDepth=sc['maxDepth'], rrDepth=5, strictNormals=True)
    t4 = time.perf_counter()
    ctx.build(); t5 = time.perf_counter()
    film = ctx.render(sc['spp'], seed=rep); t6 = time.perf_counter()
    st = ctx.stats()
    ctx.close(); t7 = time.perf_counter()
    print('rep %d: create %.3f bsdf %.3f add_hair %.3f set %.3f build %.3f (dev %.3f) render %.3f (dev %.3f) close %.3f total %.3f' % (
        rep, t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, st['build_ms'] / 1e3, t6 - t5, st['render_ms'] / 1e3, t7 - t6, t6 - t0))

#!/usr/bin/env python3
"""Generates the golden vectors under tests/golden/ from the CPU oracle (run from the repo root: python tests/golden/make_golden.py).

The reference holds no golden vector for the hair path and cannot be run here, so these fixtures are
outputs of the oracle restatement on fixed seeded inputs; they pin the oracle against accidental drift and give the GPU tests a
second, file-based target.  Needs refdata/ (data/microfacet/*.dat) and, for the render fixture, oracle/_ref.
"""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, 'tests'))
import orc
import cudapath

HAIR_RGB = (0.143016, 0.0156076, 1.80928e-005)


def sphere_dirs(rng, n):
    v = rng.normal(size=(n, 3)); v /= np.linalg.norm(v, axis=1, keepdims=True)
    return v.astype(np.float32)


def second_set():
    """Fixtures added with the mesh / roughplastic / fixed-Marschner rows (the first set is left untouched so that it keeps pinning
    the state it was generated from)."""
    rng = np.random.default_rng(0x5eed2)
    n = 4096
    s = orc.Scene()
    mats = [('roughplastic', dict(intIOR=1.55, extIOR=1.0, alpha=0.2, distribution='ggx', diffuseReflectance=HAIR_RGB)),
            ('roughplastic', dict(intIOR=1.49, extIOR=1.000277, alpha=0.1, distribution='beckmann', nonlinear=True, diffuseReflectance=(0.6, 0.5, 0.4))),
            ('roughplastic', dict(intIOR=1.55, extIOR=1.0, alpha=0.3, distribution='phong', diffuseReflectance=(0.2, 0.3, 0.4))),
            ('marschner_fixed', dict(intIOR=1.55, extIOR=1.000277)),
            ('diffuse', dict(reflectance=(0.5, 0.4, 0.3))), ('twosided', dict(reflectance=(0.5, 0.4, 0.3)))]
    for t, p in mats:
        s.add_bsdf(t, **p)
    out = dict(wi=sphere_dirs(rng, n), wo=sphere_dirs(rng, n), sample=rng.random((n, 2)).astype(np.float32), extra=rng.random((n, 4)).astype(np.float32))
    for b in range(len(mats)):
        out['eval_%d' % b], out['pdf_%d' % b] = s.bsdf_eval(b, out['wi'], out['wo'])
        wo, wt, pdf, ty = s.bsdf_sample(b, out['wi'], out['sample'], out['extra'])
        out['swo_%d' % b], out['swt_%d' % b], out['spdf_%d' % b], out['sty_%d' % b] = wo, wt, pdf, ty
    np.savez_compressed(os.path.join(HERE, 'bsdf2_golden.npz'), **out)

    # fibers + triangle meshes in one scene: chords through the head, brute-force answers and full intersection records
    ov = dict(width=32, height=24, spp=4, maxDepth=6)
    env = cudapath.bake_sunsky(**cudapath.scenes.sunsky_params('hair-on-head'))
    g = orc.scene_from_description('hair-on-head', scale=0.004, overrides=ov, envmap=env)
    m = 20000
    c = np.array([0.15, 11.0, 0.5]); r = 9.0
    p1 = c + r * sphere_dirs(rng, m); p2 = c + r * sphere_dirs(rng, m)
    d = p2 - p1; d /= np.linalg.norm(d, axis=1, keepdims=True)
    o = p1.astype(np.float32); d = d.astype(np.float32)
    sh, pr, t = g.intersect(o, d, 0.0, np.inf, mode=2)
    sh2, pr2, t2, rec = g.intersect_full(o, d, 0.0, np.inf)
    np.savez_compressed(os.path.join(HERE, 'mesh_golden.npz'), o=o, d=d, shape=sh, prim=pr, t=t, rec=rec, rec_shape=sh2, rec_prim=pr2)
    film = g.render(4, seed=9, threads=2)
    np.savez_compressed(os.path.join(HERE, 'render_mesh_golden.npz'), film=film)


THIRD_SET_MATS = [('thindielectric', dict(intIOR=1.55, extIOR=1.0, specularReflectance=HAIR_RGB, specularTransmittance=HAIR_RGB)),
                  ('thindielectric', dict(intIOR=1.5046, extIOR=1.000277, specularReflectance=(0.9, 0.5, 0.1), specularTransmittance=(2.0, 1.0, 0.5))),
                  ('marschnerdielectric', dict(intIOR=1.55, extIOR=1.0, exponent=5.0, specularTransmittance=HAIR_RGB, specularReflectance=HAIR_RGB, diffuseReflectance=HAIR_RGB)),
                  ('marschnerdielectric', dict(intIOR=1.501, extIOR=1.000277, diffuseReflectance=(0.3, 0.2, 0.1), specularReflectance=(0.4, 0.3, 0.2),
                                               specularTransmittance=(1.5, 0.6, 0.7)))]


def third_set():
    """Fixtures added with the `thindielectric` / `marschnerdielectric` rows: BSDF tuples in both measures and
    the two straight-hair dielectric scenes rendered small (ENull vertices, hits from inside the fibers)."""
    rng = np.random.default_rng(0x5eed3)
    n = 4096
    s = orc.Scene()
    for t, p in THIRD_SET_MATS:
        s.add_bsdf(t, **p)
    out = dict(wi=sphere_dirs(rng, n), sample=rng.random((n, 2)).astype(np.float32))
    for b in range(len(THIRD_SET_MATS)):
        wo, wt, pdf, ty = s.bsdf_sample(b, out['wi'], out['sample'])
        out['swo_%d' % b], out['swt_%d' % b], out['spdf_%d' % b], out['sty_%d' % b] = wo, wt, pdf, ty
        for discrete in (0, 1):        # evaluated at the sampled directions: exactly on the delta directions for the discrete components
            out['eval_%d_%d' % (b, discrete)], out['pdf_%d_%d' % (b, discrete)] = s.bsdf_eval(b, out['wi'], wo, discrete=bool(discrete))
    np.savez_compressed(os.path.join(HERE, 'bsdf3_golden.npz'), **out)
    films = {}
    for name in ('straight-hair-thindielectric', 'straight-hair-dielectric'):
        ov = dict(width=32, height=24, spp=4, maxDepth=24)
        env = cudapath.bake_sunsky(**cudapath.scenes.sunsky_params(name))
        films[name.replace('-', '_')] = orc.scene_from_description(name, scale=0.004, overrides=ov, envmap=env).render(4, seed=13, threads=2)
    np.savez_compressed(os.path.join(HERE, 'render_dielectric_golden.npz'), **films)


def main():
    if '--third' in sys.argv:
        third_set()
        print('third set written')
        return
    if '--second' in sys.argv:
        second_set()
        print('second set written')
        return
    rng = np.random.default_rng(0x5eed)
    n = 4096
    s = orc.Scene()
    s.add_bsdf('kajiyakay', diffuseReflectance=HAIR_RGB, exponent=10.0)
    s.add_bsdf('marschner', intIOR=1.55, extIOR=1.0, specularReflectance=(0.592384, 0.32628, 0.0528657))
    s.add_bsdf('marschner', intIOR=1.55, extIOR=1.0, alpha=0.2, distribution='ggx', diffuseReflectance=HAIR_RGB)
    out = dict(wi=sphere_dirs(rng, n), wo=sphere_dirs(rng, n), sample=rng.random((n, 2)).astype(np.float32))
    for b in range(3):
        out['eval_%d' % b], out['pdf_%d' % b] = s.bsdf_eval(b, out['wi'], out['wo'])
        wo, wt, pdf, ty = s.bsdf_sample(b, out['wi'], out['sample'])
        out['swo_%d' % b], out['swt_%d' % b], out['spdf_%d' % b], out['sty_%d' % b] = wo, wt, pdf, ty
    np.savez_compressed(os.path.join(HERE, 'bsdf_golden.npz'), **out)

    # intersection: furball at 0.2 % of the strands, kdbench-style chords + surface rays
    ov = dict(width=32, height=24, spp=4, maxDepth=6)
    env = cudapath.bake_sunsky(**cudapath.scenes.sunsky_params('furball'))
    g = orc.scene_from_description('furball', scale=0.01, overrides=ov, envmap=env)
    aabb, bs = g.scene_bounds()
    m = 20000
    c = 0.5 * (aabb[:3] + aabb[3:]); r = 0.5 * np.linalg.norm(aabb[3:] - aabb[:3])       # chords of the fiber ball's own bounding sphere
    p1 = c + r * sphere_dirs(rng, m); p2 = c + r * sphere_dirs(rng, m)
    d = p2 - p1; d /= np.linalg.norm(d, axis=1, keepdims=True)
    o = p1.astype(np.float32); d = d.astype(np.float32)
    sh, pr, t = g.intersect(o, d, 0.0, np.inf, mode=2)          # brute force = the reference semantics with a trivial visiting order
    np.savez_compressed(os.path.join(HERE, 'intersect_golden.npz'), o=o, d=d, shape=sh, prim=pr, t=t)

    c = orc.scene_from_description('curly-hair', scale=0.004, overrides=ov, envmap=cudapath.bake_sunsky(**cudapath.scenes.sunsky_params('curly-hair')))
    film = c.render(4, seed=5, threads=2)
    np.savez_compressed(os.path.join(HERE, 'render_golden.npz'), film=film)
    print('golden vectors written:', [f for f in os.listdir(HERE) if f.endswith('.npz')])


if __name__ == '__main__':
    main()

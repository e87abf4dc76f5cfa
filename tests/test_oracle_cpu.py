"""CPU tests (run with -m "not gpu"): the oracle against the compiled reference pieces and closed forms, the product's
host-side code (file loaders, sunsky bake, develop) against the oracle, and the committed golden vectors."""
import ctypes
import os
import struct
import numpy as np
import pytest

HAIR_RGB = (0.143016, 0.0156076, 1.80928e-005)
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
REF_GEOM = os.path.join(os.path.dirname(GOLDEN), '..', 'oracle', '_ref', 'libref_geom.so')     # reference text executed as written (oracle/ref_shim/ref_geom.cpp, ref_loader.cpp)


def sphere_dirs(rng, n):
    v = rng.normal(size=(n, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    return v.astype(np.float32)


# ------------------------------------------------------------------------------------------------ pinning against oracle/_ref
needs_ref = pytest.mark.skipif(not os.path.exists(os.path.join(os.path.dirname(GOLDEN), '..', 'oracle', '_ref', 'libref_pieces.so')),
                               reason='oracle/_ref not built (needs /root/reference)')


@needs_ref
def test_gauss_legendre_bitwise_vs_reference(oracle):
    L, R = oracle.lib(), oracle.ref_lib()
    a = [np.zeros(140, np.float32) for _ in range(4)]
    L.orc_gauss_legendre_140(oracle.p(a[0]), oracle.p(a[1])); R.ref_gauss_legendre_140(oracle.p(a[2]), oracle.p(a[3]))
    assert np.array_equal(a[0], a[2]) and np.array_equal(a[1], a[3])
    assert abs(a[1].sum() - 2.0) < 1e-5 and a[0][0] > 0.9998 and np.all(np.diff(a[0]) < 0)     # weights sum to 2, nodes descend from +1


@needs_ref
def test_interpolated_distribution_bitwise_vs_reference(oracle):
    L, R = oracle.lib(), oracle.ref_lib()
    rng = np.random.default_rng(3)
    size, num, n = 64, 64, 20000
    w = rng.random(size * num).astype(np.float32) ** 4
    w[5 * size:6 * size] = 0                                   # near-degenerate row -> uniform fallback (hpp:49-56)
    dist = (rng.random(n) * 70 - 3).astype(np.float32); u = rng.random(n).astype(np.float32)
    u[:3] = (0.0, 1.0, 0.5)
    ou, ox = np.zeros(n, np.float32), np.zeros(n, np.int32)
    ru, rx, rp, rs = np.zeros(n, np.float32), np.zeros(n, np.int32), np.zeros(n, np.float32), np.zeros(n, np.float32)
    L.orc_interp_dist_warp(oracle.p(w), size, num, n, oracle.p(dist), oracle.p(u), oracle.p(ou), oracle.p(ox))
    R.ref_interp_dist(oracle.p(w), size, num, n, oracle.p(dist), oracle.p(u), oracle.p(ru), oracle.p(rx), oracle.p(rp), oracle.p(rs))
    assert np.array_equal(ox, rx) and np.array_equal(ou, ru)


@needs_ref
def test_sunsky_bake_product_vs_oracle(cp, oracle):
    """Product bake (own Hosek-Wilkie evaluation reading refdata) vs oracle bake (the reference's skymodel.cpp compiled as is)."""
    for name in ('straight-hair', 'hair-curl'):
        sp = cp.scenes.sunsky_params(name)
        a = cp.bake_sunsky(**sp)
        b = oracle.bake_sunsky(sp['turbidity'], 0.2, sp['sunDirection'], sp['skyScale'], sp['sunScale'], sp['sunRadiusScale'], 512)
        assert a.shape == (256, 512, 3) and (b[128:] == 0).all()             # below the horizon: black
        sky = b < 20                                                         # texels without sun-disc samples
        assert np.abs(a[sky] - b[sky]).max() <= 2e-5 * b[sky].max()
        # sun texels: a QMC sample on a texel border may land in the neighbouring texel -> compare the total energy instead
        assert abs(a.sum() - b.sum()) <= 1e-5 * b.sum()
        assert (np.abs(a - b) / np.maximum(b, 1e-3) > 1e-3).sum() < 200


def test_sunsky_sun_disc_energy(oracle):
    """Closed form: the sun-disc splat adds radiance * solidAngle * W*H / (2 pi^2) / sin(theta) in total (sunsky.cpp:195-214)."""
    if not oracle.have_ref():
        pytest.skip('oracle/_ref not built')
    d = (0.3, 0.8, 0.2)
    with_sun = oracle.bake_sunsky(3.0, 0.2, d, 1.0, 1.0, 10.0, 512)
    sky_only = oracle.bake_sunsky(3.0, 0.2, d, 1.0, 0.0, 10.0, 512)
    disc = (with_sun - sky_only)
    assert disc.min() >= 0 and disc.sum() > 0
    ys, xs = np.nonzero(disc[..., 1] > 0)
    # the splat lies around the sun direction
    dn = np.array(d) / np.linalg.norm(d)
    az = np.arctan2(dn[0], -dn[2]) % (2 * np.pi); el = np.arccos(dn[1])
    assert abs(xs.mean() - az * 512 / (2 * np.pi)) < 8 and abs(ys.mean() - el * 256 / np.pi) < 8


# ------------------------------------------------------------------------------------------------ closed forms / invariants of the oracle
def test_philox_known_answer(oracle):
    """Random123 known-answer vectors for Philox4x32-10 (Salmon et al. 2011, kat_vectors)."""
    out = np.zeros(4, np.uint32)
    oracle.lib().orc_philox(0, 0, 0, 0, 0, 0, oracle.p(out))
    assert [hex(x) for x in out] == ['0x6627e8d5', '0xe169c58d', '0xbc57ac4c', '0x9b00dbd8']
    oracle.lib().orc_philox(0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, oracle.p(out))
    assert [hex(x) for x in out] == ['0x408f276d', '0x41c83b0e', '0xa20bc7c6', '0x6d5451fd']
    oracle.lib().orc_philox(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0, oracle.p(out))
    assert [hex(x) for x in out] == ['0xd16cfe09', '0x94fdcceb', '0x5001e420', '0x24126ea1']


def test_half_quantisation_matches_numpy(oracle):
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.normal(size=5000) * 100, rng.random(5000) * 1e-6, [0, 65504, 65520, 1e9, 6e-8, 5.96e-8, 2.98e-8, 1.0009766]]).astype(np.float32)
    h = np.zeros(len(x), np.uint16); f = np.zeros(len(x), np.float32)
    oracle.lib().orc_half_roundtrip(oracle.p(x), len(x), oracle.p(h), oracle.p(f))
    with np.errstate(over='ignore'):
        ref = x.astype(np.float16)
    assert np.array_equal(h, ref.view(np.uint16)) and np.array_equal(f, ref.astype(np.float32))


def make_bsdf_scene(oracle):
    s = oracle.Scene()
    s.add_bsdf('kajiyakay', diffuseReflectance=HAIR_RGB, exponent=10.0)
    s.add_bsdf('marschner', intIOR=1.55, extIOR=1.0, specularReflectance=(0.592384, 0.32628, 0.0528657))
    s.add_bsdf('marschner', intIOR=1.55, extIOR=1.0, alpha=0.2, distribution='ggx', diffuseReflectance=HAIR_RGB)
    return s


def test_kajiyakay_closed_forms(oracle):
    s = make_bsdf_scene(oracle)
    wi = np.array([[0.6, 0.0, 0.8]] * 3, np.float32)
    wo = np.array([[-0.6, 0.0, 0.8], [0.6, 0.0, 0.8], [0.0, 0.0, -1.0]], np.float32)
    ev, pdf = s.bsdf_eval(0, wi, wo)
    spec = 0.2                                   # default specularReflectance (kajiyakay.cpp:64-65)
    e = 10.0
    # forward scatter: tl=te=0.6, alpha=1 -> 0.15*spec*(e+2)/(4 pi) + diffuse/pi, times cos(theta_o)
    expect = (0.15 * spec * (e + 2) / (4 * np.pi) + np.array(HAIR_RGB) / np.pi) * 0.8
    assert np.allclose(ev[0], expect, rtol=1e-5)
    assert np.allclose(ev[1], np.array(HAIR_RGB) / np.pi * 0.8, rtol=1e-5)     # wi.x*wo.x > 0: no specular lobe (kajiyakay.cpp:157)
    assert (ev[2] == 0).all() and pdf[2] == 0                                  # below the horizon
    # pdf integrates to ~1 over the hemisphere (Phong lobe about the mirror direction + cosine lobe)
    rng = np.random.default_rng(1)
    d = sphere_dirs(rng, 400000); d[:, 2] = np.abs(d[:, 2])
    _, p = s.bsdf_eval(0, np.repeat(wi[:1], len(d), 0), d)
    assert abs(p.mean() * 2 * np.pi - 1.0) < 0.02


def test_marschner_quirks(oracle):
    s = make_bsdf_scene(oracle)
    rng = np.random.default_rng(2)
    n = 2000
    wi, wo = sphere_dirs(rng, n), sphere_dirs(rng, n)
    ev, pdf = s.bsdf_eval(2, wi, wo)
    assert (pdf == 1).all()                                                    # quirk 1
    assert np.isfinite(ev).all() and (ev >= 0).all()
    # quirk 2: the specular term depends on wi only through wi.y; the diffuse term vanishes for wo.z < 0 (T(c<0) = 0)
    wo_b = wo.copy(); wo_b[:, 2] = -np.abs(wo_b[:, 2]) - 1e-3; wo_b /= np.linalg.norm(wo_b, axis=1, keepdims=True)
    wi2 = wi.copy()
    ang = rng.random(n) * 2 * np.pi
    r = np.sqrt(np.maximum(1 - wi[:, 1] ** 2, 0))
    wi2[:, 0] = r * np.cos(ang); wi2[:, 2] = r * np.sin(ang)
    e1, _ = s.bsdf_eval(2, wi, wo_b); e2, _ = s.bsdf_eval(2, wi2.astype(np.float32), wo_b)
    assert np.allclose(e1, e2, rtol=2e-4, atol=1e-7 * e1.max())
    # sample(): weight == eval(wi, wo_sampled) (pdf 1), spec branch flagged EDeltaReflection with components 5/6/7
    smp = rng.random((n, 2)).astype(np.float32)
    swo, swt, spdf, sty = s.bsdf_sample(2, wi, smp)
    ev2, _ = s.bsdf_eval(2, wi, swo)
    assert np.array_equal(ev2, swt) and (spdf == 1).all()
    types, comps = sty & 0xff, sty >> 8
    assert set(np.unique(types)) <= {0x20, 0x2} and set(np.unique(comps[types == 0x20])) <= {5, 6, 7} and (comps[types == 0x2] == 1).all()


def test_marschner_table_symmetries(oracle):
    s = make_bsdf_scene(oracle)
    t = s.marschner_tables(1)
    tab = t['tables']
    assert np.isfinite(tab[:, 1:]).all()                 # row 0 is cos(theta_d) = 0: iorPrime = inf/0 -> NaN handling as in the reference
    assert np.allclose(tab[0, 1:, :, 0], tab[0, 1:, :, 1]) and np.allclose(tab[0, 1:, :, 0], tab[0, 1:, :, 2])     # R lobe is grey
    assert (np.abs(t['cdfs'][:, :, 0]) == 0).all() and (t['cdfs'][:, :, -1] == 1).all()
    assert (np.diff(t['cdfs'], axis=2) >= -1e-7).all()
    assert 0 < t['Fdr'] < 1 and len(t['rt']) == 100 and (np.diff(t['rt']) >= -1e-4).all()
    sw = t['specW']
    lum = lambda c: 0.212671 * c[0] + 0.715160 * c[1] + 0.072169 * c[2]
    assert abs(sw - lum((0.592384, 0.32628, 0.0528657)) / (lum((0.592384, 0.32628, 0.0528657)) + 0.5)) < 1e-6


def tiny_hair():
    # three fibers: a straight one along x, a bent one, and a 2-vertex stub
    xyz = np.array([[-1, 0, 0], [0, 0, 0], [1, 0, 0], [2, 0.5, 0],
                    [0, 1, -1], [0, 1, 0], [0.3, 1, 1],
                    [3, 3, 3], [3, 4, 3]], np.float32)
    st = np.array([1, 0, 0, 0, 1, 0, 0, 1, 0], np.uint8)
    return xyz, st


def test_cylinder_intersection_analytic(oracle):
    s = oracle.Scene()
    b = s.add_bsdf('kajiyakay')
    xyz, st = tiny_hair()
    s.add_hair(xyz, st, 0.1, b)
    s.set_camera(np.eye(4, dtype=np.float32), 35.0, width=8, height=8)
    s.build()
    o = np.array([[-0.5, 0, 5], [-0.5, 0, 0], [-0.5, 5, 0], [10, 10, 10], [3, 3.5, 8]], np.float32)
    d = np.array([[0, 0, -1], [0, 0, -1], [0, -1, 0], [0, 0, 1], [0, 0, -1]], np.float32)
    for mode in (0, 2):
        sh, pr, t = s.intersect(o, d, 0.0, np.inf, mode=mode)
        assert list(sh) == [0, 0, 0, -1, 0] and list(pr[[0, 1, 2, 4]]) == [0, 0, 0, 7]
        assert abs(t[0] - 4.9) < 1e-6 and abs(t[1] - 0.1) < 1e-6 and abs(t[2] - 4.9) < 1e-6   # t[1]: a ray from inside exits through the far wall
        # The stub fiber is the scene's z-extreme and segment bounds use radius*(1-Epsilon) (hair.cpp:375), but both kd-trees enlarge
        # their box by MTS_KD_AABB_EPSILON = 1e-3 after the build (gkdtree.h:1213-1220), so the entry distance of the shape / scene
        # box (< 4.9) is in front of the near root and the near wall is reported.
        assert abs(t[4] - 4.9) < 1e-6
    # any-hit respects [mint, maxt]
    assert list(s.intersect(o[:1], d[:1], 0.0, 4.0, mode=1)[0]) == [-1]
    assert list(s.intersect(o[:1], d[:1], 0.0, 4.95, mode=1)[0]) == [0]
    # the record: p lies on the cylinder, n is radial, s is the fiber tangent, wi = toLocal(-d)
    sh, pr, t, rec = s.intersect_full(o[:1], d[:1], 0.0, np.inf)
    p, n, sx, tx, wi = rec[0, 0:3], rec[0, 3:6], rec[0, 6:9], rec[0, 9:12], rec[0, 12:15]
    assert np.allclose(p, (-0.5, 0, 0.1), atol=1e-6) and np.allclose(n, (0, 0, 1), atol=1e-6) and np.allclose(np.abs(sx), (1, 0, 0), atol=1e-6)
    assert np.allclose(wi, (0, 0, 1), atol=1e-6) and abs(np.dot(n, sx)) < 1e-6 and np.allclose(np.cross(n, sx), tx, atol=1e-6)


def test_kdtree_bounds_are_enlarged(oracle):
    """gkdtree.h:1213-1220: getAABB() of a built kd-tree = tight bounds enlarged by 1e-3 (relative to the extent, plus absolute); the
    hair shape's box is enlarged once, the scene box (union of the shape boxes) once more."""
    s = oracle.Scene()
    b = s.add_bsdf('kajiyakay')
    xyz = np.array([[0, 0, 0], [4, 0, 0], [8, 1, 0]], np.float32)
    s.add_hair(xyz, np.array([1, 0, 0], np.uint8), 0.5, b)
    s.set_camera(np.eye(4, dtype=np.float32), 35.0, width=8, height=8)
    s.build()
    seg = s.segment_bounds(0, 2)
    tight = np.concatenate([seg[:, :3].min(axis=0), seg[:, 3:].max(axis=0)]).astype(np.float32)
    def enlarge(a):
        a = a.astype(np.float32).copy(); eps = np.float32(1e-3)
        a[:3] = a[:3] - ((a[3:] - a[:3]) * eps + eps)
        a[3:] = a[3:] + ((a[3:] - a[:3]) * eps + eps)
        return a
    aabb, _ = s.scene_bounds()
    assert np.array_equal(aabb, enlarge(enlarge(tight)))
    assert (aabb[:3] < tight[:3] - 1e-3).all() and (aabb[3:] > tight[3:] + 1e-3).all()
    # a ray grazing just outside the tight box but inside the enlarged one is still clipped to a valid interval (and misses the fiber)
    assert s.intersect([[4.0, tight[4] + 5e-4, 5.0]], [[0, 0, -1]], 0.0, np.inf, mode=2)[0][0] == -1


def test_miter_joint_has_no_gap_or_overlap(oracle):
    """At a joint between two segments the miter planes coincide: a ray through the joint hits exactly one of them."""
    s = oracle.Scene()
    b = s.add_bsdf('kajiyakay')
    xyz, st = tiny_hair()
    s.add_hair(xyz, st, 0.1, b)
    s.set_camera(np.eye(4, dtype=np.float32), 35.0, width=8, height=8)
    s.build()
    xs = np.linspace(0.9, 1.2, 301).astype(np.float32)                 # crosses the joint at x = 1 of the bent fiber
    o = np.stack([xs, np.full_like(xs, 0.02), np.full_like(xs, 5.0)], 1); d = np.tile(np.array([[0, 0, -1]], np.float32), (len(xs), 1))
    sh, pr, t = s.intersect(o, d, 0.0, np.inf, mode=2)
    assert (sh == 0).all() and set(pr) == {1, 2}
    for i in range(len(xs)):
        c = s.candidates(o[i], d[i], 0.0, np.inf)
        assert len(c[1]) == 1, 'ray %d is claimed by %s' % (i, c[1])


def test_segment_bounds_contain_the_segment(oracle):
    s = oracle.Scene()
    b = s.add_bsdf('kajiyakay')
    xyz, st = tiny_hair()
    s.add_hair(xyz, st, 0.1, b)
    s.set_camera(np.eye(4, dtype=np.float32), 35.0, width=8, height=8)
    s.build()
    boxes = s.segment_bounds(0, 6)
    assert np.allclose(boxes[0], [-1, -0.09999, -0.09999, 0, 0.09999, 0.09999], atol=2e-5)      # axis-aligned, radius*(1-Epsilon)
    aabb, _ = s.scene_bounds()
    lo, hi = boxes[:, :3].min(0), boxes[:, 3:].max(0)
    assert (aabb[:3] < lo).all() and (aabb[3:] > hi).all()                 # the tree boxes are the tight union, enlarged (gkdtree.h:1213-1220)
    assert np.allclose(aabb[:3], lo, atol=3e-3 * (1 + np.abs(hi - lo).max())) and np.allclose(aabb[3:], hi, atol=3e-3 * (1 + np.abs(hi - lo).max()))


def test_marschner_fixed_mode(oracle):
    """SURVEY M7, src/bsdfs/marschner.cpp: TRT-only eval, lobe-weighted pdf, sample() = eval/pdf with rejection of pdf > 1."""
    s = oracle.Scene()
    f = s.add_bsdf('marschner_fixed', intIOR=1.55, extIOR=1.0)
    m = s.add_bsdf('marschner', intIOR=1.55, extIOR=1.0)
    tf, tm = s.marschner_tables(f), s.marschner_tables(m)
    assert tf['eta'] == tm['eta'] == np.float32(1.55)
    assert np.array_equal(tf['tables'][0], tm['tables'][0])                 # R does not depend on sigmaA
    assert (tf['tables'][1] >= tm['tables'][1]).all() and (tf['tables'][2] > tm['tables'][2]).any()   # sigmaA 0.22 < 0.5: less absorption
    rng = np.random.default_rng(21)
    n = 4000
    wi = rng.normal(size=(n, 3)); wi /= np.linalg.norm(wi, axis=1, keepdims=True)
    wo = rng.normal(size=(n, 3)); wo /= np.linalg.norm(wo, axis=1, keepdims=True)
    ev, pdf = s.bsdf_eval(f, wi, wo)
    assert np.isfinite(ev).all() and (ev >= 0).all() and np.isfinite(pdf).all() and (pdf >= 0).all()
    assert (ev.sum(axis=1) > 0).mean() > 0.3                                 # no hemisphere test in this class; the narrow M() lobes underflow elsewhere
    smp = rng.random((n, 2)).astype(np.float32); ex = rng.random((n, 4)).astype(np.float32)
    w1, wt1, p1, ty1 = s.bsdf_sample(f, wi, smp, ex)
    w2, wt2, p2, ty2 = s.bsdf_sample(f, wi, rng.random((n, 2)).astype(np.float32), ex)
    assert np.array_equal(w1, w2) and np.array_equal(wt1, wt2)               # the handed-in 2-D sample is ignored, only the extra draws matter
    assert ((ty1 & 0xff) == 0x20).all() and set(np.unique(ty1 >> 8)) <= {0, 1, 2}
    assert np.allclose(np.linalg.norm(w1, axis=1), 1.0, atol=1e-5)
    ev_s, pdf_s = s.bsdf_eval(f, wi, w1)
    assert np.array_equal(pdf_s, p1)
    ok = (p1 > 0) & (p1 <= 1)
    assert ok.mean() > 0.3 and not wt1[~ok].any()                            # pdf > 1 is rejected with a zero weight (:530)
    assert np.allclose(wt1[ok], ev_s[ok] / p1[ok, None], rtol=1e-5, atol=1e-7)


def _upper(rng, n):
    v = rng.normal(size=(n, 3)); v /= np.linalg.norm(v, axis=1, keepdims=True); v[:, 2] = np.abs(v[:, 2])
    return v.astype(np.float32)


@pytest.mark.parametrize('distr', ['ggx', 'beckmann', 'phong'])
def test_roughplastic_consistency(oracle, distr):
    """roughplastic (src/bsdfs/roughplastic.cpp + microfacet.h): sample() weights integrate to the albedo that eval() integrates to,
    pdf() integrates to ~1 and matches the density of the sampled directions, one-sided, energy <= 1."""
    s = oracle.Scene()
    b = s.add_bsdf('roughplastic', intIOR=1.55, extIOR=1.0, alpha=0.2, distribution=distr, diffuseReflectance=(0.143016, 0.0156076, 1.80928e-005))
    rng = np.random.default_rng(31)
    n = 200000
    for wi0 in ([0.0, 0.0, 1.0], [0.5, 0.2, 0.84], [0.95, 0.0, 0.3]):
        wi = np.tile(np.array([wi0], np.float32), (n, 1)); wi /= np.linalg.norm(wi, axis=1, keepdims=True)
        wo, wt, pdf, ty = s.bsdf_sample(b, wi, rng.random((n, 2)).astype(np.float32))
        ok = pdf > 0
        assert set(np.unique(ty[ok] & 0xff)) <= {0x2, 0x8} and (wo[ok, 2] > 0).all()
        u = _upper(rng, n)
        ev, pu = s.bsdf_eval(b, wi, u)
        albedo_eval = (ev * 2 * np.pi).mean(axis=0); albedo_sample = wt.sum(axis=0) / n
        assert np.allclose(albedo_eval, albedo_sample, rtol=0.04, atol=2e-3) and (albedo_eval < 1).all()
        assert abs((pu * 2 * np.pi).mean() - ok.mean()) < 0.03        # pdf mass = fraction of samples that are not rejected below the horizon
        ev2, p2 = s.bsdf_eval(b, wi[ok], wo[ok])
        assert np.array_equal(p2, pdf[ok]) and np.allclose(wt[ok], ev2 / p2[:, None], rtol=1e-5)
    down = np.array([[0.3, 0.1, -0.9]], np.float32)
    assert not s.bsdf_eval(b, down, _upper(rng, 1))[0].any() and not s.bsdf_sample(b, down, [[0.3, 0.6]])[1].any()


# ------------------------------------------------------------------------------------------------ triangle meshes (T1)
def _unit_tri_scene(oracle, tris, pos, normals=None, two_sided=False):
    s = oracle.Scene()
    b = s.add_bsdf('twosided' if two_sided else 'diffuse', reflectance=(0.5, 0.4, 0.3))
    s.add_mesh(pos, tris, b, normals=normals)
    s.set_camera(np.eye(4, dtype=np.float32), 35.0, width=8, height=8)
    s.build()
    return s


def test_triangle_intersection_analytic(oracle):
    """TriAccel (triaccel.h:61-158): plane distance, barycentric acceptance u>=0, v>=0, u+v<=1, inclusive [mint, maxt]."""
    pos = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [5, 5, -5], [5, 5, -5], [5, 5, -5], [-5, -5, 9], [-5, -5, 9], [-5, -5, 9]], np.float32)
    s = _unit_tri_scene(oracle, [[0, 1, 2], [3, 4, 5], [6, 7, 8]], pos)   # the other two are degenerate (k = 3, never hit); they only widen the scene box
    o = np.array([[0.25, 0.25, 2], [0.25, 0.25, -3], [0.9, 0.9, 2], [-0.1, 0.5, 2], [0.5, 0.5, 2], [0.0, 0.0, 2], [5, 5, 7]], np.float32)
    d = np.array([[0, 0, -1], [0, 0, 1], [0, 0, -1], [0, 0, -1], [0, 0, -1], [0, 0, -1], [0, 0, -1]], np.float32)
    sh, pr, t = s.intersect(o, d, 0.0, np.inf, mode=2)
    assert list(sh) == [0, 0, -1, -1, 0, 0, -1]                       # inside (both sides), u+v>1, u<0, on the hypotenuse, on a vertex, degenerate
    assert np.allclose(t[[0, 1, 4, 5]], [2, 3, 2, 2]) and (pr[[0, 1, 4, 5]] == 0).all()
    # the interval is inclusive at both ends (t < mint || t > maxt rejects)
    assert s.intersect(o[:1], d[:1], 2.0, 2.5, mode=2)[0][0] == 0 and s.intersect(o[:1], d[:1], 1.5, 2.0, mode=2)[0][0] == 0
    assert s.intersect(o[:1], d[:1], 2.0001, 10.0, mode=2)[0][0] == -1 and s.intersect(o[:1], d[:1], 0.0, 1.9999, mode=2)[0][0] == -1
    # record: barycentric hit point, face normal (no vertex normals), dpdu = p1 - p0 -> s = x axis
    _, _, _, rec = s.intersect_full(o[:2], d[:2], 0.0, np.inf)
    assert np.allclose(rec[0, :3], [0.25, 0.25, 0], atol=1e-7) and np.allclose(rec[0, 3:6], [0, 0, 1]) and np.allclose(rec[0, 6:9], [1, 0, 0])
    assert np.allclose(rec[0, 12:15], [0, 0, 1]) and np.allclose(rec[1, 12:15], [0, 0, -1])      # wi.z < 0 from behind: one-sided diffuse is black there


def test_mesh_normals_and_bvh_vs_brute(oracle, cp):
    """Interpolated vertex normals; geometric normal flipped to the shading side (skdtree.h:381-391); BVH == brute force; hair + mesh."""
    xyz, idx, nrm = cp.scenes.gen_ellipsoid((0, 0, 0), (1.0, 1.5, 0.8), 10)
    s = oracle.Scene()
    b = s.add_bsdf('diffuse', reflectance=0.5)
    k = s.add_bsdf('kajiyakay')
    s.add_mesh(xyz, idx, b, normals=-nrm)                               # inward normals: geoFrame.n must follow them
    fib = np.array([[-2, 0, 1.5], [0, 0.2, 1.5], [2, 0, 1.5], [-2, 1, -1.5], [2, 1.2, -1.5]], np.float32)
    s.add_hair(fib, np.array([1, 0, 0, 1, 0], np.uint8), 0.05, k)
    s.set_camera(np.eye(4, dtype=np.float32), 35.0, width=8, height=8)
    s.build()
    aabb, _ = s.scene_bounds()
    assert np.allclose(aabb[:3], [-2.0, -1.5, -1.55], atol=2e-2) and np.allclose(aabb[3:], [2.0, 1.5, 1.55], atol=2e-2)   # union of both kinds of shapes
    rng = np.random.default_rng(3)
    o = rng.normal(size=(4000, 3)); o = (4 * o / np.linalg.norm(o, axis=1, keepdims=True)).astype(np.float32)
    d = -o / 4 + 0.35 * rng.normal(size=(4000, 3)); d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    a = s.intersect(o, d, 0.0, np.inf, mode=0); bf = s.intersect(o, d, 0.0, np.inf, mode=2)
    assert np.array_equal(a[0], bf[0]) and np.array_equal(a[1], bf[1]) and np.array_equal(a[2], bf[2])
    assert (a[0] == 0).sum() > 500 and (a[0] == 1).sum() > 20
    assert np.array_equal(s.intersect(o, d, 1e-4, 3.5, mode=1)[0] >= 0, s.intersect(o, d, 1e-4, 3.5, mode=3)[0] >= 0)
    sh, pr, t, rec = s.intersect_full(o, d, 0.0, np.inf)
    m = sh == 0
    P, N, WI = rec[m, :3], rec[m, 3:6], rec[m, 12:15]
    q = P / np.array([1.0, 1.5, 0.8])
    assert np.abs(np.linalg.norm(q, axis=1) - 1).max() < 0.06           # on the faceted ellipsoid
    assert ((N * P).sum(1) < 0).all()                                    # shading normals point inward as given
    assert (WI[:, 2] < 0).mean() > 0.97                                  # rays from outside arrive on the back side (silhouette facets aside)


def test_diffuse_and_twosided_closed_forms(oracle):
    s = oracle.Scene()
    d = s.add_bsdf('diffuse', reflectance=(0.5, 0.4, 0.3)); t2 = s.add_bsdf('twosided', reflectance=(2.0, 1.0, 0.5))   # > 1: rescaled by 0.99/max
    up = np.array([[0.3, 0.1, 0.9]], np.float32); up /= np.linalg.norm(up); down = up * np.array([1, 1, -1], np.float32)
    wo = np.array([[-0.2, 0.4, 0.6]], np.float32); wo /= np.linalg.norm(wo)
    ev, pdf = s.bsdf_eval(d, up, wo)
    assert np.allclose(ev[0], np.array([0.5, 0.4, 0.3]) * wo[0, 2] / np.pi, rtol=1e-6) and np.isclose(pdf[0], wo[0, 2] / np.pi, rtol=1e-6)
    assert not s.bsdf_eval(d, down, wo)[0].any() and not s.bsdf_eval(d, up, wo * np.array([1, 1, -1], np.float32))[0].any()
    ev2, pdf2 = s.bsdf_eval(t2, down, wo * np.array([1, 1, -1], np.float32))
    assert np.allclose(ev2[0], np.array([0.99, 0.495, 0.2475]) * wo[0, 2] / np.pi, rtol=1e-6) and np.isclose(pdf2[0], wo[0, 2] / np.pi, rtol=1e-6)
    assert not s.bsdf_eval(t2, down, wo)[0].any()                        # opposite hemispheres
    smp = np.array([[0.3, 0.7]], np.float32)
    w1, wt1, p1, ty1 = s.bsdf_sample(d, up, smp); w2, wt2, p2, ty2 = s.bsdf_sample(t2, down, smp)
    assert w1[0, 2] > 0 and np.allclose(wt1[0], [0.5, 0.4, 0.3]) and np.isclose(p1[0], w1[0, 2] / np.pi, rtol=1e-6) and ty1[0] == 0x2
    assert np.allclose(w2[0], w1[0] * np.array([1, 1, -1])) and ty2[0] == (0x2 | (1 << 8))      # flipped: component index + 1
    assert not s.bsdf_sample(d, down, smp)[1].any()


def _fresnel_unpolarised(c, eta):
    """Textbook unpolarised Fresnel reflectance from the outside (cos > 0) of a dielectric with relative index eta."""
    st = np.sqrt(max(0.0, 1 - c * c)) / eta
    ct = np.sqrt(max(0.0, 1 - st * st))
    rs = (c - eta * ct) / (c + eta * ct); rp = (eta * c - ct) / (eta * c + ct)
    return 0.5 * (rs * rs + rp * rp)


def test_thindielectric_closed_forms(oracle):
    """src/bsdfs/thindielectric.cpp:143-245: R' = R + T^2 R / (1 - R^2); reflection with probability R', else straight through
    (ENull); non-zero eval/pdf only in the discrete measure; works from both sides."""
    s = oracle.Scene()
    b = s.add_bsdf('thindielectric', intIOR=1.55, extIOR=1.0, specularReflectance=(0.9, 0.5, 0.1), specularTransmittance=(2.0, 1.0, 0.5))   # T > 1: rescaled
    rng = np.random.default_rng(5)
    wi = sphere_dirs(rng, 4096); smp = rng.random((4096, 2)).astype(np.float32)
    wo, wt, pdf, ty = s.bsdf_sample(b, wi, smp)
    for k in range(0, 4096, 97):
        R = _fresnel_unpolarised(abs(float(wi[k, 2])), 1.55); Rp = R + (1 - R) ** 2 * R / (1 - R * R)
        if smp[k, 0] <= np.float32(Rp) - 1e-6:
            assert ty[k] == 0x20 and np.allclose(wo[k], wi[k] * [-1, -1, 1]) and np.isclose(pdf[k], Rp, rtol=2e-5) and np.allclose(wt[k], [0.9, 0.5, 0.1])
        elif smp[k, 0] > np.float32(Rp) + 1e-6:
            assert ty[k] == (0x1 | (1 << 8)) and np.allclose(wo[k], -wi[k]) and np.isclose(pdf[k], 1 - Rp, rtol=2e-5) and np.allclose(wt[k], [0.99, 0.495, 0.2475])
    # discrete measure: eval = colour * probability, pdf = probability, exactly at the two delta directions and nowhere else
    ev, p = s.bsdf_eval(b, wi, wo, discrete=True)
    refl = (ty & 0xff) == 0x20
    assert np.allclose(p, pdf, rtol=1e-6) and np.allclose(ev[refl], np.float32([0.9, 0.5, 0.1]) * pdf[refl, None], rtol=1e-6)
    assert np.allclose(ev[~refl], np.float32([0.99, 0.495, 0.2475]) * pdf[~refl, None], rtol=1e-6)
    off = sphere_dirs(rng, 4096)
    nz = s.bsdf_eval(b, wi, off, discrete=True)[0].any(axis=1)     # DeltaEpsilon = 1e-3: a 2.6 degree cap around each delta direction
    near = (np.abs((wi * [-1, -1, 1] * off).sum(1) - 1) <= 1e-3) | (np.abs((-wi * off).sum(1) - 1) <= 1e-3)
    assert nz.sum() < 20 and (nz == near).mean() > 0.999
    ev0, p0 = s.bsdf_eval(b, wi, wo)                               # solid-angle measure: a delta BSDF evaluates to zero
    assert not ev0.any() and not p0.any()
    assert (s.bsdf_sample(b, wi, smp)[2] > 0).all()                 # back side (wi.z < 0) samples as well


def test_marschnerdielectric_as_committed(oracle):
    """src/bsdfs/marschnerdielectric.cpp as committed: eval() == 0 in every measure, pdf() = cosine density of the diffuse component,
    sample() = thin dielectric with probability (s+t)/(d+s+t) and a zero-weight diffuse branch."""
    s = oracle.Scene()
    d, r, t = (0.3, 0.2, 0.1), (0.4, 0.3, 0.2), (0.5, 0.6, 0.7)
    b = s.add_bsdf('marschnerdielectric', intIOR=1.55, extIOR=1.0, exponent=5.0, diffuseReflectance=d, specularReflectance=r, specularTransmittance=t)
    lum = lambda c: 0.212671 * c[0] + 0.715160 * c[1] + 0.072169 * c[2]
    w = (lum(r) + lum(t)) / (lum(d) + lum(r) + lum(t))
    rng = np.random.default_rng(6)
    wi = sphere_dirs(rng, 4096); wo = sphere_dirs(rng, 4096); smp = rng.random((4096, 2)).astype(np.float32)
    ev, pdf = s.bsdf_eval(b, wi, wo)
    assert not ev.any() and not s.bsdf_eval(b, wi, wo, discrete=True)[0].any() and not s.bsdf_eval(b, wi, wo, discrete=True)[1].any()
    both = (wi[:, 2] > 0) & (wo[:, 2] > 0)
    assert np.allclose(pdf[both], wo[both, 2] / np.pi, rtol=1e-6) and not pdf[~both].any()
    swo, wt, p, ty = s.bsdf_sample(b, wi, smp)
    spec = smp[:, 0] <= np.float32(w)
    assert ((ty[~spec] & 0xff) == 0x2).all() and not wt[~spec].any() and (swo[~spec, 2] >= 0).all()       # dead diffuse branch
    assert np.isin(ty[spec] & 0xff, (0x1, 0x20)).all()
    refl = spec & ((ty & 0xff) == 0x20); thru = spec & ((ty & 0xff) == 0x1)
    assert np.allclose(wt[refl], np.float32(r)) and np.allclose(wt[thru], np.float32(t)) and np.allclose(swo[thru], -wi[thru]) and np.allclose(p[refl] + 0 * p[refl], p[refl])
    k = int(np.nonzero(refl)[0][0])
    R = _fresnel_unpolarised(abs(float(wi[k, 2])), 1.55); Rp = R + (1 - R) ** 2 * R / (1 - R * R)
    assert np.isclose(p[k], Rp, rtol=2e-5)


def test_dielectric_scenes_render_on_the_oracle(cp, oracle):
    """The thin dielectric has no smooth component: no shadow rays (path.cpp:174-175).  marschnerdielectric has one, so every vertex
    draws an emitter sample although its eval() is zero; both images are finite, non-black and differ from each other."""
    films = {}
    for name in ('straight-hair-thindielectric', 'straight-hair-dielectric'):
        ov = dict(width=48, height=48, spp=4)
        env = cp.bake_sunsky(**cp.scenes.sunsky_params(name))
        sc = oracle.scene_from_description(name, scale=0.004, overrides=ov, envmap=env)
        films[name] = sc.render(4, seed=3)
        st = sc.last_stats
        assert st['paths'] == 48 * 48 * 4 and np.isfinite(films[name]).all() and films[name][..., :3].sum() > 0
        assert (st['shadow_rays'] == 0) == (name == 'straight-hair-thindielectric')
    assert not np.allclose(films['straight-hair-thindielectric'], films['straight-hair-dielectric'])


def _write_rgbe(path, rgbe, rle, crlf=False):
    """Radiance RGBE file from an (h, w, 4) uint8 array; rle=True writes new-style run-length encoded scanlines."""
    h, w = rgbe.shape[:2]
    nl = b'\r\n' if crlf else b'\n'
    out = bytearray(b'#?RADIANCE' + nl + b'# made by a test' + nl + b'FORMAT=32-bit_rle_rgbe' + nl + nl + ('-Y %d +X %d' % (h, w)).encode() + nl)
    if not rle:
        out += rgbe.tobytes()
    else:
        for y in range(h):
            out += bytes([2, 2, w >> 8, w & 255])
            for c in range(4):
                row = rgbe[y, :, c]; x = 0
                while x < w:
                    run = 1
                    while x + run < w and run < 127 and row[x + run] == row[x]:
                        run += 1
                    if run >= 3:
                        out += bytes([128 + run, int(row[x])]); x += run
                    else:
                        n = 1
                        while x + n < w and n < 128 and not (x + n + 2 < w and row[x + n] == row[x + n + 1] == row[x + n + 2]):
                            n += 1
                        out += bytes([n]) + row[x:x + n].tobytes(); x += n
    open(path, 'wb').write(bytes(out))


def _ref_load_rgbe(path):
    L = ctypes.CDLL(REF_GEOM)
    w = ctypes.c_int(); h = ctypes.c_int(); err = ctypes.create_string_buffer(256)
    if L.ref_load_rgbe(str(path).encode(), None, ctypes.byref(w), ctypes.byref(h), err) != 0:
        raise RuntimeError(err.value.decode())
    out = np.zeros((h.value, w.value, 3), np.float32)
    L.ref_load_rgbe(str(path).encode(), out.ctypes.data_as(ctypes.c_void_p), ctypes.byref(w), ctypes.byref(h), err)
    return out


def test_rgbe_reader(cp, oracle, tmp_path):
    """Bitmap::readRGBE (src/libcore/bitmap.cpp:3590-3678): flat and run-length encoded files, CR/LF headers, widths that forbid RLE,
    zero exponents; product loader == oracle loader == direct decode, error messages for broken files."""
    rng = np.random.default_rng(12)
    for k, (w, h, rle, crlf) in enumerate([(16, 5, False, False), (16, 5, True, False), (37, 9, True, True), (5, 4, False, False), (300, 3, True, False)]):
        q = rng.integers(0, 256, size=(h, w, 4), dtype=np.uint8)
        q[:, :, 3] = rng.integers(100, 150, size=(h, w), dtype=np.uint8)
        q[0, : w // 2] = q[0, 0]                                   # long runs
        q[h - 1, w // 3:, 3] = 0                                   # zero exponent = black
        if not rle and w >= 8:
            q[0, 0, 0] = 7                                          # a flat file must not start with the RLE marker 2,2
        path = tmp_path / ('t%d.hdr' % k)
        _write_rgbe(path, q, rle, crlf)
        expect = q[..., :3].astype(np.float32) * np.ldexp(np.float32(1), q[..., 3].astype(np.int32) - 136)[..., None]
        expect[q[..., 3] == 0] = 0
        a = cp.load_rgbe(path); b = oracle.load_rgbe(path)
        assert a.shape == (h, w, 3) and np.array_equal(a, expect) and np.array_equal(b, expect)
        if os.path.exists(REF_GEOM):                              # Bitmap::readRGBE itself, cut out of the reference and executed as written
            c = _ref_load_rgbe(path)
            assert c.shape == (h, w, 3) and np.array_equal(c, expect)
    raw = open(path, 'rb').read()
    bad = tmp_path / 'bad.hdr'
    for data, msg in [(b'RADIANCE\n' + raw[11:], 'Invalid header'), (raw.replace(b'FORMAT=32-bit_rle_rgbe', b'FORMAT=32-bit_rle_xyze'), 'invalid format'),
                      (raw[:len(raw) - 40], 'end of file')]:
        bad.write_bytes(data)
        with pytest.raises(cp.CudapathError, match=msg):
            cp.load_rgbe(bad)
        with pytest.raises(RuntimeError, match=msg):
            oracle.load_rgbe(bad)
        if os.path.exists(REF_GEOM):
            with pytest.raises(RuntimeError):
                _ref_load_rgbe(bad)
    with pytest.raises(cp.CudapathError, match='could not be found'):
        cp.load_rgbe(tmp_path / 'nothing.hdr')
    ref = '/root/reference/models/teapot/textures/envmap.hdr'       # the one image file the reference ships for this emitter
    if os.path.exists(ref):
        import hashlib
        import json
        a = cp.load_rgbe(ref)
        gold = json.load(open(os.path.join(GOLDEN, 'envmap_hdr.json')))
        assert list(a.shape) == gold['shape'] and hashlib.sha256(a.tobytes()).hexdigest() == gold['sha256'] and np.array_equal(a, oracle.load_rgbe(ref))
        assert np.array_equal(a, _ref_load_rgbe(ref))


def test_obj_loader(cp, tmp_path):
    """WavefrontOBJ + computeNormals: vertex merge, n-gon fans, negative indices, toWorld on points and normals, generated
    angle-weighted normals, faceNormals / flipNormals (obj.cpp:244-349, 608-700; trimesh.cpp:608-672)."""
    # a unit cube as 6 quads, no normals: 8 merged vertices, 12 triangles, smooth normals along the diagonals
    v = [(x, y, z) for x in (0, 1) for y in (0, 1) for z in (0, 1)]
    quads = [(1, 2, 4, 3), (5, 7, 8, 6), (1, 5, 6, 2), (3, 4, 8, 7), (1, 3, 7, 5), (2, 6, 8, 4)]
    path = tmp_path / 'cube.obj'
    path.write_text('# cube\n' + ''.join('v %d %d %d\n' % p for p in v) + 'g faces\n' + ''.join('f %d %d %d %d\n' % q for q in quads))
    xyz, idx, nrm = cp.load_obj_file(str(path))
    assert xyz.shape == (8, 3) and idx.shape == (12, 3) and nrm.shape == (8, 3)
    p = xyz[idx]; fn = np.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0])
    assert ((fn * (p.mean(1) - 0.5)).sum(1) > 0).all()                      # outward winding preserved by the fans
    assert np.allclose(nrm, (xyz - 0.5) / np.linalg.norm(xyz - 0.5, axis=1, keepdims=True), atol=1e-6)
    _, idx_f, nrm_f = cp.load_obj_file(str(path), faceNormals=True, flipNormals=True)
    assert nrm_f is None and np.array_equal(idx_f[:, [1, 0, 2]], idx)        # flipped winding, no normals
    _, _, nrm_fl = cp.load_obj_file(str(path), flipNormals=True)
    assert np.allclose(nrm_fl, -nrm, atol=1e-6)
    # explicit normals, v//vn corners, negative indices, a scaling + translating toWorld (normals use the inverse transpose)
    path2 = tmp_path / 'tri.obj'
    path2.write_text('v 0 0 0\nv 1 0 0\nv 0 1 0\nvn 0 0 2\nvn 1 0 1\nvt 0.25 0.75\nf -3//1 -2//1 -1//2\nf 1/1/1 2/1/1 3/1/2\n')
    tw = np.array([[2, 0, 0, 5], [0, 1, 0, 0], [0, 0, 4, 0], [0, 0, 0, 1]], np.float32)
    xyz, idx, nrm = cp.load_obj_file(str(path2), toWorld=tw)
    assert xyz.shape == (6, 3) and idx.shape == (2, 3)                       # same position+normal but different uv: not merged
    assert np.array_equal(xyz[:3], np.array([[5, 0, 0], [7, 0, 0], [5, 1, 0]], np.float32))
    assert np.allclose(nrm[0], [0, 0, 1]) and np.allclose(nrm[2], np.array([0.5, 0, 0.25]) / np.hypot(0.5, 0.25), atol=1e-6)
    with pytest.raises(cp.CudapathError):
        cp.load_obj_file(str(tmp_path / 'missing.obj'))
    bad = tmp_path / 'bad.obj'; bad.write_text('v 0 0 0\nf 1 2 3\n')
    with pytest.raises(cp.CudapathError, match='Out of bounds'):
        cp.load_obj_file(str(bad))


def test_compute_normals_pinned_against_reference_text(cp, tmp_path):
    """TriMesh::computeNormals (src/librender/trimesh.cpp:606-676) + unitAngle (include/mitsuba/core/util.h:309-314), cut out of the reference
    at build time and executed as written (oracle/ref_shim/ref_trimesh.cpp), against the PRODUCT's OBJ loader on a bumpy sphere with a
    degenerate triangle and an unreferenced vertex: generated vertex normals, flipNormals, faceNormals (winding swap, no normals).  The
    reference calls libm asinf, the product rounds correctly: normals agree to an ulp or two, most of them to the bit."""
    if not os.path.exists(REF_GEOM):
        pytest.skip('oracle/_ref/libref_geom.so not built (needs /root/reference)')
    L = ctypes.CDLL(REF_GEOM)
    rng = np.random.default_rng(12)
    nu, nv = 48, 24
    th = (np.arange(1, nv) / nv * np.pi)[:, None]; ph = (np.arange(nu) / nu * 2 * np.pi)[None, :]
    r = 1 + 0.15 * rng.random((nv - 1, nu))
    pts = np.stack([r * np.sin(th) * np.cos(ph), r * np.cos(th) * np.ones_like(ph), r * np.sin(th) * np.sin(ph)], -1).reshape(-1, 3)
    pts = np.vstack([pts, [[0, 1.2, 0], [0, -1.2, 0], [9, 9, 9]]]).astype(np.float32)            # poles + a vertex no face uses
    top, bot = len(pts) - 3, len(pts) - 2
    faces = []
    for j in range(nv - 2):
        for i in range(nu):
            a = j * nu + i; b = j * nu + (i + 1) % nu; c = a + nu; d = b + nu
            faces.append((a, b, d, c))                                                            # quads: fan-triangulated by the loader
    for i in range(nu):
        faces.append((top, (i + 1) % nu, i)); faces.append((bot, (nv - 2) * nu + i, (nv - 2) * nu + (i + 1) % nu))
    faces.append((0, 0, 1))                                                                       # degenerate
    path = tmp_path / 'bumpy.obj'
    path.write_text(''.join('v %.9g %.9g %.9g\n' % tuple(p) for p in pts) + ''.join('f ' + ' '.join(str(k + 1) for k in f) + '\n' for f in faces))
    for face, flip in ((False, False), (False, True), (True, True), (True, False)):
        xyz, idx, nrm = cp.load_obj_file(str(path), faceNormals=face, flipNormals=flip)
        xyz0, idx0, _ = cp.load_obj_file(str(path), faceNormals=True, flipNormals=False)          # the mesh before computeNormals touches it
        assert np.array_equal(xyz, xyz0)
        tri = np.ascontiguousarray(idx0, np.uint32).copy(); ref_n = np.zeros_like(xyz)
        has = L.ref_compute_normals(xyz.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(len(xyz)), tri.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(len(tri)),
                                    int(face), int(flip), ref_n.ctypes.data_as(ctypes.c_void_p))
        assert np.array_equal(idx, tri)                                                           # winding after computeNormals
        assert bool(has) == (nrm is not None)
        if nrm is not None:
            assert len(nrm) > 1100 and np.abs(nrm - ref_n).max() <= 4e-7 and (nrm == ref_n).mean() > 0.8
            assert (np.abs(np.linalg.norm(nrm, axis=1) - 1) < 1e-6).all()


def _ref_obj_load(L, path, toWorld=None):
    """WavefrontOBJ text loader of the reference executed as written (oracle/ref_shim/ref_obj.cpp) -> (xyz, idx, normals or None)."""
    tw = np.ascontiguousarray(np.eye(4) if toWorld is None else toWorld, np.float32)
    h = ctypes.c_void_p(); nv = ctypes.c_size_t(); nt = ctypes.c_size_t(); hn = ctypes.c_int(); ht = ctypes.c_int(); err = ctypes.create_string_buffer(512)
    if L.ref_obj_load(str(path).encode(), tw.ctypes.data_as(ctypes.c_void_p), 0, ctypes.byref(h), ctypes.byref(nv), ctypes.byref(nt), ctypes.byref(hn), ctypes.byref(ht),
                      err, ctypes.c_size_t(512)) != 0:
        raise RuntimeError(err.value.decode())
    xyz = np.zeros((nv.value, 3), np.float32); nrm = np.zeros((nv.value, 3), np.float32) if hn.value else None; idx = np.zeros((nt.value, 3), np.uint32)
    L.ref_obj_copy(h, xyz.ctypes.data_as(ctypes.c_void_p), None if nrm is None else nrm.ctypes.data_as(ctypes.c_void_p), None, idx.ctypes.data_as(ctypes.c_void_p))
    L.ref_obj_free(h)
    return xyz, idx, nrm


def test_obj_text_loader_pinned_against_reference_text(cp, tmp_path):
    """The line loop of WavefrontOBJ(props), fetch_line, parse and createMesh (src/shapes/obj.cpp:165-187, 245-328, 371-390, 577-715) with
    tokenize / trim (util.cpp:83-104), Transform(Matrix4x4) (matrix.inl:138-193) and the point / normal transforms (transform.h:108-125,
    203-211), cut out of the reference at build time and executed as written, against the PRODUCT's OBJ loader: a file with every corner
    syntax (v, v/vt, v//vn, v/vt/vn), negative indices, n-gons, duplicated positions, a zero normal, trailing blanks, CR line ends,
    backslash continuation lines, groups / materials (collapsed), under the identity and under a sheared toWorld.  Vertex order after the
    merge, indices and transformed normals must be BIT-identical; malformed files must raise the same message on both sides."""
    if not os.path.exists(REF_GEOM):
        pytest.skip('oracle/_ref/libref_geom.so not built (needs /root/reference)')
    L = ctypes.CDLL(REF_GEOM)
    rng = np.random.default_rng(5)
    n = 300
    v = rng.normal(size=(n, 3)).astype(np.float32); v[50:60] = v[40:50]                           # equal positions: merged when normal and uv agree
    vn = rng.normal(size=(40, 3)).astype(np.float32); vn[7] = 0                                   # a zero normal is kept as it is (obj.cpp:658-659)
    vt = rng.random((30, 2)).astype(np.float32)
    lines = ['# torture', 'mtllib foo.mtl', 'o thing']
    for i, p in enumerate(v):
        lines.append('v %.9g %.9g %.9g' % tuple(p) + ('  \t ' if i % 7 == 0 else '') + ('\r' if i % 5 == 0 else ''))
        if i == 100:
            lines += ['g part_a', 'usemtl red']
    lines += ['vn %.9g %.9g %.9g' % tuple(q) for q in vn] + ['vt %.9g %.9g' % tuple(q) for q in vt] + ['s off']
    for k in range(400):
        ids = rng.integers(1, n + 1, int(rng.integers(3, 7)))
        toks = []
        for a in ids:
            a = int(a) if k % 3 else int(a) - n - 1
            toks.append(('%d' % a, '%d/%d' % (a, rng.integers(1, 31)), '%d//%d' % (a, rng.integers(1, 41)), '%d/%d/%d' % (a, rng.integers(1, 31), -int(rng.integers(1, 41))),
                         '%d/%d/%d' % (a, -int(rng.integers(1, 31)), rng.integers(1, 41)))[k % 5])
        lines += ['f ' + ' '.join(toks[:2]) + ' \\', ' '.join(toks[2:])] if k % 11 == 0 else ['f ' + ' '.join(toks)]
        if k == 150:
            lines += ['g part_b', 'usemtl blue']
    lines += ['   f 1 2 3   ', 'f 4 5', 'f 6', '']                                               # leading blanks; short faces repeat their last corner (obj.cpp:311-314)
    path = tmp_path / 'torture.obj'
    path.write_bytes(('\n'.join(lines) + '\n').encode())
    shear = np.array([[1.5, 0.2, 0, 3], [0.1, 0.8, -0.3, -1], [0, 0.4, 2.0, 0.5], [0, 0, 0, 1]], np.float32)
    for tw in (None, shear):
        rx, ri, rn = _ref_obj_load(L, path, tw)
        assert len(rx) > 1500 and len(ri) > 900 and rn is not None
        px, pi, _ = cp.load_obj_file(str(path), toWorld=tw, faceNormals=True)                     # the mesh before computeNormals touches it
        assert np.array_equal(px.view(np.uint32), rx.view(np.uint32)) and np.array_equal(pi, ri)
        _, pi2, pn = cp.load_obj_file(str(path), toWorld=tw)                                      # given normals are kept
        assert np.array_equal(pi2, ri) and np.array_equal(pn.view(np.uint32), rn.view(np.uint32))
    cases = {'empty face': 'v 0 0 0\nv 1 0 0\nv 0 1 0\nf\n', 'four tokens': 'v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1/1/1/1 2 3\n', 'vertex': 'v 0 0 0\nf 1 2 3\n',
             'normal': 'v 0 0 0\nv 1 0 0\nv 0 1 0\nvn 0 0 1\nf 1//1 2//2 3//1\n', 'uv': 'v 0 0 0\nv 1 0 0\nv 0 1 0\nvt 0 0\nf 1/1 2/5 3/1\n',
             'negative': 'v 0 0 0\nv 1 0 0\nv 0 1 0\nf -1 -2 -4\n'}
    for name, text in cases.items():
        bad = tmp_path / 'bad.obj'; bad.write_text(text)
        with pytest.raises(RuntimeError) as ref_err:
            _ref_obj_load(L, bad)
        with pytest.raises(cp.CudapathError) as prod_err:
            cp.load_obj_file(str(bad))
        assert str(prod_err.value) == str(ref_err.value), name
    ok = tmp_path / 'ok.obj'; ok.write_text('v 0 0 0\nv 1 0 0\nv 0 1 0\nvn 0 0 1\nf 1/ 2// /3')   # stray slashes, no final newline
    assert np.array_equal(cp.load_obj_file(str(ok), faceNormals=True)[1], _ref_obj_load(L, ok)[1])


def test_sun_disc_splat_pinned_against_reference_text(cp):
    """The QMC rasterisation of the sun disc in SunSkyEmitter (src/emitters/sunsky.cpp:180-211) with toSphere / fromSphere (sunmodel.h:90-105),
    squareToUniformCone (warp.cpp:54-63), sample02 (qmc.h:43-60,82-87,115-120) and coordinateSystem (util.cpp:592-601), cut out of the
    reference at build time and executed as written (oracle/ref_shim/ref_sunsplat.cpp), against the PRODUCT's bake with the sky switched off:
    sample count, texels hit and accumulated radiance BIT-identical (one sample up to ~10^5 per texel, sun at the zenith, near the horizon
    and behind the seam).  The sun radiance itself (computeSunRadiance -> RGB) comes from the product and is an input on the reference side."""
    if not os.path.exists(REF_GEOM):
        pytest.skip('oracle/_ref/libref_geom.so not built (needs /root/reference)')
    L = ctypes.CDLL(REF_GEOM)
    L.ref_sun_splat.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_float, ctypes.c_float]
    for sun_dir, turbidity, res, radius_scale in (((0.3, 0.8, 0.2), 3.0, 512, 1.0), ((-0.5, 0.1, -0.7), 5.0, 256, 4.0), ((0, 1, 0), 2.0, 64, 1.0),
                                                  ((0.7, 0.02, 0.1), 8.0, 1024, 10.0), ((-1e-3, 0.5, 1.0), 3.0, 128, 6.0)):
        prod = cp.bake_sunsky(turbidity=turbidity, sunDirection=sun_dir, skyScale=0.0, sunScale=1.5, sunRadiusScale=radius_scale, resolution=res)
        radiance = cp.sun_radiance(turbidity, sun_dir) * np.float32(1.5)
        assert np.isfinite(radiance).all() and (radiance >= 0).all() and radiance[0] > 0          # a sun on the horizon at turbidity 8 is red only
        ref = np.zeros((res // 2, res, 3), np.float32); d = np.asarray(sun_dir, np.float32)
        L.ref_sun_splat(ref.ctypes.data, res, radiance.ctypes.data, d.ctypes.data, 1.0, radius_scale)
        assert (ref.sum(-1) > 0).sum() >= 3
        assert np.array_equal(ref.view(np.uint32), prod.view(np.uint32)), (sun_dir, res)


def test_film_filter_table_and_splat(oracle):
    s = oracle.Scene()
    s.add_hair(*tiny_hair(), 0.1, s.add_bsdf('kajiyakay'))
    s.set_camera(np.eye(4, dtype=np.float32), 35.0, width=8, height=6)
    s.set_film('tent')
    s.build()
    t = s.filter_table()
    assert t[31] == 0 and abs(t[:31].sum() * 2 / 31 - 1.0) < 1e-6 and np.all(np.diff(t[:31]) < 0)    # discretised tent, unit integral
    f = s.splat(np.array([[3.5, 2.5]], np.float32), np.array([[1, 2, 3]], np.float32), np.array([1], np.float32))
    assert f[2, 3, 4] == pytest.approx(t[0] * t[0]) and np.allclose(f[2, 3, :3] / f[2, 3, 4], [1, 2, 3])
    assert np.count_nonzero(f[..., 4]) == 1                              # centred sample: neighbours get index >= 31 -> weight 0
    f = s.splat(np.array([[4.0, 3.0]], np.float32), np.array([[1, 1, 1]], np.float32), np.array([1], np.float32))
    assert np.count_nonzero(f[..., 4]) == 4 and np.allclose(f[2:4, 3:5, 4], t[15] * t[15])
    bad = s.splat(np.array([[4.0, 3.0]] * 3, np.float32), np.array([[np.nan, 0, 0], [-1, 0, 0], [np.inf, 0, 0]], np.float32), np.ones(3, np.float32))
    assert (bad == 0).all()                                              # invalid samples are dropped whole (imageblock.h:148-151)


def test_camera_rays(oracle):
    s = oracle.Scene()
    s.add_hair(*tiny_hair(), 0.1, s.add_bsdf('kajiyakay'))
    cam = np.eye(4, dtype=np.float32); cam[:3, 3] = (1, 2, 3)
    s.set_camera(cam, 90.0, nearClip=0.5, farClip=100.0, width=64, height=32)
    s.build()
    o, d, mn, mx = s.camera_rays(np.array([[32, 16], [0, 16], [64, 16], [32, 0]], np.float32))
    assert np.allclose(o, (1, 2, 3)) and np.allclose(d[0], (0, 0, 1), atol=1e-6) and mn[0] == pytest.approx(0.5) and mx[0] == pytest.approx(100.0)
    assert np.allclose(d[1], np.array([1, 0, 1]) / np.sqrt(2), atol=1e-6) and np.allclose(d[2], np.array([-1, 0, 1]) / np.sqrt(2), atol=1e-6)   # x is mirrored (perspective.cpp:150)
    assert np.allclose(d[3], np.array([0, 0.5, 1]) / np.linalg.norm([0, 0.5, 1]), atol=1e-6) and mn[1] == pytest.approx(0.5 * np.sqrt(2), rel=1e-6)


def test_envmap_sampling_matches_pdf(oracle):
    """Importance sampling consistency: E[value/pdf] over samples == integral of the map; pdfDirect agrees with sampleDirect's pdf."""
    rng = np.random.default_rng(5)
    env = (rng.random((16, 32, 3)) ** 3).astype(np.float32); env[3, 7] = 50
    s = oracle.Scene()
    s.add_hair(*tiny_hair(), 0.1, s.add_bsdf('kajiyakay'))
    s.set_envmap(env)
    s.set_camera(np.eye(4, dtype=np.float32), 35.0, width=8, height=8)
    s.build()
    n = 200000
    d, v, pdf, dist = s.env_sample(np.zeros((n, 3), np.float32), rng.random((n, 2)).astype(np.float32))
    ok = pdf > 0
    assert ok.mean() > 0.99 and np.allclose(np.linalg.norm(d[ok], axis=1), 1, atol=1e-5)
    # the tent offset can carry a sample of the first/last row across the pole, where the reference's (u,v) <-> direction
    # mapping is not one-to-one (envmap.cpp:577,594-599): pdfDirect and sampleDirect legitimately disagree there
    ok &= np.abs(d[:, 1]) < 0.97
    rgb, pdf2 = s.env_eval(d[ok])
    assert np.allclose(pdf2, pdf[ok], rtol=2e-3, atol=1e-6)
    assert np.allclose(rgb, v[ok] * pdf[ok][:, None], rtol=2e-3, atol=1e-4)
    # Monte-Carlo estimate of the map's integral vs a brute-force quadrature of evalEnvironment
    est = v[pdf > 0].mean(0) * (pdf > 0).mean()
    dirs = sphere_dirs(rng, 400000)
    quad = s.env_eval(dirs)[0].mean(0) * 4 * np.pi
    assert np.allclose(est, quad, rtol=0.03)
    rows, cols, rw, nrm = s.env_tables()
    assert rows[0] == 0 and rows[-1] == 1 and np.all(np.diff(rows) >= 0) and np.allclose(cols[:, -1], 1)


# ------------------------------------------------------------------------------------------------ product host code vs oracle (no GPU needed)
def write_ascii(path, xyz, st):
    with open(path, 'w') as f:
        for p, s in zip(xyz, st):
            if s:
                f.write('\n')
            f.write('%.9g %.9g %.9g\n' % tuple(p))


def test_hair_loader_binary_and_ascii(cp, oracle, tmp_path):
    rng = np.random.default_rng(11)
    xyz, st = cp.scenes.gen_curly(strands=200, segments=20)
    # make the loader work for its living: duplicates, nearly collinear runs, a degenerate 1-vertex fiber
    xyz = xyz.copy(); st = st.copy()
    xyz[5] = xyz[4]                                           # exact duplicate -> dropped (hair.cpp:712-714)
    xyz[30:34] = xyz[29] + np.outer(np.arange(1, 5), [0.01, 0.0, 0.0]).astype(np.float32)   # collinear run -> merged (hair.cpp:699-704)
    st[100] = 1
    b = str(tmp_path / 'h.mitshair'); a = str(tmp_path / 'h.txt')
    cp.scenes.write_mitshair(b, xyz, st); write_ascii(a, xyz, st)
    tw = np.array([[2, 0, 0, 1], [0, 0, -2, 0], [0, 2, 0, -3], [0, 0, 0, 1]], np.float32)      # rotation+scale 2 -> radius doubles
    for path in (b, a):
        for kw in (dict(), dict(toWorld=tw, angleThreshold=5.0)):
            pa = cp.load_hair_file(path, radius=0.01, **kw)
            pb = oracle.load_hair_file(path, radius=0.01, **kw)
            assert np.array_equal(pa[0], pb[0]) and np.array_equal(pa[1], pb[1]) and pa[2] == pb[2]
            assert len(pa[1]) < len(st) and pa[1][0] == 1
    assert cp.load_hair_file(b, radius=0.01, toWorld=tw)[2] == pytest.approx(0.02)
    x1 = cp.load_hair_file(b, radius=0.01)[0]; x2 = cp.load_hair_file(a, radius=0.01)[0]
    assert np.array_equal(x1, x2)                              # %.9g round-trips fp32
    with pytest.raises(cp.CudapathError):
        cp.load_hair_file(str(tmp_path / 'missing.mitshair'))
    with pytest.raises(cp.CudapathError):
        cp.load_hair_file(b, reduction=1.5)
    # truncated binary file
    raw = open(b, 'rb').read()
    open(str(tmp_path / 'trunc.mitshair'), 'wb').write(raw[:len(raw) // 2])
    with pytest.raises(cp.CudapathError):
        cp.load_hair_file(str(tmp_path / 'trunc.mitshair'))
    # empty fiber list
    open(str(tmp_path / 'empty.mitshair'), 'wb').write(b'BINARY_HAIR' + struct.pack('<I', 0))
    assert len(cp.load_hair_file(str(tmp_path / 'empty.mitshair'))[1]) == 0


@pytest.mark.skipif(not os.path.exists(REF_GEOM), reason='oracle/_ref/libref_geom.so not built (needs /root/reference)')
def test_hair_loaders_pinned_against_reference_constructor(cp, oracle, tmp_path):
    """HairShape::HairShape(const Properties &) (src/shapes/hair.cpp:609-785), cut out of the reference and executed as written
    (oracle/ref_shim/ref_loader.cpp), against the oracle's loader AND the product's host loader: binary and ASCII files, exact duplicates,
    nearly collinear runs at several angle thresholds, comment / blank-line fiber breaks, a toWorld with rotation, scale and translation, and the
    `reduction` parameter, whose draws come from Mitsuba's Random -- the SFMT-19937 state, recursion and seeding cut out of src/libcore/random.cpp
    into the same shim -- vertices, fiber flags and the scaled radius are bit-identical, and so are the three random streams."""
    L = ctypes.CDLL(REF_GEOM); L.ref_hair_load.restype = ctypes.c_void_p; L.ref_hair_load_reduced.restype = ctypes.c_void_p; L.ref_hair_load_radius.restype = ctypes.c_float
    nr = 5000
    ra = np.zeros(nr, np.float32); L.ref_random_floats(nr, ra.ctypes.data_as(ctypes.c_void_p))
    rb = np.zeros(nr, np.float32); oracle.lib().orc_random_floats(ctypes.c_uint64(5489), ctypes.c_uint64(nr), rb.ctypes.data_as(ctypes.c_void_p))
    rc = np.zeros(nr, np.float32); cp.lib().cudapath_random_floats(ctypes.c_uint64(5489), ctypes.c_uint64(nr), rc.ctypes.data_as(ctypes.c_void_p))
    assert np.array_equal(ra, rb) and np.array_equal(ra, rc) and 0 <= ra.min() and ra.max() < 1 and abs(ra.mean() - 0.5) < 0.02
    rng = np.random.default_rng(61)
    xyz, st = cp.scenes.gen_curly(strands=300, segments=24)
    xyz = xyz.copy(); st = st.copy()
    for k in rng.integers(1, len(st) - 6, 200):               # gentle bends: merged or kept depending on the threshold
        if not st[k:k + 5].any():
            d = xyz[k] - xyz[k - 1]
            for j in range(4):
                xyz[k + j] = xyz[k + j - 1] + d * (1 + 0.01 * j) + rng.normal(size=3).astype(np.float32) * np.float32(2e-3) * np.linalg.norm(d)
    for k in rng.integers(1, len(st) - 1, 60):
        if not st[k]:
            xyz[k] = xyz[k - 1]                                   # exact duplicates
    st[rng.integers(1, len(st), 20)] = 1                         # short fibers, some of one vertex
    b = str(tmp_path / 'h.mitshair'); a = str(tmp_path / 'h.txt')
    cp.scenes.write_mitshair(b, xyz, st); write_ascii(a, xyz, st)
    th = np.deg2rad(30.0); c, s_ = np.cos(th), np.sin(th)
    tw = np.array([[1.5 * c, -1.5 * s_, 0, 0.25], [1.5 * s_, 1.5 * c, 0, -1.0], [0, 0, 1.5, 3.0], [0, 0, 0, 1]], np.float32)
    counts = {}
    for path in (b, a):
        for kw in (dict(), dict(angleThreshold=0.2), dict(toWorld=tw, angleThreshold=5.0), dict(toWorld=tw, radius=0.003), dict(reduction=0.3), dict(toWorld=tw, reduction=0.9),
                   dict(reduction=0.999)):
            radius = kw.get('radius', 0.01); ang = kw.get('angleThreshold', 1.0); m = kw.get('toWorld', np.eye(4, dtype=np.float32)); red = kw.get('reduction', 0.0)
            err = ctypes.create_string_buffer(256)
            m32 = np.ascontiguousarray(m, np.float32)
            h = L.ref_hair_load_reduced(path.encode(), ctypes.c_float(radius), ctypes.c_float(ang), m32.ctypes.data_as(ctypes.c_void_p), ctypes.c_float(red), err)
            assert h, err.value
            h = ctypes.c_void_p(h)
            n = L.ref_hair_load_count(h) - 1                      # the reference appends the sentinel flag, not a vertex: count = vertices
            n = L.ref_hair_load_count(h)
            rx = np.zeros((n, 3), np.float32); rs = np.zeros(n, np.uint8)
            L.ref_hair_load_copy(h, rx.ctypes.data_as(ctypes.c_void_p), rs.ctypes.data_as(ctypes.c_void_p))
            rr = float(L.ref_hair_load_radius(h))
            args = dict(radius=radius, angleThreshold=ang, reduction=red); args.update({k: v for k, v in kw.items() if k == 'toWorld'})
            for loader in (cp.load_hair_file, oracle.load_hair_file):
                px, ps, pr = loader(path, **args)[:3]
                assert np.array_equal(px, rx) and np.array_equal(ps, rs) and np.float32(pr) == np.float32(rr)
            assert 0 < n < len(st)
            if red:
                counts[(path, red)] = int(rs.sum())
    for path in (b, a):      # the first fiber of a file is never drawn for (no marker precedes it); about 1 - reduction of the rest survive
        assert 0.55 * int(st.sum()) < counts[(path, 0.3)] < 0.85 * int(st.sum()) and counts[(path, 0.9)] < 0.2 * int(st.sum()) and 1 <= counts[(path, 0.999)] <= 3
    for bad in (-0.1, 1.0):
        with pytest.raises(cp.CudapathError, match='reduction'):
            cp.load_hair_file(b, reduction=bad)
    # a file shorter than the 11-byte magic: the reference's FileStream throws, so do both loaders
    tiny = str(tmp_path / 'tiny.txt'); open(tiny, 'w').write('0 0 0\n')
    err = ctypes.create_string_buffer(256)
    assert not L.ref_hair_load(tiny.encode(), ctypes.c_float(0.01), ctypes.c_float(1.0), np.eye(4, dtype=np.float32).ctypes.data_as(ctypes.c_void_p), err)
    with pytest.raises(cp.CudapathError, match='11 bytes'):
        cp.load_hair_file(tiny)
    with pytest.raises(RuntimeError, match='11 bytes'):
        oracle.load_hair_file(tiny)


def test_develop(cp):
    film = np.zeros((2, 3, 5), np.float32)
    film[0, 0] = (2, 4, 6, 2, 2); film[1, 2] = (1, 1, 1, 0.5, 0.5)
    rgb = cp.develop(film)
    assert np.allclose(rgb[0, 0], (1, 2, 3)) and np.allclose(rgb[1, 2], (2, 2, 2)) and (rgb[0, 1] == 0).all()      # weight 0 -> 0


def test_develop_ldr(cp):
    """ldrfilm develop: value/weight, 2^exposure, sRGB or pow(1/gamma), round and clamp to 8 bits (fmtconv.cpp:984-995,1104-1160)."""
    rng = np.random.default_rng(41)
    film = np.zeros((6, 7, 5), np.float32)
    film[..., :3] = rng.random((6, 7, 3)).astype(np.float32) * 3
    film[..., 4] = rng.random((6, 7)).astype(np.float32) * 2 + 0.5
    film[0, 0] = (1, 1, 1, 1, 0); film[0, 1] = (50, 0, 1e-4, 1, 1); film[0, 2] = (np.nan, -1, 0.5, 1, 1)      # zero weight, clamp, NaN / negative
    v = film[..., :3] / np.where(film[..., 4:5] != 0, film[..., 4:5], np.inf)
    def q(x):
        with np.errstate(invalid='ignore'):
            y = x * 255.0 + 0.5
            return np.where(y > 0, np.minimum(y, 255.0), 0).astype(np.uint8)     # NaN -> 0
    with np.errstate(invalid='ignore'):
        srgb = np.where(v <= 0.0031308, 12.92 * v, 1.055 * np.power(v.astype(np.float64), 1 / 2.4) - 0.055)
        g22 = np.power((v * 2.0).astype(np.float64), 1 / 2.2)
    a = cp.develop_ldr(film); b = cp.develop_ldr(film, gamma=2.2, exposure=1.0)
    assert a.dtype == np.uint8 and a.shape == (6, 7, 3)
    assert np.abs(a.astype(int) - q(srgb).astype(int)).max() <= 1 and np.abs(b.astype(int) - q(g22).astype(int)).max() <= 1   # fp32 vs fp64 pow at a rounding boundary
    assert tuple(a[0, 0]) == (0, 0, 0) and a[0, 1, 0] == 255 and a[0, 2, 0] == 0 and a[0, 2, 1] == 0
    assert np.array_equal(cp.develop_ldr(film, gamma=1.0), q(v))


def test_develop_pinned_against_reference_text(cp):
    """Film develop against the reference executed as written: undoGamma / applyGamma / convertScalar cut out of src/libcore/fmtconv.cpp
    (:1093-1160) at build time, inside the ESpectrumAlphaWeight -> ERGB pixel loop of :984-995 (oracle/ref_shim/ref_develop.cpp), with what
    LDRFilm::develop passes (gamma, 2^exposure; src/films/ldrfilm.cpp:300-321).  Linear output bit-identical; 8-bit output identical except
    where the reference's libm powf and the product's correctly rounded pow land on different sides of a rounding boundary (<= 1 level)."""
    if not os.path.exists(REF_GEOM):
        pytest.skip('oracle/_ref/libref_geom.so not built (needs /root/reference)')
    L = ctypes.CDLL(REF_GEOM)
    rng = np.random.default_rng(77)
    h, w = 256, 384
    film = np.zeros((h, w, 5), np.float32)
    film[..., 4] = rng.random((h, w)).astype(np.float32) * 60 + 4                         # filter weights of ~64 spp
    film[..., :3] = (rng.random((h, w, 3)) ** 3).astype(np.float32) * film[..., 4:5] * 1.5    # values in [0, 1.5): dark tones, mid tones, clipped highlights
    film[..., 3] = film[..., 4]
    film[0, :8, 4] = 0                                                                    # pixels no sample touched
    film[1, :8, :3] = 0; film[2, :4, 0] = -1.0; film[2, 4:8, 1] = np.inf; film[3, :4, 2] = np.nan
    film[4, :16, 0] = np.linspace(0.0031308 - 2e-6, 0.0031308 + 2e-6, 16, dtype=np.float32) * film[4, :16, 4]   # around the sRGB knee
    n = h * w
    ref_lin = np.zeros((h, w, 3), np.float32)
    L.ref_develop_hdr(film.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(n), ref_lin.ctypes.data_as(ctypes.c_void_p))
    assert np.array_equal(cp.develop(film), ref_lin, equal_nan=True)
    for gamma, exposure in ((-1.0, 0.0), (2.2, 0.0), (2.2, 1.0), (1.0, -0.5), (1.8, 0.25)):
        ref8 = np.zeros((h, w, 3), np.uint8)
        L.ref_develop_ldr(film.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(n), ctypes.c_float(gamma), ctypes.c_float(exposure), ref8.ctypes.data_as(ctypes.c_void_p))
        got = cp.develop_ldr(film, gamma=gamma, exposure=exposure)
        diff = np.abs(got.astype(int) - ref8.astype(int))
        assert diff.max() <= 1 and (diff != 0).mean() < 1e-4, (gamma, exposure, int(diff.max()), float((diff != 0).mean()))
        if gamma == 1.0:
            assert np.array_equal(got, ref8)


def test_c_abi_exports_every_declared_symbol(cp):
    """The shared library loads (without a GPU) and exports every function include/cudapath.h declares."""
    import re
    hdr = open(os.path.join(os.path.dirname(GOLDEN), '..', 'include', 'cudapath.h')).read()
    names = sorted(set(re.findall(r'\b(cudapath_[a-z0-9_]+)\s*\(', hdr)))
    assert len(names) > 35
    L = cp.lib()
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_header_is_plain_c_and_cli_fails_loudly_without_gpu(cp, tmp_path):
    """include/cudapath.h compiles as C99 on its own; the native front end (csrc/cp_cli.cpp) refuses to run without a CUDA device."""
    import shutil
    import subprocess
    hdr = os.path.join(os.path.dirname(GOLDEN), '..', 'include', 'cudapath.h')
    if shutil.which('gcc'):
        assert subprocess.run(['gcc', '-std=c99', '-Wall', '-Werror', '-fsyntax-only', '-x', 'c', hdr], capture_output=True).returncode == 0
    assert os.path.exists(cp.CLI_PATH)
    r = subprocess.run([cp.CLI_PATH, '-h'], capture_output=True, text=True)
    assert r.returncode == 0 and 'scene.xml' in r.stdout
    import torch
    if not torch.cuda.is_available():
        path = cp.scenes.write_scene('straight-hair', str(tmp_path), scale=0.002)
        r = subprocess.run([cp.CLI_PATH, '-o', str(tmp_path / 'o.png'), path], capture_output=True, text=True)
        assert r.returncode == 1 and 'no CUDA device' in r.stderr and not (tmp_path / 'o.png').exists()


def test_no_gpu_fails_loudly(cp):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip('a GPU is present')
    with pytest.raises(cp.CudapathError, match='CUDA'):
        cp.Context(0)


def test_scene_xml_text(cp):
    txt = cp.scenes.scene_xml('hair-curl')
    assert txt.count('<shape type="hair">') == 4 and txt.count('<bsdf type="marschner"') == 4 and 'version="0.6.0"' in txt
    assert cp.scenes.SCENES['straight-hair']['shapes'][0]['bsdf']['type'] == 'kajiyakay'


def test_lookat_without_up_and_degenerate_cases(cp, tmp_path):
    """<lookat> as scenehandler.cpp:362-398 + Transform::lookAt (transform.cpp:191-214) treat it: no `up` (or a zero one) picks an axis with
    coordinateSystem(); coinciding origin / target and an `up` parallel to the view direction raise the reference's messages (dry run: no GPU)."""
    def scene(lookat):
        p = tmp_path / 'l.xml'
        p.write_text('<scene version="0.6.0"><sensor type="perspective"><transform name="toWorld">%s</transform><film type="ldrfilm"/></sensor>'
                     '<bsdf type="diffuse" id="d"/><shape type="rectangle"><ref id="d"/></shape></scene>' % lookat)
        return str(p)
    assert 'sensor perspective' in cp.validate_scene_xml(scene('<lookat origin="0,0,5" target="0,0,0"/>'))
    assert 'sensor perspective' in cp.validate_scene_xml(scene('<lookat origin="1,2,3" target="0,0,0" up="0,0,0"/>'))
    with pytest.raises(cp.CudapathError, match='coincide'):
        cp.validate_scene_xml(scene('<lookat origin="1,2,3" target="1,2,3"/>'))
    with pytest.raises(cp.CudapathError, match='linearly independent'):
        cp.validate_scene_xml(scene('<lookat origin="0,0,0" target="0,2,0" up="0,1,0"/>'))
    with pytest.raises(cp.CudapathError, match="invalid 'up'"):
        cp.validate_scene_xml(scene('<lookat origin="0,0,0" target="0,2,0" up="0,1"/>'))


# ------------------------------------------------------------------------------------------------ pinning against the reference's own plugin sources
REF_BSDF = os.path.join(os.path.dirname(GOLDEN), '..', 'oracle', '_ref', 'libref_bsdf.so')
needs_ref_bsdf = pytest.mark.skipif(not os.path.exists(REF_BSDF), reason='oracle/_ref/libref_bsdf.so not built (needs /root/reference)')


class RefBSDF:
    """src/bsdfs/{kajiyakay,thindielectric,marschnerdielectric}.cpp compiled unmodified (oracle/Makefile `ref`, oracle/ref_shim/ref_bsdf.cpp)."""
    def __init__(self, plugin, floats=None, spectra=None):
        self.L = ctypes.CDLL(REF_BSDF)
        self.L.ref_bsdf_create.restype = ctypes.c_void_p
        floats = floats or {}; spectra = spectra or {}
        fn = (ctypes.c_char_p * max(1, len(floats)))(*[k.encode() for k in floats]); fv = np.array(list(floats.values()) or [0], np.float32)
        sn = (ctypes.c_char_p * max(1, len(spectra)))(*[k.encode() for k in spectra]); sv = np.array([c for v in spectra.values() for c in v] or [0], np.float32)
        self.h = ctypes.c_void_p(self.L.ref_bsdf_create(plugin.encode(), len(floats), fn, fv.ctypes.data_as(ctypes.c_void_p), len(spectra), sn, sv.ctypes.data_as(ctypes.c_void_p)))
        assert self.h.value

    def eval(self, wi, wo, discrete=False):
        n = len(wi); ev = np.zeros((n, 3), np.float32); pdf = np.zeros(n, np.float32)
        wi = np.ascontiguousarray(wi, np.float32); wo = np.ascontiguousarray(wo, np.float32)
        self.L.ref_bsdf_eval(self.h, n, wi.ctypes.data_as(ctypes.c_void_p), wo.ctypes.data_as(ctypes.c_void_p), 4 if discrete else 1, ev.ctypes.data_as(ctypes.c_void_p), pdf.ctypes.data_as(ctypes.c_void_p))
        return ev, pdf

    def sample(self, wi, smp, extra=None):
        n = len(wi); wo = np.zeros((n, 3), np.float32); wt = np.zeros((n, 3), np.float32); pdf = np.zeros(n, np.float32); ty = np.zeros(n, np.int32)
        wi = np.ascontiguousarray(wi, np.float32); smp = np.ascontiguousarray(smp, np.float32)
        ex = None if extra is None else np.ascontiguousarray(extra, np.float32)
        self.L.ref_bsdf_sample_ex(self.h, n, wi.ctypes.data_as(ctypes.c_void_p), smp.ctypes.data_as(ctypes.c_void_p), None if ex is None else ex.ctypes.data_as(ctypes.c_void_p),
                                  wo.ctypes.data_as(ctypes.c_void_p), wt.ctypes.data_as(ctypes.c_void_p), pdf.ctypes.data_as(ctypes.c_void_p), ty.ctypes.data_as(ctypes.c_void_p))
        return wo, wt, pdf, ty

    def type(self):
        return int(self.L.ref_bsdf_type(self.h))


def _ulp_diff(a, b):
    """Distance in units in the last place between two float32 arrays (same-sign finite values; 0 where bit-identical)."""
    ia = np.ascontiguousarray(a, np.float32).view(np.int32).astype(np.int64); ib = np.ascontiguousarray(b, np.float32).view(np.int32).astype(np.int64)
    ia = np.where(ia < 0, -(ia & 0x7fffffff), ia); ib = np.where(ib < 0, -(ib & 0x7fffffff), ib)
    return np.abs(ia - ib)


@needs_ref_bsdf
@pytest.mark.parametrize('plugin,props', [
    ('kajiyakay', dict(diffuseReflectance=HAIR_RGB, exponent=10.0)),                                                      # C1 block, models/straight-hair/scene_kkay.xml
    ('kajiyakay', dict(diffuseReflectance=(0.7, 0.6, 0.5), specularReflectance=(0.6, 0.6, 0.6), exponent=30.0)),           # energy-conservation rescale
    ('kajiyakay', dict()),                                                                                                  # plugin defaults
    ('thindielectric', dict(intIOR=1.55, extIOR=1.0, specularReflectance=HAIR_RGB, specularTransmittance=HAIR_RGB)),       # scene_thindielectric.xml
    ('thindielectric', dict(specularReflectance=(0.9, 0.5, 0.1), specularTransmittance=(2.0, 1.0, 0.5))),
    ('marschnerdielectric', dict(intIOR=1.55, extIOR=1.0, exponent=5.0, specularTransmittance=HAIR_RGB, specularReflectance=HAIR_RGB, diffuseReflectance=HAIR_RGB)),   # scene_dielectric.xml
    ('marschnerdielectric', dict(diffuseReflectance=(0.3, 0.2, 0.1), specularReflectance=(0.4, 0.3, 0.2), specularTransmittance=(1.5, 0.6, 0.7)))])
def test_oracle_pinned_against_compiled_reference_plugins(oracle, plugin, props):
    """The oracle's KajiyaKay / ThinDielectric / MarschnerDielectric against the reference's own plugin sources, compiled unmodified from
    /root/reference (scaffolding only for the library interfaces): identical decisions (sampled component, zero / non-zero), values equal to
    the last bit wherever no libm call is involved and within 2 ulp where the reference calls powf / sincosf (the oracle rounds correctly)."""
    floats = {k: float(v) for k, v in props.items() if not isinstance(v, tuple)}
    spectra = {k: v for k, v in props.items() if isinstance(v, tuple)}
    ref = RefBSDF(plugin, floats, spectra)
    s = oracle.Scene()
    q = dict(props)
    if plugin == 'thindielectric':
        q.setdefault('intIOR', 1.5046); q.setdefault('extIOR', 1.000277)
    if plugin == 'marschnerdielectric':
        q.setdefault('intIOR', 1.501); q.setdefault('extIOR', 1.000277)
    b = s.add_bsdf(plugin, **q)
    assert ref.type() & 0x1ff == {'kajiyakay': 0x2 | 0x8, 'thindielectric': 0x1 | 0x20, 'marschnerdielectric': 0x1 | 0x20 | 0x2}[plugin]
    rng = np.random.default_rng(41)
    n = 200000
    wi = sphere_dirs(rng, n); wo = sphere_dirs(rng, n); smp = rng.random((n, 2)).astype(np.float32)
    wi[:100, 2] = 0.0; wi[100:200] = [0, 0, 1]; wo[:50] = wi[:50] * [-1, -1, 1]
    # ---- sample
    rwo, rwt, rpdf, rty = ref.sample(wi, smp)
    owo, owt, opdf, oty = s.bsdf_sample(b, wi, smp)
    alive = (rwt != 0).any(axis=1)
    assert np.array_equal(alive, (owt != 0).any(axis=1))
    # a sample that yields nothing leaves the record's type fields in whatever state the code path left them; compare where it matters
    assert np.array_equal(rty[alive], oty[alive])
    assert np.abs(rwo[alive] - owo[alive]).max() <= 2e-5 and (rwo[alive] == owo[alive]).mean() > 0.85          # unit vectors: absolute error (sincosf / powf vs correctly rounded)
    # kajiyakay: weight and pdf are eval / pdf at the sampled direction, and sqrt(1 - pow(y, 2/(e+1))) of the Phong lobe amplifies one ulp of
    # powf near y = 1 -- so the functions are compared sharply below at IDENTICAL directions (the reference's own wo), and here in the bulk
    assert (rwt[alive] == owt[alive]).mean() > 0.8 and np.median(np.abs(rpdf[alive] - opdf[alive])) == 0
    assert np.allclose(rwt[alive], owt[alive], rtol=5e-2, atol=0)
    if plugin != 'kajiyakay':                                   # no libm on these paths: bit-identical
        assert np.array_equal(rwt[alive], owt[alive]) and np.array_equal(rpdf[alive], opdf[alive]) and np.array_equal(rwo[alive], owo[alive])
    # ---- eval / pdf at random pairs and at the reference's own sampled directions, both measures
    for w2 in (wo, rwo):
        for discrete in (False, True):
            rev, rp = ref.eval(wi, w2, discrete)
            oev, op = s.bsdf_eval(b, wi, w2, discrete=discrete)
            assert np.array_equal(rev != 0, oev != 0) and np.array_equal(rp != 0, op != 0)
            assert np.allclose(rev, oev, rtol=2e-6, atol=0) and np.allclose(rp, op, rtol=2e-6, atol=0)
            if plugin != 'kajiyakay':
                assert np.array_equal(rev, oev) and np.array_equal(rp, op)
            else:
                assert (_ulp_diff(rev, oev) == 0).mean() > 0.98 and _ulp_diff(rev, oev).max() <= 4 and _ulp_diff(rp, op).max() <= 4 and (_ulp_diff(rp, op) == 0).mean() > 0.98


@needs_ref_bsdf
@pytest.mark.parametrize('props', [
    dict(intIOR=1.55, extIOR=1.0, specularReflectance=(0.592384, 0.32628, 0.0528657)),                                  # C2: models/hair-curl/marschner_scene.xml (blonde), beckmann 0.1
    dict(intIOR=1.55, extIOR=1.0, alpha=0.2, distribution='ggx', diffuseReflectance=HAIR_RGB),                           # C3 / C4: models/straight-hair/scene_marschner.xml:31-39
    dict(nonlinear=True, diffuseReflectance=(0.3, 0.2, 0.1))])                                                            # plugin defaults (bk7 / air), nonlinear diffuse term
def test_oracle_marschner_pinned_against_compiled_reference_plugin(oracle, props):
    """The as-built `marschner` plugin: src/bsdfs/marschner_diffuse.cpp compiled UNMODIFIED from /root/reference (with its own microfacet.h,
    rtrans.h, the two Tungsten headers and src/libcore/spline.cpp; data/microfacet/*.dat read by its own loader) against the oracle's
    Marschner.  The reference calls libm (expf, logf, asinf, atan2f ...), the oracle rounds correctly, and M() turns one ulp into 3e-5,
    so values agree to the 1e-4 of BASELINE.json's north_star rather than to the bit; decisions (lobe, flags, zero / non-zero) are identical."""
    os.environ['REF_DATA_DIR'] = oracle.DATA_DIR
    distr = {'beckmann': 0.0, 'ggx': 1.0, 'phong': 2.0}
    floats = {k: (distr[v] if k == 'distribution' else float(v)) for k, v in props.items() if not isinstance(v, tuple)}
    spectra = {k: v for k, v in props.items() if isinstance(v, tuple)}
    ref = RefBSDF('marschner', floats, spectra)
    s = oracle.Scene()
    q = dict(props); q.setdefault('intIOR', 1.5046); q.setdefault('extIOR', 1.000277)
    b = s.add_bsdf('marschner', **q)
    rng = np.random.default_rng(43)
    n = 100000
    wi = sphere_dirs(rng, n); wo = sphere_dirs(rng, n); smp = rng.random((n, 2)).astype(np.float32)
    rev, rp = ref.eval(wi, wo); oev, op = s.bsdf_eval(b, wi, wo)
    assert np.array_equal((rev != 0).any(axis=1), (oev != 0).any(axis=1)) and np.array_equal(rp, op)        # pdf == 1 (or 0): the quirk, bit for bit
    scale = np.maximum(np.abs(oev).max(axis=1, keepdims=True), 1e-6)
    err = np.abs(rev - oev) / scale
    assert err.max() < 1e-4, 'eval: max relative error %.3g' % err.max()
    assert np.median(err) < 2e-7
    rwo, rwt, rpdf, rty = ref.sample(wi, smp)
    owo, owt, opdf, oty = s.bsdf_sample(b, wi, smp)
    same = (rty == oty) & ((rwt != 0).any(axis=1) == (owt != 0).any(axis=1))
    assert same.mean() > 0.9995, 'only %.5f of the samples take the same decisions' % same.mean()       # a lobe boundary moved by an ulp of the table sums
    assert np.abs(rwo[same] - owo[same]).max() < 2e-3 and np.median(np.abs(rwo[same] - owo[same])) < 1e-6
    assert np.array_equal(rpdf[same], opdf[same])
    werr = np.abs(rwt[same] - owt[same]) / np.maximum(np.abs(owt[same]).max(axis=1, keepdims=True), 1e-6)
    assert np.median(werr) < 1e-6 and (werr < 1e-3).mean() > 0.999
    # the functions at IDENTICAL inputs: eval at the reference's own sampled directions
    rev2, _ = ref.eval(wi, rwo); oev2, _ = s.bsdf_eval(b, wi, rwo)
    err2 = np.abs(rev2 - oev2) / np.maximum(np.abs(oev2).max(axis=1, keepdims=True), 1e-6)
    assert err2.max() < 1e-4
    # a first random number of exactly 0 (one in 2^24 draws; it happens once in the first sample index of the full curly-hair scene): the
    # reference's sample() returns wo = (0, inf, 0) with an infinite weight, Li becomes NaN and ImageBlock::put drops the sample
    smp0 = smp[:512].copy(); smp0[:, 0] = 0.0
    rwo0, rwt0, _, _ = ref.sample(wi[:512], smp0); owo0, owt0, _, _ = s.bsdf_sample(b, wi[:512], smp0)
    assert np.array_equal(np.isfinite(rwo0), np.isfinite(owo0)) and np.array_equal(np.isfinite(rwt0), np.isfinite(owt0))
    assert not np.isfinite(rwt0).all()


@needs_ref_bsdf
@pytest.mark.parametrize('plugin,props', [
    ('diffuse', dict(reflectance=(0.5, 0.4, 0.3))), ('twosided', dict(reflectance=(1.5, 0.4, 0.3))),
    ('roughplastic', dict(intIOR=1.55, extIOR=1.0, alpha=0.2, distribution='ggx', diffuseReflectance=HAIR_RGB)),            # default BSDF of models/*/scene.xml
    ('roughplastic', dict(alpha=0.1, distribution='beckmann', nonlinear=True, diffuseReflectance=(0.6, 0.5, 0.4))),
    ('roughplastic', dict(intIOR=1.55, extIOR=1.0, alpha=0.3, distribution='phong', diffuseReflectance=(0.2, 0.3, 0.4))),
    ('marschner_fixed', dict(intIOR=1.55, extIOR=1.000277)),
    # the scene-driven full-lobe mode: marschner.cpp with the five edits listed in oracle/Makefile (R and TT kept in eval, sigmaA / betaR / scale angle from the scene)
    ('marschner_full', dict(intIOR=1.55, extIOR=1.000277)),
    ('marschner_full', dict(intIOR=1.5, extIOR=1.0, sigmaAr=0.6, sigmaAg=0.9, sigmaAb=1.6, betaR=0.17, scaleAngleRad=-0.05))])
def test_oracle_mesh_and_next_row_bsdfs_pinned_against_compiled_reference_plugins(oracle, plugin, props):
    """src/bsdfs/{diffuse,twosided,roughplastic,marschner}.cpp compiled unmodified (marschner.cpp is the class the fork's build leaves out:
    the `fixed` mode, with its two extra sampler draws) against the oracle restatements; math::erf / erfinv / hypot2 behind the Beckmann
    visible-normal sampling are the reference's own (src/libcore/math.cpp:25-86, cut out at build time)."""
    os.environ['REF_DATA_DIR'] = oracle.DATA_DIR
    distr = {'beckmann': 0.0, 'ggx': 1.0, 'phong': 2.0}
    floats = {k: (distr[v] if k == 'distribution' else float(v)) for k, v in props.items() if not isinstance(v, tuple)}
    spectra = {k: v for k, v in props.items() if isinstance(v, tuple)}
    ref = RefBSDF(plugin, floats, spectra)
    s = oracle.Scene()
    q = dict(props)
    if plugin == 'roughplastic':
        q.setdefault('intIOR', 1.49); q.setdefault('extIOR', 1.000277)
    if plugin == 'marschner_full':
        q['sigmaA'] = (q.pop('sigmaAr', 0.22), q.pop('sigmaAg', 0.22), q.pop('sigmaAb', 0.22))
    b = s.add_bsdf(plugin, **q)
    rng = np.random.default_rng(47)
    n = 100000
    wi = sphere_dirs(rng, n); wo = sphere_dirs(rng, n); smp = rng.random((n, 2)).astype(np.float32); extra = rng.random((n, 4)).astype(np.float32)
    rev, rp = ref.eval(wi, wo); oev, op = s.bsdf_eval(b, wi, wo)
    assert np.array_equal((rev != 0).any(axis=1), (oev != 0).any(axis=1)) and np.array_equal(rp != 0, op != 0)
    tol = {'marschner_fixed': 3e-4, 'marschner_full': 3e-4, 'roughplastic': 1e-4}.get(plugin, 0.0)       # M() of the Marschner model turns one ulp of a libm call into 3e-5 (cp_bsdf.cuh)
    err = np.abs(rev - oev) / np.maximum(np.abs(oev).max(axis=1, keepdims=True), 1e-6)
    perr = np.abs(rp - op) / np.maximum(np.abs(op), 1e-6)
    assert err.max() <= tol and perr.max() <= tol, '%s: eval %.3g pdf %.3g' % (plugin, err.max(), perr.max())
    assert np.median(err) < 2e-7 and np.median(perr) < 2e-7
    rwo, rwt, rpdf, rty = ref.sample(wi, smp, extra)
    owo, owt, opdf, oty = s.bsdf_sample(b, wi, smp, extra)
    alive = (rwt != 0).any(axis=1)
    same = (alive == (owt != 0).any(axis=1)) & (~alive | (rty == oty))
    assert same.mean() > 0.999, '%s: only %.5f of the samples take the same decisions' % (plugin, same.mean())
    ok = same & alive
    assert np.median(np.abs(rwo[ok] - owo[ok])) < 1e-6 and (np.abs(rwo[ok] - owo[ok]).max(axis=1) < 1e-3).mean() > 0.999
    werr = np.abs(rwt[ok] - owt[ok]) / np.maximum(np.abs(owt[ok]).max(axis=1, keepdims=True), 1e-6)
    assert np.median(werr) < 1e-6 and (werr < 1e-3).mean() > 0.995
    if plugin in ('diffuse', 'twosided'):
        assert np.array_equal(rwt[ok], owt[ok]) and np.allclose(rpdf[ok], opdf[ok], rtol=2e-6, atol=1e-6) and np.abs(rwo[ok] - owo[ok]).max() < 1e-5      # sincosf vs correctly rounded




@pytest.mark.skipif(not os.path.exists(REF_GEOM), reason='oracle/_ref/libref_geom.so not built (needs /root/reference)')
def test_oracle_hair_geometry_pinned_against_reference_text(cp, oracle):
    """HairKDTree::intersect (the FP64 mitred-cylinder test with its miter helpers, hair.cpp:480-596), getAABB(index) (:246-286,368-397),
    HairShape::fillIntersectionRecord (:825-862) and solveQuadraticDouble (util.cpp:487-525), cut out of the reference by line pattern at
    build time and executed as written (oracle/ref_shim/ref_geom.cpp), against the oracle: hit / miss decisions, t, stored hit points, segment
    bounds and intersection frames are bit-identical."""
    L = ctypes.CDLL(REF_GEOM); L.ref_hair_create.restype = ctypes.c_void_p
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rng = np.random.default_rng(53)
    for name, scale in (('curly-hair', 0.002), ('straight-hair', 0.004), ('furball', 0.004)):
        sh = cp.scenes.SCENES[name]['shapes'][0]
        xyz, starts = cp.scenes.generate(sh, scale)
        xyz = np.ascontiguousarray(xyz, np.float32); starts = np.ascontiguousarray(starts, np.uint8)
        radius = float(sh['radius'])
        h = ctypes.c_void_p(L.ref_hair_create(P(xyz), P(starts), len(starts), ctypes.c_float(radius)))
        nseg = L.ref_hair_segment_count(h)
        segs = np.zeros(nseg, np.uint32); L.ref_hair_segments(h, P(segs))
        s = oracle.Scene()
        b = s.add_bsdf('kajiyakay')
        s.add_hair(xyz, starts, radius, b)
        s.set_camera(np.eye(4, dtype=np.float32), 35.0, width=16, height=16); s.build()
        # ---- segment bounds
        rb = np.zeros((nseg, 6), np.float32); L.ref_hair_segment_bounds(h, P(rb))
        ob = s.segment_bounds(0, nseg)
        assert np.array_equal(rb, ob.reshape(-1, 6))
        # ---- cylinder test: rays aimed at (and near) random segments from all around, from far away, from on the surface and from inside
        n = 200000
        iv = segs[rng.integers(0, nseg, n)]
        a = xyz[iv]; bb = xyz[iv + 1]
        target = a + (bb - a) * rng.random((n, 1)).astype(np.float32) + (rng.normal(size=(n, 3)) * radius * 0.8).astype(np.float32)
        dist = np.where(rng.random(n) < 0.3, rng.random(n) * 3 * radius, 0.5 + 20 * rng.random(n)).astype(np.float32)
        d = sphere_dirs(rng, n)
        o = (target - d * dist[:, None]).astype(np.float32)
        mint = np.where(rng.random(n) < 0.5, 0.0, 1e-4).astype(np.float32)
        maxt = np.where(rng.random(n) < 0.2, dist, np.inf).astype(np.float32)
        rhit = np.zeros(n, np.int32); rt = np.zeros(n, np.float32); rp = np.zeros((n, 3), np.float32)
        L.ref_hair_intersect(h, n, P(o), P(d), P(iv), P(mint), P(maxt), P(rhit), P(rt), P(rp))
        ohit, ot, op = s.segment_intersect(0, o, d, iv, mint, maxt)
        assert 0.2 < rhit.mean() < 0.95
        assert np.array_equal(rhit, ohit) and np.array_equal(rt, ot) and np.array_equal(rp, op)
        # ---- intersection records at the stored hit points
        k = np.nonzero(rhit)[0][:50000]
        rr = np.zeros((len(k), 12), np.float32); L.ref_hair_records(h, len(k), P(np.ascontiguousarray(iv[k])), P(np.ascontiguousarray(rp[k])), P(rr))
        orr = s.segment_records(0, iv[k], rp[k])
        assert np.array_equal(rr, orr)


@pytest.mark.skipif(not os.path.exists(REF_GEOM), reason='oracle/_ref/libref_geom.so not built (needs /root/reference)')
def test_oracle_triangle_and_box_tests_pinned_against_reference_text(oracle):
    """TriAccel::load / rayIntersect (include/mitsuba/render/triaccel.h:37-158) and TAABB::rayIntersect (include/mitsuba/core/aabb.h:308-338),
    cut out of the reference headers and executed as written, against the oracle: bit-identical, degenerate triangles, axis-parallel rays and
    rays starting inside the box included."""
    L = ctypes.CDLL(REF_GEOM); O = oracle.lib()
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rng = np.random.default_rng(59)
    n = 300000
    A = rng.normal(size=(n, 3)).astype(np.float32); B = (A + rng.normal(size=(n, 3)) * 0.5).astype(np.float32); C = (A + rng.normal(size=(n, 3)) * 0.5).astype(np.float32)
    C[:100] = B[:100]; C[100:200] = A[100:200] + 2 * (B[100:200] - A[100:200])                       # degenerate: zero area / collinear
    target = (A + (B - A) * rng.random((n, 1)) * 0.7 + (C - A) * rng.random((n, 1)) * 0.7).astype(np.float32)
    d = sphere_dirs(rng, n); dist = (0.1 + 10 * rng.random(n)).astype(np.float32)
    o = (target - d * dist[:, None]).astype(np.float32)
    mint = np.where(rng.random(n) < 0.5, 0.0, 1e-4).astype(np.float32); maxt = np.where(rng.random(n) < 0.2, dist, np.inf).astype(np.float32)
    out = []
    for fn in (L.ref_triaccel, O.orc_triaccel_batch):
        acc = np.zeros((n, 10), np.float32); hit = np.zeros(n, np.int32); tuv = np.zeros((n, 3), np.float32)
        args = (P(A), P(B), P(C), P(o), P(d), P(mint), P(maxt), P(acc), P(hit), P(tuv))
        fn(n, *args) if fn is L.ref_triaccel else fn(ctypes.c_uint64(n), *args)
        out.append((acc, hit, tuv))
    assert 0.2 < out[0][1].mean() < 0.9 and (out[0][0][:100, 0] == 3).all()
    for a, b in zip(out[0], out[1]):
        assert np.array_equal(a, b, equal_nan=True)
    bmin = rng.normal(size=(n, 3)).astype(np.float32); bmax = (bmin + rng.random((n, 3)) * 2).astype(np.float32)
    d2 = sphere_dirs(rng, n); d2[:1000, 0] = 0; d2[1000:2000, 1] = -0.0; d2[2000:3000] = [0, 0, 1]
    o2 = ((bmin + bmax) * 0.5 + rng.normal(size=(n, 3)) * 0.7).astype(np.float32)
    res = []
    for fn in (L.ref_aabb_ray, O.orc_aabb_ray_batch):
        hit = np.zeros(n, np.int32); nf = np.zeros((n, 2), np.float32)
        args = (P(bmin), P(bmax), P(o2), P(d2), P(hit), P(nf))
        fn(n, *args) if fn is L.ref_aabb_ray else fn(ctypes.c_uint64(n), *args)
        res.append((hit, nf))
    assert 0.05 < res[0][0].mean() < 0.95
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1], equal_nan=True)


@pytest.mark.skipif(not os.path.exists(REF_GEOM), reason='oracle/_ref/libref_geom.so not built (needs /root/reference)')
def test_oracle_scene_query_pinned_against_reference_text(cp, oracle):
    """The two-level ray query as the reference writes it -- ShapeKDTree::rayIntersect closest / shadow (skdtree.cpp:112-142,207-226: scene box
    clip, adaptive epsilon only when ray.mint == Epsilon, the shadow overload without the inner clamp), HairKDTree::rayIntersect closest /
    visibility (hair.cpp:199-237: per-shape clip) and the kd-tree box enlargement (gkdtree.h:1219-1220) -- cut out of the reference and executed
    as written, with the Havran traversal replaced by a scan over all primitives (same answers up to the order of equal-t hits), against the
    oracle's brute-force query on a four-shape scene: scene bounds, closest shape / segment / t and occlusion are bit-identical."""
    L = ctypes.CDLL(REF_GEOM); L.ref_hair_create.restype = ctypes.c_void_p; L.ref_scene_create.restype = ctypes.c_void_p
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rng = np.random.default_rng(67)
    sc = cp.scenes.SCENES['hair-curl']
    s = oracle.Scene(); b = s.add_bsdf('kajiyakay')
    handles = []; keep = []
    for sh in sc['shapes']:
        xyz, starts = cp.scenes.generate(sh, 0.0015)
        xyz = np.ascontiguousarray(xyz, np.float32); starts = np.ascontiguousarray(starts, np.uint8); keep += [xyz, starts]
        handles.append(L.ref_hair_create(P(xyz), P(starts), len(starts), ctypes.c_float(sh['radius'])))
        s.add_hair(xyz, starts, sh['radius'], b)
    s.set_camera(np.eye(4, dtype=np.float32), 35.0, width=16, height=16); s.build()
    arr = (ctypes.c_void_p * len(handles))(*handles)
    scene = ctypes.c_void_p(L.ref_scene_create(len(handles), arr))
    rb = np.zeros(6, np.float32); L.ref_scene_bounds(scene, P(rb))
    ob, _ = s.scene_bounds()
    assert np.array_equal(rb, ob)
    n = 40000
    allv = np.concatenate(keep[0::2])
    centre = allv.mean(axis=0); rad = float(np.linalg.norm(allv - centre, axis=1).max())
    # a third from outside (mint 0), a third from fiber surfaces with mint == Epsilon (adaptive epsilon), a third with a finite maxt
    tgt = allv[rng.integers(0, len(allv), n)] + (rng.normal(size=(n, 3)) * 0.0006).astype(np.float32)
    d = sphere_dirs(rng, n)
    kind = rng.integers(0, 3, n)
    dist = np.where(kind == 0, 2.5 * rad, rng.random(n) * 0.05 * rad).astype(np.float32)
    o = (tgt - d * dist[:, None]).astype(np.float32)
    mint = np.where(kind == 0, 0.0, np.float32(1e-4)).astype(np.float32)
    maxt = np.where(kind == 2, (dist * (0.5 + rng.random(n))).astype(np.float32), np.inf).astype(np.float32)
    o[:50] = rb[:3]; o[50:100, 1] = rb[4]                                   # origins on the corner / a face of the scene box
    rsh = np.zeros(n, np.int32); riv = np.zeros(n, np.uint32); rt = np.zeros(n, np.float32); rocc = np.zeros(n, np.int32)
    L.ref_scene_intersect(scene, n, P(o), P(d), P(mint), P(maxt), P(rsh), P(riv), P(rt), P(rocc))
    osh, oiv, ot = s.intersect(o, d, mint, maxt, mode=2)
    oany = s.intersect(o, d, mint, maxt, mode=3)[0] >= 0
    assert 0.3 < (rsh >= 0).mean() < 0.99 and len(set(rsh[rsh >= 0].tolist())) == 4
    assert np.array_equal(rocc != 0, oany)
    assert np.array_equal(rsh, osh) and np.array_equal(rt, ot)
    same = riv == oiv
    assert same.mean() > 0.999 and np.array_equal(rt[~same], ot[~same])      # equal t, other segment of a miter joint: visiting order


@pytest.mark.skipif(not os.path.exists(os.path.join(os.path.dirname(GOLDEN), '..', 'oracle', '_ref', 'libref_path.so')), reason='oracle/_ref/libref_path.so not built (needs /root/reference)')
@pytest.mark.parametrize('name,variant', [('straight-hair', {}), ('curly-hair', {}), ('hair-on-head', {}), ('straight-hair-default', {}),
                                          ('straight-hair-thindielectric', dict(maxDepth=24)), ('straight-hair-dielectric', dict(maxDepth=24, hideEmitters=True)),
                                          ('curly-hair', dict(fixed=True)), ('straight-hair', dict(maxDepth=-1, strictNormals=False, hideEmitters=True, rrDepth=2)),
                                          ('furball', dict(maxDepth=2)),
                                          # the `sobol` sampler: path.cpp asks the sampler object itself, in its own order
                                          ('straight-hair', dict(sobol=True)), ('curly-hair', dict(sobol=True, fixed=True, scramble=7)),
                                          ('straight-hair-thindielectric', dict(sobol=True, maxDepth=24)), ('hair-on-head', dict(sobol=True, rrDepth=2, maxDepth=20))])
def test_oracle_li_pinned_against_compiled_reference_integrator(cp, oracle, name, variant):
    """MIPathTracer::Li: src/integrators/path/path.cpp compiled UNMODIFIED from /root/reference (oracle/_ref/libref_path.so) and run on the
    oracle's own scene components through a callback table -- same camera rays, same Philox counters -- against the oracle's Li():
    radiance of every path is bit-identical.  Covers emitter sampling + MIS, the ESmooth test, ENull / delta vertices and the `scattered`
    flag with hideEmitters, strictNormals, maxDepth (finite, 2, infinite), Russian roulette from rrDepth, the extra sampler draws of the
    fixed Marschner, fibers and meshes.  With sobol=True the random numbers come from the oracle's restatement of the `sobol` sampler: the
    compiled path.cpp draws from it as a stateful object (next2D / next1D in whatever order it asks), the oracle's Li() in the order it believes
    the reference asks -- equal radiance pins the consumption order: emitter sample, BSDF sample, the BSDF's own draws, roulette."""
    ov = dict(width=40, height=32, spp=4, maxDepth=variant.get('maxDepth', 8))
    if variant.get('fixed'):
        sh = dict(cp.scenes.SCENES[name]['shapes'][0], bsdf=dict(type='marschner_fixed', id='hair', intIOR=1.55, extIOR=1.0)); ov['shapes'] = [sh]
    env = cp.bake_sunsky(**cp.scenes.sunsky_params(name))
    sc = oracle.scene_from_description(name, scale=0.004, overrides=ov, envmap=env)
    sc.set_integrator(maxDepth=ov['maxDepth'], rrDepth=variant.get('rrDepth', 5), strictNormals=variant.get('strictNormals', True), hideEmitters=variant.get('hideEmitters', False))
    if variant.get('sobol'):
        sc.set_sampler('sobol', scramble=variant.get('scramble', 0))
    ys, xs, ss = np.meshgrid(np.arange(32), np.arange(40), np.arange(4), indexing='ij')
    xy = np.stack([xs.ravel(), ys.ravel()], axis=1).astype(np.uint32); samp = ss.ravel().astype(np.uint32)
    ours, _ = sc.render_samples(xy, samp, 4, seed=21)
    ref, alpha, depth = sc.render_samples_ref_li(xy, samp, 4, seed=21)
    assert np.isfinite(ref).all() and ref.sum() > 0 and depth.max() >= 2
    assert np.array_equal(ours, ref), '%d of %d paths differ' % ((ours != ref).any(axis=1).sum(), len(ref))
    assert (alpha == 1).all()                                                     # films without an alpha channel: EOpacity is masked out
    if variant.get('sobol'):
        sc.set_sampler('philox')
        other, _ = sc.render_samples(xy, samp, 4, seed=21)
        assert not np.array_equal(other, ours)


REF_SOBOL = os.path.join(os.path.dirname(GOLDEN), '..', 'oracle', '_ref', 'libref_sobol.so')


@pytest.mark.skipif(not os.path.exists(REF_SOBOL), reason='oracle/_ref/libref_sobol.so not built (needs /root/reference)')
def test_sobol_sampler_pinned_against_compiled_reference_plugin(cp, oracle):
    """src/samplers/sobol.cpp + sobolseq.cpp compiled UNMODIFIED (oracle/_ref/libref_sobol.so), driven like renderBlock drives a sampler
    (setFilmResolution(size, true), generate(pixel), next2D / next1D, advance), against the oracle's restatement reading the mirrored direction
    numbers: every number bit-identical -- pixel offsets inside [0, 1), power-of-two and odd film sizes, scrambled and not, 300 dimensions deep."""
    R = ctypes.CDLL(REF_SOBOL); R.ref_sobol_create.restype = ctypes.c_void_p
    rng = np.random.default_rng(73)
    for (w, h, spp, scramble) in ((1024, 1024, 64, 0), (1280, 720, 64, 0), (96, 54, 8, 0), (512, 512, 16, 12345), (3, 2, 4, 0), (2048, 2048, 256, 2 ** 40 + 17)):
        hdl = ctypes.c_void_p(R.ref_sobol_create(spp, ctypes.c_ulonglong(scramble), w, h)); assert hdl.value
        sc = oracle.Scene(); b = sc.add_bsdf('diffuse', reflectance=0.5); sc.add_rectangle(None, False, b)
        sc.set_camera(np.eye(4, dtype=np.float32), width=w, height=h); sc.set_sampler('sobol', scramble=scramble)
        pattern = np.array([2] + [2, 2, 1] * 60, np.int32)                      # pixel sample, then (emitter, bsdf, roulette) per vertex
        n = int(pattern.sum())
        for _ in range(12):
            px, py = int(rng.integers(0, w)), int(rng.integers(0, h)); first = int(rng.integers(0, spp)); cnt = min(3, spp - first)
            ref = np.zeros(cnt * n, np.float32)
            assert R.ref_sobol_sequence(hdl, px, py, first, cnt, len(pattern), pattern.ctypes.data_as(ctypes.c_void_p), ref.ctypes.data_as(ctypes.c_void_p)) == 0
            ours = sc.sobol_sequence(px, py, first, cnt, pattern)
            assert np.array_equal(ours.ravel(), ref)
            assert (ours[:, :2] >= 0).all() and (ours[:, :2] < 1).all() and (ours >= 0).all() and (ours < 1).all()
    # beyond 1024 dimensions the plugin raises (sobol.cpp:222-224): so does the restatement
    sc = oracle.Scene(); b = sc.add_bsdf('diffuse', reflectance=0.5); sc.add_rectangle(None, False, b)
    sc.set_camera(np.eye(4, dtype=np.float32), width=64, height=64); sc.set_sampler('sobol')
    with pytest.raises(RuntimeError, match='direction number table'):
        sc.sobol_sequence(3, 4, 0, 1, np.full(600, 2, np.int32))


@pytest.mark.skipif(not os.path.exists(REF_GEOM), reason='oracle/_ref/libref_geom.so not built (needs /root/reference)')
@pytest.mark.parametrize('which', ['sunsky', 'hdr-rotated'])
def test_oracle_envmap_pinned_against_reference_text(cp, oracle, which):
    """The environment-map emitter -- EnvironmentMap::configure (CDFs, row weights, normalisation), evalEnvironment, sampleDirect with
    internalSampleDirection / sampleReuse, pdfDirect with internalPdfDirection, fillDirectSamplingRecord (src/emitters/envmap.cpp), MIPMap::evalTexel /
    evalBilinear (mipmap.h), BSphere::rayIntersect, solveQuadratic and squareToTent -- cut out of the reference and executed as written over
    IEEE half texels (oracle/ref_shim/ref_env.cpp), against the oracle: column CDFs bit-identical, row tables to one ulp; eval / pdf / sampled direction / value to the last
    bits (the reference calls libm's sinf / sincosf / atan2f / acosf, the oracle rounds correctly)."""
    L = ctypes.CDLL(REF_GEOM); L.ref_env_create.restype = ctypes.c_void_p
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rng = np.random.default_rng(71)
    if which == 'sunsky':
        img = cp.bake_sunsky(**cp.scenes.sunsky_params('curly-hair')); tw = np.eye(4, dtype=np.float32); scale = 1.0
    else:
        yy, xx = np.mgrid[0:48, 0:96]                                       # smooth sky + a bright lobe, black rows below the horizon
        img = np.stack([2 + np.sin(xx / 9.0) * np.cos(yy / 7.0), 1.5 + np.cos(xx / 11.0), 1 + 0.5 * np.sin(yy / 5.0)], axis=2) * 3.0
        img = (img + 400.0 * np.exp(-((xx - 62) ** 2 + (yy - 7) ** 2) / 6.0)[..., None] * [1.0, 0.9, 0.6]).astype(np.float32); img[40:] = 0
        tw = np.array([[0, 0, 2, 0], [0, 2, 0, 0], [-2, 0, 0, 0], [0, 0, 0, 1]], np.float32); scale = 2.5                          # quarter turn, scale 2: exact inverse
    tl = np.linalg.inv(tw.astype(np.float64)).astype(np.float32)
    assert np.array_equal((tw.astype(np.float64) @ tl.astype(np.float64)), np.eye(4))
    s = oracle.scene_from_description('curly-hair', scale=0.002, overrides=dict(width=16, height=16, spp=1, maxDepth=2), envmap=img)
    s.set_envmap(img, toWorld=tw, scale=scale); s.build()
    aabb, bs = s.scene_bounds()
    img = np.ascontiguousarray(img, np.float32); h, w = img.shape[:2]
    e = ctypes.c_void_p(L.ref_env_create(P(img), w, h, P(np.ascontiguousarray(tw)), P(np.ascontiguousarray(tl)), ctypes.c_float(scale), P(np.ascontiguousarray(bs[:3], np.float32)), ctypes.c_float(bs[3])))
    rows = np.zeros(h + 1, np.float32); cols = np.zeros((h, w + 1), np.float32); rw = np.zeros(h, np.float32); nrm = ctypes.c_float(0)
    L.ref_env_tables(e, P(rows), P(cols), P(rw), ctypes.byref(nrm))
    orows, ocols, orw, onrm = s.env_tables()
    # column CDFs and the normalisation are bit-identical; the row weights are sin((y + 0.5) pi / h) -- libm's sinf in the reference, correctly
    # rounded in the oracle: a handful differ by one ulp, which moves the same number of row-CDF entries by an ulp
    assert np.array_equal(cols, ocols, equal_nan=True) and np.isclose(nrm.value, onrm, rtol=2e-7, atol=0)
    assert np.abs(rw - orw).max() <= 6e-8 and (rw != orw).mean() < 0.1 and np.abs(rows - orows).max() <= 2e-7 and (rows != orows).mean() < 0.1
    n = 200000
    d = sphere_dirs(rng, n); d[:10] = [0, 1, 0]; d[10:20] = [0, -1, 0]; d[20:30] = [0, 0, 1]
    rgb = np.zeros((n, 3), np.float32); pdf = np.zeros(n, np.float32)
    L.ref_env_eval(e, n, P(d), P(rgb), P(pdf))
    orgb, opdf = s.env_eval(d)
    assert np.array_equal(rgb != 0, orgb != 0) and np.array_equal(pdf != 0, opdf != 0)
    # one ulp of atan2f / acosf in the texture coordinate, times 512 texels, times the gradient at the rim of the sun disc: a few 1e-5 at worst
    tol = lambda a, b: np.abs(a - b) <= 2e-4 * np.maximum(np.abs(b), 2e-2 * np.abs(b).max())
    assert tol(rgb, orgb).all() and tol(pdf, opdf).all()
    assert (rgb == orgb).mean() > 0.8 and (pdf == opdf).mean() > 0.8
    ref = (np.asarray(bs[:3]) + rng.normal(size=(n, 3)) * 0.3 * bs[3] / 1.5).astype(np.float32)
    smp = rng.random((n, 2)).astype(np.float32)
    sd = np.zeros((n, 3), np.float32); sv = np.zeros((n, 3), np.float32); spd = np.zeros((n, 2), np.float32)
    L.ref_env_sample(e, n, P(ref), P(smp), P(sd), P(sv), P(spd))
    od, ov, opdf2, odist = s.env_sample(ref, smp)
    live = spd[:, 0] != 0
    assert np.array_equal(live, opdf2 != 0) and live.mean() > 0.9
    assert np.abs(sd[live] - od[live]).max() < 2e-5 and tol(sv[live], ov[live]).all() and tol(spd[live, 0], opdf2[live]).all()
    assert np.abs(spd[live, 1] - odist[live]).max() <= 1e-5 * bs[3]
    o = ref; ok = np.zeros(n, np.int32)
    L.ref_env_fill(e, n, P(o), P(d), P(ok))
    assert ok.mean() > 0.99                                                   # reference points inside the scene sphere: the record can be filled


@pytest.mark.skipif(not os.path.exists(REF_GEOM), reason='oracle/_ref/libref_geom.so not built (needs /root/reference)')
@pytest.mark.parametrize('which', ['sunsky', 'odd-size'])
def test_envmap_mip_pyramid_and_ewa_pinned_against_reference_text(cp, oracle, which):
    """The filtered environment lookup of camera rays that leave the scene: the reference's Resampler<float> (rfilter.h:107-460),
    LanczosSincFilter::eval (lanczos.cpp:42-55), MIPMap::eval and evalEWA (mipmap.h:629-725, 760-836), hypot2 / log2 (math.cpp) cut out and
    executed as written, driven like TMIPMap's constructor drives them (oracle/ref_shim/ref_env.cpp) -- against the oracle AND the product's
    host-side pyramid builder (csrc/cp_host_mip.cpp).  Pyramid: same level sizes; texels identical except where libm's sinf (reference) and the
    correctly rounded sine (oracle, product) differ in the last bit of a filter weight, which can move a value across a half rounding
    boundary (one half ulp).  Lookups: footprints from a hundredth of a texel to a quarter of the map, isotropic and 40:1 anisotropic."""
    L = ctypes.CDLL(REF_GEOM); L.ref_env_create.restype = ctypes.c_void_p
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rng = np.random.default_rng(5)
    if which == 'sunsky':
        img = cp.bake_sunsky(**cp.scenes.sunsky_params('hair-curl'))
    else:                                                                   # odd sizes exercise max(1, (size + 1) / 2) and the 1-wide tail of the pyramid
        yy, xx = np.mgrid[0:37, 0:101]
        img = np.stack([2 + np.sin(xx / 5.0) * np.cos(yy / 4.0), 1.5 + np.cos(xx / 7.0), 1 + 0.5 * np.sin(yy / 3.0)], axis=2)
        img = (img + 300.0 * np.exp(-((xx - 70) ** 2 + (yy - 9) ** 2) / 3.0)[..., None] * [1.0, 0.9, 0.6]).astype(np.float32); img[30:] = 0; img[3, 5] = -1.0
    img = np.ascontiguousarray(img, np.float32); h, w = img.shape[:2]
    tw = np.eye(4, dtype=np.float32)
    s = oracle.scene_from_description('curly-hair', scale=0.002, overrides=dict(width=16, height=16, spp=1, maxDepth=2), envmap=img)
    aabb, bs = s.scene_bounds()
    e = ctypes.c_void_p(L.ref_env_create(P(img), w, h, P(tw), P(tw), ctypes.c_float(1.0), P(np.ascontiguousarray(bs[:3], np.float32)), ctypes.c_float(bs[3])))
    olev = s.env_mip_levels()
    plev = cp.env_pyramid(img)
    lw = ctypes.c_int(); lh = ctypes.c_int()
    nlev = L.ref_env_mip_level(e, 0, ctypes.byref(lw), ctypes.byref(lh), None)
    assert nlev == len(olev) == len(plev) and olev[-1].shape[:2] == (1, 1)
    for l in range(nlev):
        L.ref_env_mip_level(e, l, ctypes.byref(lw), ctypes.byref(lh), None)
        ref = np.zeros((lh.value, lw.value, 3), np.float32)
        L.ref_env_mip_level(e, l, ctypes.byref(lw), ctypes.byref(lh), P(ref))
        assert ref.shape == olev[l].shape == plev[l].shape and (ref >= 0).all()
        prod = plev[l].astype(np.float16).astype(np.float32)               # the device quantises the product's fp32 levels to half (round to nearest even)
        assert np.array_equal(prod, olev[l]), 'level %d: product and oracle pyramids differ' % l
        half_ulp = np.maximum(np.abs(ref), 6.1e-5) * 2.0 ** -10
        assert (np.abs(ref - olev[l]) <= half_ulp).all() and (ref == olev[l]).mean() > 0.97, 'level %d' % l
    n = 60000
    d = sphere_dirs(rng, n)
    foot = np.exp(rng.uniform(np.log(1e-5), np.log(0.4), size=(n, 1))).astype(np.float32)                # angular footprint (radians), log-uniform
    aniso = np.where(rng.random((n, 1)) < 0.5, 1.0, np.exp(rng.uniform(0, np.log(40.0), size=(n, 1)))).astype(np.float32)
    t1 = np.cross(d, sphere_dirs(rng, n)); t1 /= np.linalg.norm(t1, axis=1, keepdims=True); t2 = np.cross(d, t1)
    rx = (d + t1 * foot).astype(np.float32); ry = (d + t2 * foot / aniso).astype(np.float32)
    d[:4] = [[0, 1, 0], [0, -1, 0], [0, 0, 1], [1, 0, 0]]; rx[4] = d[4]; ry[5] = d[5]                       # poles, the seam, degenerate footprints
    ref = np.zeros((n, 3), np.float32)
    L.ref_env_eval_filtered(e, n, P(d), P(rx), P(ry), P(ref))
    ours = s.env_eval_filtered(d, rx, ry)
    assert np.isfinite(ref).all() and np.array_equal(np.isfinite(ours), np.isfinite(ref))
    scale = float(np.abs(ref).max())
    err = np.abs(ours - ref) / np.maximum(np.abs(ref), 1e-3 * scale)
    print('%s: %d levels; filtered lookups: %.4f bit-identical, rel err q99 %.2e max %.2e' % (which, nlev, (ours == ref).all(axis=1).mean(), np.quantile(err, 0.99), err.max()))
    # the last bit of atan2f / acosf / sincosf in a texture coordinate, and the half-ulp texel differences above, scaled by the texel gradient
    assert np.quantile(err, 0.99) <= 2e-4 and err.max() <= 5e-3
    assert (ours == ref).all(axis=1).mean() > 0.5


@pytest.mark.skipif(not os.path.exists(REF_GEOM), reason='oracle/_ref/libref_geom.so not built (needs /root/reference)')
@pytest.mark.parametrize('rfilter,param', [('tent', 0.0), ('box', 0.0), ('gaussian', 0.0), ('gaussian', 0.8)])
def test_film_splat_pinned_against_reference(cp, oracle, rfilter, param):
    """ImageBlock::put (imageblock.h:124-186) cut out of the reference and executed as written, over the reference's own filter files compiled
    unmodified (src/libcore/rfilter.cpp: 31-tap discretisation; src/rfilters/{tent,box,gaussian}.cpp), against the oracle's film: filter tables
    and the accumulated film (R, G, B, alpha, weight) are bit-identical, invalid samples are refused by both."""
    L = ctypes.CDLL(REF_GEOM); L.ref_filter_create.restype = ctypes.c_void_p
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    tab = np.zeros(32, np.float32); info = np.zeros(2, np.float32)
    f = ctypes.c_void_p(L.ref_filter_create({'tent': 0, 'box': 1, 'gaussian': 2}[rfilter], ctypes.c_float(param), P(tab), P(info)))
    s = oracle.Scene(); b = s.add_bsdf('kajiyakay')
    s.add_hair(np.array([[0, 0, 0], [0, 1, 0], [0.1, 2, 0]], np.float32), np.array([1, 0, 0], np.uint8), 0.05, b)
    W, H = 37, 23
    s.set_camera(np.eye(4, dtype=np.float32), 35.0, width=W, height=H); s.set_film(rfilter, param); s.build()
    otab = s.filter_table()
    if rfilter == 'gaussian':                                      # expf of libm vs correctly rounded
        assert np.abs(tab - otab).max() <= 2e-7 * otab.max() and (tab == otab).mean() > 0.5
    else:
        assert np.array_equal(tab, otab)
    rng = np.random.default_rng(73)
    n = 60000
    pos = (rng.random((n, 2)) * [W + 4, H + 4] - 2).astype(np.float32)          # also outside the film: clipped taps
    rgb = (rng.random((n, 3)) ** 3 * 10).astype(np.float32); alpha = rng.random(n).astype(np.float32)
    rgb[:20, 0] = np.nan; rgb[20:40, 1] = -1.0; rgb[40:60, 2] = np.inf; alpha[60:80] = -0.5
    film = np.zeros((H, W, 5), np.float32); ok = np.zeros(n, np.int32)
    L.ref_film_put(f, W, H, n, P(pos), P(rgb), P(alpha), P(film), P(ok))
    ofilm = s.splat(pos, rgb, alpha)
    assert ok[:80].sum() == 0 and ok[80:].all()
    if rfilter == 'gaussian':
        assert np.allclose(film, ofilm, rtol=1e-6, atol=1e-6)
    else:
        assert np.array_equal(film, ofilm)
    assert film[..., 4].sum() > 0


@pytest.mark.skipif(not os.path.exists(REF_GEOM), reason='oracle/_ref/libref_geom.so not built (needs /root/reference)')
def test_oracle_camera_pinned_against_reference_text(cp, oracle):
    """PerspectiveCameraImpl::configure / sampleRayDifferential (src/sensors/perspective.cpp:126-180,271-298) with the reference's own fp32
    Transform algebra -- Transform::operator*, translate, scale, perspective (transform.cpp), the matrix product (matrix.h:743-757) and
    Matrix::invert (matrix.inl:138-193: Gauss-Jordan in fp32) -- cut out of the reference and executed as written, against the oracle:
    sampleToCamera, the near-plane differentials, ray origins, directions, intervals and differential directions are bit-identical for the
    cameras of all four scene files; Matrix::invert alone is bit-identical on random and on scene matrices."""
    L = ctypes.CDLL(REF_GEOM); L.ref_camera_create.restype = ctypes.c_void_p
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rng = np.random.default_rng(79)
    for name, (w, h) in (('straight-hair', (512, 512)), ('hair-curl', (1200, 1000)), ('furball', (2048, 2048)), ('curly-hair', (333, 777))):
        sc = cp.scenes.SCENES[name]
        tw = np.ascontiguousarray(np.array(sc['camera'], np.float32).reshape(4, 4))
        for fov, near, far in ((sc['fov'], 1e-2, 1e4), (61.3, 0.5, 250.0)):
            cam = ctypes.c_void_p(L.ref_camera_create(P(tw), ctypes.c_float(fov), ctypes.c_float(near), ctypes.c_float(far), w, h))
            s2c = np.zeros(16, np.float32); dxdy = np.zeros(6, np.float32)
            L.ref_camera_matrices(cam, P(s2c), P(dxdy))
            s = oracle.Scene(); b = s.add_bsdf('kajiyakay')
            s.add_hair(np.array([[0, 0, 0], [0, 1, 0], [0.1, 2, 0]], np.float32), np.array([1, 0, 0], np.uint8), 0.05, b)
            s.set_camera(tw, fov, nearClip=near, farClip=far, width=w, height=h); s.build()
            n = 50000
            pxy = (rng.random((n, 2)) * [w, h]).astype(np.float32); pxy[:4] = [[0, 0], [w, h], [w, 0], [0.5, h - 0.5]]
            ro = np.zeros((n, 3), np.float32); rd = np.zeros((n, 3), np.float32); rmm = np.zeros((n, 2), np.float32); rxy = np.zeros((n, 6), np.float32)
            L.ref_camera_rays(cam, n, P(pxy), P(ro), P(rd), P(rmm), P(rxy))
            oo, od, omin, omax = s.camera_rays(pxy)
            assert np.array_equal(ro, oo) and np.array_equal(rd, od) and np.array_equal(rmm[:, 0], omin) and np.array_equal(rmm[:, 1], omax)
            # the differential directions reach the film through evalEnvironment only; compare them through a full sample below (Li test) and here via dx/dy
            odx = s.camera_differentials() if hasattr(s, 'camera_differentials') else None
            if odx is not None:
                assert np.array_equal(dxdy, odx)
    for k in range(200):
        m = rng.normal(size=(4, 4)).astype(np.float32)
        if k % 3 == 0:
            m[3] = [0, 0, 0, 1]
        if k == 0:
            m = np.ascontiguousarray(np.array(cp.scenes.SCENES['hair-curl']['camera'], np.float32).reshape(4, 4))
        a = np.zeros(16, np.float32); ok = L.ref_matrix_invert(P(np.ascontiguousarray(m)), P(a))
        bmat = oracle.matrix_invert(m)
        assert ok == 1 and np.array_equal(a.reshape(4, 4), bmat)
    sing = np.zeros(16, np.float32)
    assert L.ref_matrix_invert(P(np.zeros((4, 4), np.float32)), P(sing)) == 0 and oracle.matrix_invert(np.zeros((4, 4), np.float32)) is None


# ------------------------------------------------------------------------------------------------ golden vectors
def test_validate_scene_xml_dry_run(cp, tmp_path):
    """cudapath_validate_scene_xml: the scene loader without a GPU -- lists what a file would create, names what is unsupported."""
    for name in ('hair-curl', 'straight-hair-default', 'hair-on-head', 'straight-hair-dielectric', 'straight-hair-thindielectric', 'straight-hair'):
        d = tmp_path / name
        path = cp.scenes.write_scene(name, str(d), scale=0.002)
        rep = cp.validate_scene_xml(path)
        sc = cp.scenes.SCENES[name]
        assert sum(r.startswith('bsdf') for r in rep) == len(sc['shapes']) and sum(r.startswith('shape') for r in rep) == len(sc['shapes'])
        assert not any('missing' in r for r in rep) and rep[-1] == 'sampleCount %d' % sc['spp']
        assert any(r.startswith('integrator path') for r in rep) and any(r.startswith('emitter sunsky') for r in rep)
    xml = open(path).read()
    bad = tmp_path / 'bad.xml'
    bad.write_text(xml.replace('type="kajiyakay"', 'type="roughdielectric"'))
    with pytest.raises(cp.CudapathError, match='roughdielectric'):
        cp.validate_scene_xml(str(bad))
    bad.write_text(xml.replace('models/hair.mitshair', 'models/nothing.mitshair'))
    assert any('file missing' in r for r in cp.validate_scene_xml(str(bad)))
    bad.write_text(xml.replace('<scene version="0.6.0">', '<scene>'))
    with pytest.raises(cp.CudapathError, match='version'):
        cp.validate_scene_xml(str(bad))
    bad.write_text(xml.replace('value="8"', 'value="$depth"', 1))
    with pytest.raises(cp.CudapathError, match='undefined'):
        cp.validate_scene_xml(str(bad))
    assert cp.validate_scene_xml(str(bad), defines={'depth': 8})


def test_oracle_matches_committed_golden_vectors(oracle):
    g = np.load(os.path.join(GOLDEN, 'bsdf_golden.npz'))
    s = make_bsdf_scene(oracle)
    for b in range(3):
        ev, pdf = s.bsdf_eval(b, g['wi'], g['wo'])
        assert np.array_equal(ev, g['eval_%d' % b]) and np.array_equal(pdf, g['pdf_%d' % b])
        wo, wt, p, ty = s.bsdf_sample(b, g['wi'], g['sample'])
        assert np.array_equal(wo, g['swo_%d' % b]) and np.array_equal(wt, g['swt_%d' % b]) and np.array_equal(ty, g['sty_%d' % b])


def test_oracle_render_matches_golden_film(cp, oracle):
    if not oracle.have_ref():
        pytest.skip('oracle/_ref not built')
    g = np.load(os.path.join(GOLDEN, 'render_golden.npz'))
    ov = dict(width=32, height=24, spp=4, maxDepth=6)
    env = cp.bake_sunsky(**cp.scenes.sunsky_params('curly-hair'))
    film = oracle.scene_from_description('curly-hair', scale=0.004, overrides=ov, envmap=env).render(4, seed=5, threads=2)
    assert np.allclose(film, g['film'], rtol=1e-5, atol=1e-6)


SECOND_SET_MATS = [('roughplastic', dict(intIOR=1.55, extIOR=1.0, alpha=0.2, distribution='ggx', diffuseReflectance=(0.143016, 0.0156076, 1.80928e-005))),
                   ('roughplastic', dict(intIOR=1.49, extIOR=1.000277, alpha=0.1, distribution='beckmann', nonlinear=True, diffuseReflectance=(0.6, 0.5, 0.4))),
                   ('roughplastic', dict(intIOR=1.55, extIOR=1.0, alpha=0.3, distribution='phong', diffuseReflectance=(0.2, 0.3, 0.4))),
                   ('marschner_fixed', dict(intIOR=1.55, extIOR=1.000277)),
                   ('diffuse', dict(reflectance=(0.5, 0.4, 0.3))), ('twosided', dict(reflectance=(0.5, 0.4, 0.3)))]


def test_oracle_matches_second_golden_set(cp, oracle):
    """bsdf2 / mesh / render_mesh fixtures (tests/golden/make_golden.py --second): roughplastic, fixed Marschner, diffuse, fibers + meshes."""
    g = np.load(os.path.join(GOLDEN, 'bsdf2_golden.npz'))
    s = oracle.Scene()
    for t, p in SECOND_SET_MATS:
        s.add_bsdf(t, **p)
    for b in range(len(SECOND_SET_MATS)):
        ev, pdf = s.bsdf_eval(b, g['wi'], g['wo'])
        assert np.array_equal(ev, g['eval_%d' % b], equal_nan=True) and np.array_equal(pdf, g['pdf_%d' % b], equal_nan=True)
        wo, wt, p, ty = s.bsdf_sample(b, g['wi'], g['sample'], g['extra'])
        assert np.array_equal(wo, g['swo_%d' % b], equal_nan=True) and np.array_equal(wt, g['swt_%d' % b], equal_nan=True) and np.array_equal(ty, g['sty_%d' % b])
    if not oracle.have_ref():
        pytest.skip('oracle/_ref not built')
    m = np.load(os.path.join(GOLDEN, 'mesh_golden.npz'))
    ov = dict(width=32, height=24, spp=4, maxDepth=6)
    env = cp.bake_sunsky(**cp.scenes.sunsky_params('hair-on-head'))
    sc = oracle.scene_from_description('hair-on-head', scale=0.004, overrides=ov, envmap=env)
    sh, pr, t = sc.intersect(m['o'], m['d'], 0.0, np.inf, mode=2)
    assert np.array_equal(sh, m['shape']) and np.array_equal(pr, m['prim']) and np.array_equal(t, m['t'])
    assert (sh == 0).sum() > 50 and (sh == 1).sum() > 1000 and (sh == 2).sum() > 100
    film = sc.render(4, seed=9, threads=2)
    assert np.allclose(film, np.load(os.path.join(GOLDEN, 'render_mesh_golden.npz'))['film'], rtol=1e-5, atol=1e-6)


THIRD_SET_MATS = [('thindielectric', dict(intIOR=1.55, extIOR=1.0, specularReflectance=HAIR_RGB, specularTransmittance=HAIR_RGB)),
                  ('thindielectric', dict(intIOR=1.5046, extIOR=1.000277, specularReflectance=(0.9, 0.5, 0.1), specularTransmittance=(2.0, 1.0, 0.5))),
                  ('marschnerdielectric', dict(intIOR=1.55, extIOR=1.0, exponent=5.0, specularTransmittance=HAIR_RGB, specularReflectance=HAIR_RGB, diffuseReflectance=HAIR_RGB)),
                  ('marschnerdielectric', dict(intIOR=1.501, extIOR=1.000277, diffuseReflectance=(0.3, 0.2, 0.1), specularReflectance=(0.4, 0.3, 0.2),
                                               specularTransmittance=(1.5, 0.6, 0.7)))]


def test_oracle_matches_third_golden_set(cp, oracle):
    """bsdf3 / render_dielectric fixtures (tests/golden/make_golden.py --third): thindielectric and marschnerdielectric."""
    g = np.load(os.path.join(GOLDEN, 'bsdf3_golden.npz'))
    s = oracle.Scene()
    for t, p in THIRD_SET_MATS:
        s.add_bsdf(t, **p)
    for b in range(len(THIRD_SET_MATS)):
        wo, wt, p, ty = s.bsdf_sample(b, g['wi'], g['sample'])
        assert np.array_equal(wo, g['swo_%d' % b]) and np.array_equal(wt, g['swt_%d' % b]) and np.array_equal(p, g['spdf_%d' % b]) and np.array_equal(ty, g['sty_%d' % b])
        for discrete in (0, 1):
            ev, pdf = s.bsdf_eval(b, g['wi'], wo, discrete=bool(discrete))
            assert np.array_equal(ev, g['eval_%d_%d' % (b, discrete)]) and np.array_equal(pdf, g['pdf_%d_%d' % (b, discrete)])
    if not oracle.have_ref():
        pytest.skip('oracle/_ref not built')
    films = np.load(os.path.join(GOLDEN, 'render_dielectric_golden.npz'))
    for name in ('straight-hair-thindielectric', 'straight-hair-dielectric'):
        ov = dict(width=32, height=24, spp=4, maxDepth=24)
        env = cp.bake_sunsky(**cp.scenes.sunsky_params(name))
        film = oracle.scene_from_description(name, scale=0.004, overrides=ov, envmap=env).render(4, seed=13, threads=2)
        assert np.allclose(film, films[name.replace('-', '_')], rtol=1e-5, atol=1e-6)


def test_fast_accelerator_matches_checker(cp, oracle):
    """The CPU-baseline build of the oracle (liboracle_fast.so: reference optimisation flags, libm, and the ray-query accelerator of
    oracle/o_hair.h under ORC_FAST -- 8-wide BVH over pre-split segment references, fp32 pre-test, mailbox) answers ray queries like
    the checker build: same hit / miss, shape and primitive, and the same fp32 distance for fibers (the exact FP64 test is the shared,
    unchanged function); triangles agree to an ulp (their fp32 arithmetic is compiled with -funsafe-math-optimizations there).
    The baseline that bench.py times is therefore the same algorithm as the checker, only faster."""
    import importlib.util
    fast_path = os.path.join(os.path.dirname(oracle.ORACLE_LIB), 'liboracle_fast.so')
    if not os.path.exists(fast_path):
        pytest.skip('liboracle_fast.so not built')
    spec = importlib.util.spec_from_file_location('orc_fast', oracle.__file__)
    fast = importlib.util.module_from_spec(spec); spec.loader.exec_module(fast)
    fast.ORACLE_LIB = fast_path
    assert 'pre-split' in fast.accel_description() and 'checker' in oracle.accel_description()
    env = np.ones((16, 32, 3), np.float32)
    rng = np.random.default_rng(5)
    for name, scale in (('hair-curl', 0.02), ('furball', 0.02), ('straight-hair', 0.02), ('hair-on-head', 0.004)):
        a = oracle.scene_from_description(name, scale=scale, envmap=env)
        b = fast.scene_from_description(name, scale=scale, envmap=env)
        aabb, bs = a.scene_bounds()
        n = 60000
        c = 0.5 * (aabb[:3] + aabb[3:]); r = 0.5 * np.linalg.norm(aabb[3:] - aabb[:3])
        p1 = c + r * sphere_dirs(rng, n); p2 = c + r * sphere_dirs(rng, n)
        d = p2 - p1; d /= np.linalg.norm(d, axis=1, keepdims=True)
        o = p1.astype(np.float32); d = d.astype(np.float32)
        sa, pa, ta = a.intersect(o, d, 0.0, np.inf)
        sb, pb, tb = b.intersect(o, d, 0.0, np.inf)
        hit = sa >= 0
        assert hit.sum() > 500 and np.array_equal(sa, sb)
        # secondary-like rays from the hit points (mint = Epsilon: adaptive epsilon, rays starting on / inside fibers), closest and any-hit
        hp = (o[hit] + d[hit] * ta[hit][:, None]).astype(np.float32); d2 = sphere_dirs(rng, int(hit.sum()))
        s2a, p2a, t2a = a.intersect(hp, d2, 1e-4, np.inf); s2b, p2b, t2b = b.intersect(hp, d2, 1e-4, np.inf)
        oa, _, _ = a.intersect(hp, d2, 1e-4, 5.0, mode=1); ob, _, _ = b.intersect(hp, d2, 1e-4, 5.0, mode=1)
        assert np.array_equal(oa >= 0, ob >= 0)
        for (s1, q1, t1, s2, q2, t2) in ((sa, pa, ta, sb, pb, tb), (s2a, p2a, t2a, s2b, p2b, t2b)):
            same = (s1 == s2) & (q1 == q2)
            ties = ~same & (np.abs(t1 - t2) <= 1e-6 * np.maximum(np.abs(t1), 1.0))          # equal-t hits resolve by visiting order
            assert (same | ties).all(), '%s: %d rays answer differently' % (name, int((~(same | ties)).sum()))
            assert ties.sum() <= 1e-3 * len(s1)
            h = (s1 >= 0) & same
            if name == 'hair-on-head':
                assert np.allclose(t1[h], t2[h], rtol=5e-5, atol=2e-4)       # Wald test: cancellation in (n_d - o.n) at |o| ~ 20
            else:
                assert np.array_equal(t1[h], t2[h])


REF_SUN = os.path.join(os.path.dirname(GOLDEN), '..', 'oracle', '_ref', 'libref_sun.so')


@pytest.mark.skipif(not os.path.exists(REF_SUN), reason='oracle/_ref/libref_sun.so not built (needs /root/reference)')
def test_sun_radiance_pinned_against_reference_text(cp, oracle):
    """SURVEY E1, the spectral side of the sunsky bake: computeSunRadiance with its absorption tables (src/emitters/sunsky/sunmodel.h:252-371),
    Spectrum::fromContinuousSpectrum / fromXYZ, ProductSpectrum, InterpolatedSpectrum::eval and ContinuousSpectrum::average
    (src/libcore/spectrum.cpp) and the adaptive Gauss-Lobatto rule behind it (src/libcore/quad.cpp:287-420) cut out of the reference and
    compiled as written (oracle/ref_shim/ref_sunrad.cpp), against the product (cudapath_sun_radiance) and the oracle.  Both integrate the
    piecewise-linear products exactly (Simpson per linear piece) where the reference runs its adaptive rule to a 1e-4 tolerance: the three
    agree to 1e-4 (measured 4e-5) of the largest channel from the zenith to the horizon and from clear to hazy air."""
    R = ctypes.CDLL(REF_SUN); L = oracle.lib()
    worst_p = worst_o = 0.0
    for turb in (2.0, 3.0, 4.5, 6.0, 10.0):
        for d in ((0, 1, 0), (0.3, 0.9, 0.2), (-0.5, 0.3, 0.7), (0.9, 0.05, 0.1), (0.2, 0.5, -0.8), (-0.7, 0.02, -0.1),
                  (-0.376047, 0.758426, 0.532333)):                                                              # the last one: models/*/scene.xml
            d = np.array(d, np.float32); d /= np.linalg.norm(d)
            theta = np.float32(np.arccos(np.float32(d[1])))
            ref = np.zeros(3, np.float32); orc_rgb = np.zeros(3, np.float32)
            R.ref_sun_radiance(ctypes.c_float(theta), ctypes.c_float(turb), ref.ctypes.data_as(ctypes.c_void_p))
            assert L.orc_sun_radiance(oracle.REF_LIB.encode(), ctypes.c_float(theta), ctypes.c_float(turb), orc_rgb.ctypes.data_as(ctypes.c_void_p)) == 0
            prod = cp.sun_radiance(turb, d)
            assert (ref >= 0).all() and ref[0] > 0 and np.array_equal(ref == 0, prod == 0) and np.array_equal(ref == 0, orc_rgb == 0)
            worst_p = max(worst_p, float(np.abs(prod - ref).max() / ref.max())); worst_o = max(worst_o, float(np.abs(orc_rgb - ref).max() / ref.max()))
    print('sun radiance vs the compiled reference: product %.2e, oracle %.2e of the largest channel' % (worst_p, worst_o))
    assert worst_p <= 1e-4 and worst_o <= 1e-4            # the tolerance the reference itself asks of its integrator (GaussLobattoIntegrator(10000, Epsilon, Epsilon), spectrum.cpp:547)


@needs_ref_bsdf
def test_fresnel_diffuse_reflectance_pinned_against_reference_text(oracle):
    """fresnelDiffuseReflectance(eta, fast = false) (src/libcore/util.cpp:807-862) over the adaptive Gauss-Lobatto rule (src/libcore/quad.cpp:287-420),
    both cut out of the reference and compiled as written, against the oracle's restatement (o_math.h): the same bits from eta = 0.4 to 3."""
    R = ctypes.CDLL(REF_BSDF); R.ref_fresnel_diffuse_reflectance.restype = ctypes.c_float
    L = oracle.lib()
    for eta in np.concatenate([np.linspace(0.4, 0.98, 30), np.linspace(1.02, 3.0, 60), [1.49 / 1.000277, 1.000277 / 1.49, 1.5, 1 / 1.5]]).astype(np.float32):
        a = R.ref_fresnel_diffuse_reflectance(ctypes.c_float(eta)); b = L.orc_fresnel_diffuse_reflectance(ctypes.c_float(eta))
        assert np.float32(a) == np.float32(b) and 0 < a < 1, (eta, a, b)


@needs_ref_bsdf
@pytest.mark.parametrize('plugin,props', [
    ('plastic', dict(intIOR=1.5, extIOR=1.0, nonlinear=True, diffuseReflectance=(0.9, 0.9, 0.9))),        # models/teapot/scene.xml:31-38
    ('plastic', dict(diffuseReflectance=(0.2, 0.5, 0.7), specularReflectance=(0.9, 0.8, 1.3))),
    ('twosided:plastic', dict(intIOR=1.5, extIOR=1.0, nonlinear=True, diffuseReflectance=(0.9, 0.9, 0.9))),
    ('twosided:roughplastic', dict(intIOR=1.55, extIOR=1.0, alpha=0.2, distribution='ggx', diffuseReflectance=(0.4, 0.3, 0.2))),
    ('mirror', dict(specularReflectance=(0.9, 0.9, 0.9))), ('twosided:mirror', dict(specularReflectance=(1.2, 0.9, 0.6)))])          # models/teapot/mirror_scene.xml:32-36
def test_oracle_plastic_and_twosided_pinned_against_compiled_reference_plugins(oracle, plugin, props):
    """src/bsdfs/plastic.cpp compiled unmodified (with fresnelDiffuseReflectance and the Gauss-Lobatto rule cut out of libcore), alone and inside
    src/bsdfs/twosided.cpp, against the oracle: both measures (the delta reflection lives in the discrete one), all of eval / pdf / sample.
    plastic: bit-identical.  twosided over roughplastic: the tolerances of the one-sided test."""
    os.environ['REF_DATA_DIR'] = oracle.DATA_DIR
    distr = {'beckmann': 0.0, 'ggx': 1.0, 'phong': 2.0}
    floats = {k: (distr[v] if k == 'distribution' else float(v)) for k, v in props.items() if not isinstance(v, tuple)}
    spectra = {k: v for k, v in props.items() if isinstance(v, tuple)}
    ref = RefBSDF(plugin, floats, spectra)
    s = oracle.Scene()
    base = plugin.split(':')[-1]
    q = dict(props); q.setdefault('intIOR', 1.49); q.setdefault('extIOR', 1.000277)
    b = s.add_bsdf(base, **q)
    if plugin.startswith('twosided'):
        s.set_twosided(b)
    rng = np.random.default_rng(53)
    n = 100000
    wi = sphere_dirs(rng, n); wo = sphere_dirs(rng, n); smp = rng.random((n, 2)).astype(np.float32)
    wo[: n // 4] = wi[: n // 4] * np.array([-1, -1, 1], np.float32)             # exact mirror pairs: the delta component
    exact = base in ('plastic', 'mirror')
    for discrete in (False, True):
        rev, rp = ref.eval(wi, wo, discrete); oev, op = s.bsdf_eval(b, wi, wo, discrete)
        if exact:
            assert np.array_equal(rev, oev) and np.array_equal(rp, op)
            assert base == 'mirror' or ((rev != 0).any() and (rp != 0).any())
            if base == 'mirror':
                assert not (rev != 0).any() and (rp != 0).all() == discrete     # eval() identically zero, pdf() == 1 in the discrete measure: the file as committed
        else:
            assert np.array_equal((rev != 0).any(axis=1), (oev != 0).any(axis=1)) and np.array_equal(rp != 0, op != 0)
            assert (np.abs(rev - oev) / np.maximum(np.abs(oev).max(axis=1, keepdims=True), 1e-6)).max() <= 1e-4
    if plugin.startswith('twosided'):
        rev, _ = ref.eval(wi, wo, False)
        assert base == 'mirror' or (rev[wi[:, 2] < 0] != 0).any()                # the back side scatters as well
    rwo, rwt, rpdf, rty = ref.sample(wi, smp)
    owo, owt, opdf, oty = s.bsdf_sample(b, wi, smp)
    if exact:
        alive = (rwt != 0).any(axis=1)
        assert np.array_equal(alive, (owt != 0).any(axis=1)) and np.array_equal(rty[alive], oty[alive])
        delta = alive & ((rty & 0xff) == 0x20)                                   # EDeltaReflection (bsdf.h:84): bit-identical
        assert delta.sum() > 1000 and (base == 'mirror' or (alive & ~delta).sum() > 1000)              # both components are drawn
        assert np.array_equal(rwt[delta], owt[delta]) and np.array_equal(rpdf[delta], opdf[delta]) and np.array_equal(rwo[delta], owo[delta])
        # the diffuse lobe draws its direction with sincosf in the reference (warp.cpp:43-52) and correctly rounded in the oracle: one ulp in wo, hence in Fo
        assert np.abs(rwo[alive] - owo[alive]).max() < 1e-5 and np.abs(rwt[alive] - owt[alive]).max() < 2e-5 and np.abs(rpdf[alive] - opdf[alive]).max() < 1e-6 and np.median(np.abs(rwt[alive] - owt[alive])) == 0
    else:
        alive = (rwt != 0).any(axis=1)
        same = (alive == (owt != 0).any(axis=1)) & (~alive | (rty == oty))
        assert same.mean() > 0.999


@pytest.mark.skipif(not os.path.exists(REF_GEOM), reason='oracle/_ref/libref_geom.so not built (needs /root/reference)')
def test_rectangle_and_checkerboard_pinned_against_reference_text(cp, oracle):
    """Rectangle::configure / getAABB / rayIntersect / fillIntersectionRecord (src/shapes/rectangle.cpp:100-171) over the reference's Transform
    (transform.h / transform.cpp / matrix.inl text) and Checkerboard::eval behind Texture2D::eval (src/textures/checkerboard.cpp:65-72,
    src/librender/texture.cpp:112-121), cut out and executed as written (oracle/ref_shim/ref_rect.cpp), against the oracle: hit / miss, t, hit
    point, both normals, uv and the bounds bit-identical on the floor of models/teapot/scene.xml:56-62 and on a rotated, flipped rectangle; texture
    colours identical on 10^5 lookups including negative and huge coordinates."""
    R = ctypes.CDLL(REF_GEOM); R.ref_rect_create.restype = ctypes.c_void_p
    rng = np.random.default_rng(59)
    floor = np.array([-39.9766, 39.9766, -1.74743e-006, 0, 4.94249e-006, 2.47125e-006, -56.5355, 0, -39.9766, -39.9766, -5.2423e-006, 0, 0, 0, 0, 1], np.float32).reshape(4, 4)
    c, s_ = np.cos(0.7), np.sin(0.7)
    rot = (np.array([[c, 0, s_, 1.5], [0, 1, 0, -2.0], [-s_, 0, c, 0.25], [0, 0, 0, 1]]) @ np.diag([3.0, 0.5, 1.0, 1.0])).astype(np.float32)
    for tw, flip in ((floor, False), (rot, False), (rot, True), (np.eye(4, dtype=np.float32), True)):
        err = ctypes.create_string_buffer(256)
        h = ctypes.c_void_p(R.ref_rect_create(tw.ctypes.data_as(ctypes.c_void_p), 1 if flip else 0, err, 256)); assert h.value, err.value
        s = oracle.Scene(); b = s.add_bsdf('diffuse', reflectance=0.5)
        s.add_rectangle(tw, flip, b)
        s.set_camera(np.eye(4, dtype=np.float32), width=8, height=8); s.build()
        rb = np.zeros(6, np.float32); R.ref_rect_bounds(h, rb.ctypes.data_as(ctypes.c_void_p))
        n = 40000
        centre = (tw @ np.array([0, 0, 0, 1], np.float32))[:3]; ext = float(np.abs(tw[:3, :3]).sum(axis=1).max())
        target = centre + (rng.random((n, 3)).astype(np.float32) - 0.5) * 2.4 * ext * np.array([1, 1, 1], np.float32)
        target = (tw @ np.concatenate([(rng.random((n, 2)) * 2.6 - 1.3), np.zeros((n, 1)), np.ones((n, 1))], axis=1).T).T[:, :3].astype(np.float32)
        o = (centre + sphere_dirs(rng, n) * ext * (0.5 + 2 * rng.random((n, 1)))).astype(np.float32)
        d = target - o; d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
        mint = np.full(n, 1e-4, np.float32); maxt = np.where(rng.random(n) < 0.2, ext, np.inf).astype(np.float32)
        hit = np.zeros(n, np.int32); t = np.zeros(n, np.float32); rec = np.zeros((n, 14), np.float32)
        R.ref_rect_intersect(h, n, o.ctypes.data_as(ctypes.c_void_p), d.ctypes.data_as(ctypes.c_void_p), mint.ctypes.data_as(ctypes.c_void_p), maxt.ctypes.data_as(ctypes.c_void_p),
                             hit.ctypes.data_as(ctypes.c_void_p), t.ctypes.data_as(ctypes.c_void_p), rec.ctypes.data_as(ctypes.c_void_p))
        assert (hit >= 0).all() and 0.2 < hit.mean() < 0.9
        sh, pr, ot, orec = s.intersect_full(o, d, mint, maxt)
        # the scene query clips the ray against the (enlarged) scene box first (skdtree.cpp:112-142): compare where that box cannot matter
        ouv, ogn = s.intersect_uv(o, d, mint, maxt)
        assert np.array_equal(sh >= 0, hit == 1)
        k = hit == 1
        assert np.array_equal(ot[k], t[k]) and np.array_equal(orec[k, 0:3], rec[k, 0:3])          # t, p
        assert np.array_equal(ogn[k], rec[k, 3:6]) and np.array_equal(orec[k, 3:6], rec[k, 6:9])    # geoFrame.n, shFrame.n
        assert np.array_equal(ouv[k], rec[k, 12:14])
        bmin, bmax = s.scene_bounds()[0][:3], s.scene_bounds()[0][3:]
        assert (bmin <= rb[:3]).all() and (bmax >= rb[3:]).all() and np.allclose(bmin, rb[:3], rtol=2e-3, atol=2e-3 * ext)   # the tree box is the shape box enlarged (gkdtree.h:1213-1220)
    sheared = np.eye(4, dtype=np.float32); sheared[0, 1] = 0.3
    err = ctypes.create_string_buffer(256)
    assert not R.ref_rect_create(sheared.ctypes.data_as(ctypes.c_void_p), 0, err, 256) and b'shear' in err.value
    s = oracle.Scene(); b = s.add_bsdf('diffuse', reflectance=0.5)
    with pytest.raises(RuntimeError, match='shear'):
        s.add_rectangle(sheared, False, b)
    # checkerboard
    n = 100000
    uv = np.concatenate([rng.random((n // 2, 2)) * 4 - 2, (rng.random((n // 2, 2)) - 0.5) * 1e6]).astype(np.float32)
    uv[:8] = [[0, 0], [0.5, 0.5], [0.25, 0.75], [-0.25, 0.25], [1, 1], [-1e-8, 0.3], [0.49999997, 0.1], [2.5, -3.5]]
    c0 = np.array([0.725, 0.71, 0.68], np.float32); c1 = np.array([0.325, 0.31, 0.25], np.float32)
    for (uo, vo, us, vs) in ((0, 0, 10, 10), (0.25, -0.5, 3, 0.5), (0, 0, 1, 1)):
        ref = np.zeros((n, 3), np.float32)
        R.ref_checkerboard_eval(c0.ctypes.data_as(ctypes.c_void_p), c1.ctypes.data_as(ctypes.c_void_p), ctypes.c_float(uo), ctypes.c_float(vo), ctypes.c_float(us), ctypes.c_float(vs),
                                n, uv.ctypes.data_as(ctypes.c_void_p), ref.ctypes.data_as(ctypes.c_void_p))
        s = oracle.Scene(); b = s.add_bsdf('diffuse', reflectance=0.5); s.set_checkerboard(b, c0, c1, uo, vo, us, vs)
        wi = np.tile(np.array([[0, 0, 1]], np.float32), (n, 1))
        _, wt, _, _ = s.bsdf_sample_uv(b, wi, np.full((n, 2), 0.5, np.float32), uv)           # SmoothDiffuse::sample returns the reflectance (diffuse.cpp:140-156)
        assert np.array_equal(wt, ref) and 0.3 < (ref[:, 0] == c0[0]).mean() < 0.7


def test_teapot_scene_plugins_host_side(cp, oracle, tmp_path):
    """Host side of SURVEY 8f rank 2 without a GPU: the product's fresnelDiffuseReflectance equals the oracle's (and through it the reference text)
    bit for bit; the OBJ loader carries `vt` coordinates through the vertex merge (src/shapes/obj.cpp:633-636,672-675); a scene file with the
    plugin set of models/teapot/scene.xml validates, and the reference's own file does when the tree is present."""
    L = oracle.lib()
    for eta in np.concatenate([np.linspace(0.4, 0.98, 15), np.linspace(1.02, 3.0, 30)]).astype(np.float32):
        assert cp.fresnel_diffuse_reflectance(float(eta)) == np.float32(L.orc_fresnel_diffuse_reflectance(ctypes.c_float(eta)))
    obj = tmp_path / 'q.obj'
    obj.write_text('v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nvt 0 0\nvt 1 0\nvt 1 1\nvt 0 1\nvt 0.5 0.5\nf 1/1 2/2 3/3\nf 1/1 3/3 4/4\nf 1/5 2/2 3/3\n')
    xyz, idx, nrm, uv = cp.load_obj_file(str(obj), texcoords=True)
    assert len(xyz) == 5 and uv.shape == (5, 2)                                   # the corner with another uv is a vertex of its own
    assert np.array_equal(uv[idx[0]], [[0, 1], [1, 1], [1, 0]]) and np.array_equal(uv[idx[2][0]], [0.5, 0.5])    # flipTexCoords default: v -> 1 - v (obj.cpp:215)
    obj2 = tmp_path / 'p.obj'; obj2.write_text('v 0 0 0\nv 1 0 0\nv 1 1 0\nf 1 2 3\n')
    assert cp.load_obj_file(str(obj2), texcoords=True)[3] is None
    xml = tmp_path / 's.xml'
    xml.write_text('''<scene version="0.6.0"><sensor type="perspective"><sampler type="sobol"><integer name="sampleCount" value="4"/></sampler>
      <film type="ldrfilm"><integer name="width" value="8"/><integer name="height" value="8"/></film></sensor>
      <bsdf type="twosided" id="m"><bsdf type="plastic"><string name="intIOR" value="polypropylene"/><boolean name="nonlinear" value="true"/></bsdf></bsdf>
      <bsdf type="twosided" id="f"><bsdf type="diffuse"><texture name="reflectance" type="checkerboard"><float name="uscale" value="10"/></texture></bsdf></bsdf>
      <shape type="rectangle"><transform name="toWorld"><scale x="3" y="2"/></transform><ref id="f"/></shape>
      <shape type="rectangle"><boolean name="flipNormals" value="true"/><ref id="m"/></shape></scene>''')
    rep = cp.validate_scene_xml(str(xml))
    assert rep.count('shape rectangle') == 2 and 'bsdf plastic' in rep and 'texture checkerboard' in rep and sum('twosided' in r for r in rep) == 2
    bad = tmp_path / 'b.xml'
    bad.write_text(xml.read_text().replace('type="checkerboard"', 'type="bitmap"'))
    with pytest.raises(cp.CudapathError, match='bitmap'):
        cp.validate_scene_xml(str(bad))
    ref = '/root/reference/models/teapot/scene.xml'
    if os.path.exists(ref):
        rep = cp.validate_scene_xml(ref)
        assert 'shape rectangle' in rep and 'bsdf plastic' in rep and 'texture checkerboard' in rep and 'sampleCount 64' in rep


def _read_exr(path):
    """Minimal reader of single-part, uncompressed scan-line OpenEXR files (half or float channels): returns {channel: (h, w) float32}."""
    import struct
    b = open(path, 'rb').read()
    assert b[:4] == bytes([0x76, 0x2f, 0x31, 0x01]) and struct.unpack('<I', b[4:8])[0] == 2
    pos = 8; attrs = {}
    while b[pos] != 0:
        e = b.index(b'\0', pos); name = b[pos:e].decode(); pos = e + 1
        e = b.index(b'\0', pos); typ = b[pos:e].decode(); pos = e + 1
        size = struct.unpack('<i', b[pos:pos + 4])[0]; pos += 4
        attrs[name] = (typ, b[pos:pos + size]); pos += size
    pos += 1
    assert attrs['compression'] == ('compression', b'\0') and attrs['lineOrder'] == ('lineOrder', b'\0')
    x0, y0, x1, y1 = struct.unpack('<4i', attrs['dataWindow'][1]); w, h = x1 - x0 + 1, y1 - y0 + 1
    chans = []; c = attrs['channels'][1]; q = 0
    while c[q] != 0:
        e = c.index(b'\0', q); nm = c[q:e].decode(); q = e + 1
        pt = struct.unpack('<i', c[q:q + 4])[0]; q += 16; chans.append((nm, pt))
    assert [n for n, _ in chans] == sorted(n for n, _ in chans)
    offs = struct.unpack('<%dQ' % h, b[pos:pos + 8 * h])
    out = {n: np.zeros((h, w), np.float32) for n, _ in chans}
    for y in range(h):
        p = offs[y]; yy, nbytes = struct.unpack('<ii', b[p:p + 8]); p += 8
        assert yy == y0 + y
        for n, pt in chans:
            dt, sz = (np.float16, 2) if pt == 1 else (np.float32, 4)
            out[n][y] = np.frombuffer(b, dt, w, p).astype(np.float32); p += sz * w
    return out


def test_exr_writer(cp, tmp_path):
    """hdrfilm's default output (src/films/hdrfilm.cpp:213-246: openexr, rgb, float16): cudapath_write_exr against an independent parse of the
    file format and, where OpenCV was built with OpenEXR, against its reader; the half conversion against numpy's IEEE round-to-nearest-even."""
    rng = np.random.default_rng(71)
    x = np.concatenate([rng.normal(size=20000) * 10.0 ** rng.integers(-9, 6, 20000), [0, -0.0, 65504, 65519.99, 65520, 1e9, -1e9, np.inf, -np.inf, 6e-8, 2.98e-8, 2.99e-8, 5.96e-8, 6.1e-5,
                                                                                       6.097e-5, 1.0009765625, 1.00048828125, 1.00146484375]]).astype(np.float32)
    hq = np.zeros(len(x), np.uint16)
    assert cp.lib().cudapath_float_to_half(x.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint64(len(x)), hq.ctypes.data_as(ctypes.c_void_p)) == 0
    with np.errstate(over='ignore'):
        assert np.array_equal(hq, x.astype(np.float16).view(np.uint16))
    w, h = 37, 23
    img = (rng.random((h, w, 3)) * np.array([1.0, 50.0, 0.01])).astype(np.float32); img[3, 4] = (0, 1e6, 65504)
    for half in (1, 0):
        path = str(tmp_path / ('a%d.exr' % half))
        assert cp.lib().cudapath_write_exr(path.encode(), img.ctypes.data_as(ctypes.c_void_p), w, h, half) == 0
        ch = _read_exr(path)
        with np.errstate(over='ignore'):
            want = img.astype(np.float16).astype(np.float32) if half else img
        assert np.array_equal(ch['R'], want[..., 0]) and np.array_equal(ch['G'], want[..., 1]) and np.array_equal(ch['B'], want[..., 2])
        os.environ['OPENCV_IO_ENABLE_OPENEXR'] = '1'
        try:
            import cv2
            im = cv2.imread(path, cv2.IMREAD_UNCHANGED)
        except Exception:
            im = None
        if im is not None:
            assert im.shape == (h, w, 3) and np.array_equal(im[..., ::-1].astype(np.float32), want)
    assert cp.lib().cudapath_write_exr(str(tmp_path / 'no' / 'such' / 'dir.exr').encode(), img.ctypes.data_as(ctypes.c_void_p), w, h, 1) != 0


@pytest.mark.skipif(not os.path.isdir('/root/reference/models'), reason='needs /root/reference')
def test_every_reference_scene_file_validates(cp):
    """Every scene file the reference ships under models/ goes through the loader's dry run: all of them parse and name only plugins of this
    path -- except the three the reference itself cannot load: scene_dielectric2.xml is not well-formed (an unclosed <float>), teapot/dielectric.xml
    nests a transmissive `dielectric` in `twosided` (TwoSidedBRDF::configure raises, twosided.cpp:106-108: same message here) and
    teapot/mirror_scene.xml runs under the irrcache / photonmapper integrators, which are not this path (its `mirror` BSDF, the fork's own plugin, is)."""
    import glob
    expected_failures = {'straight-hair/scene_dielectric2.xml': 'parse error', 'teapot/dielectric.xml': 'transmission component', 'teapot/mirror_scene.xml': 'irrcache'}
    files = sorted(glob.glob('/root/reference/models/*/*.xml'))
    assert len(files) >= 17
    for f in files:
        key = f.split('models/')[1]
        if key in expected_failures:
            with pytest.raises(cp.CudapathError, match=expected_failures[key]):
                cp.validate_scene_xml(f)
            continue
        rep = cp.validate_scene_xml(f)
        assert any(r.startswith('bsdf') for r in rep) and any(r.startswith('shape') for r in rep) and rep[-1].startswith('sampleCount'), key

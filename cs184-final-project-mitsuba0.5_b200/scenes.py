"""Scene definitions for the four hair scenes named in BASELINE.json, plus seeded procedural fiber generators.

The reference's geometry blobs (models/*/models/*.mitshair) are missing (.MISSING_LARGE_BLOBS), so every scene is
described here by the *parameters* of its XML file (camera matrix, fov, fiber radius, sun direction, BSDF block -- cited
per scene) and a deterministic procedural stand-in for the fibers, sized and placed to fill the XML camera's frustum.
`write_scene()` emits a scene directory (scene.xml + models/*.mitshair in the reference's BINARY_HAIR format,
src/shapes/hair.cpp:92-98,641-679) so that the unchanged loader path (XML -> HairShape file loader) is exercised.
Real .mitshair files can be dropped into the same directory layout instead.
"""
import os
import struct
import numpy as np

# ------------------------------------------------------------------------------------------------ scene parameters
_CAM_STRAIGHT = [0.999887, 0.00390257, 0.0145262, -0.234672, 6.98571e-010, 0.965755, -0.259457, 16.5124,
                 -0.0150413, 0.259428, 0.965645, -25.3482, 0, 0, 0, 1]       # models/straight-hair/scene_kkay.xml:12, curly-hair/scene.xml:12
_CAM_CURL = [-1, 4.24672e-010, 1.50958e-007, -0.055286, 1.11022e-016, 0.999996, -0.00281317, 5.92976,
             -1.50959e-007, -0.00281317, -0.999996, 17.0651, 0, 0, 0, 1]      # models/hair-curl/marschner_scene.xml:12
_CAM_FURBALL = [-0.704024, 0.0939171, 0.703939, -10.6677, 1.05829e-008, 0.991217, -0.132245, 14.3141,
                -0.710177, -0.0931033, -0.69784, 10.2879, 0, 0, 0, 1]         # models/furball/scene.xml:12
_SUN_A = (0.19033, 0.758426, -0.623349)     # straight-hair / curly-hair scene files
_SUN_B = (-0.376047, 0.758426, 0.532333)    # hair-curl / furball scene files
_SUNSKY = dict(turbidity=3.0, skyScale=5.0, sunScale=19.0912, sunRadiusScale=37.9165)
_HAIR_RGB = (0.143016, 0.0156076, 1.80928e-005)
# Marschner block of models/straight-hair/scene_marschner.xml:31-39
_MARSCHNER_C3 = dict(type='marschner', alpha=0.2, distribution='ggx', intIOR=1.55, extIOR=1.0, diffuseReflectance=_HAIR_RGB)
# roughplastic block of the DEFAULT scene files, models/{straight-hair,curly-hair,furball}/scene.xml:31-38
_ROUGHPLASTIC = dict(type='roughplastic', alpha=0.2, distribution='ggx', intIOR=1.55, extIOR=1.0, nonlinear=False, diffuseReflectance=_HAIR_RGB)
_KKAY = dict(type='kajiyakay', diffuseReflectance=_HAIR_RGB, exponent=10.0)   # models/straight-hair/scene_kkay.xml:31-34

SCENES = {
    # C1: models/straight-hair/scene_kkay.xml with maxDepth=8, 512x512, 16 spp
    'straight-hair': dict(camera=_CAM_STRAIGHT, fov=35.0, sun=_SUN_A, width=512, height=512, spp=16, maxDepth=8,
                          shapes=[dict(generator='straight', radius=0.00566563, bsdf=dict(_KKAY, id='hair'))]),
    # C2: models/hair-curl/marschner_scene.xml at 1024x1024, 64 spp, maxDepth 65; four shapes / four marschner BSDFs that only
    # differ in specularReflectance (marschner_scene.xml:31-93), radius 0.000444
    'hair-curl': dict(camera=_CAM_CURL, fov=35.0, sun=_SUN_B, width=1024, height=1024, spp=64, maxDepth=65,
                      shapes=[dict(generator='curl', group=g, radius=0.000444,
                                   bsdf=dict(type='marschner', id=name, intIOR=1.55, extIOR=1.0, specularReflectance=spec))
                              for g, (name, spec) in enumerate([('black_hair', (6.344e-006, 7.62186e-012, 6.53751e-030)),
                                                                ('red_hair', (0.0112431, 6.77287e-005, 1.13705e-011)),
                                                                ('brown_hair', (0.143016, 0.0156076, 1.80928e-005)),
                                                                ('blonde_hair', (0.592384, 0.32628, 0.0528657))])]),
    # C3: curly-hair geometry/camera (models/curly-hair/scene.xml) + the Marschner block, 1024x1024, 1024 spp
    'curly-hair': dict(camera=_CAM_STRAIGHT, fov=35.0, sun=_SUN_A, width=1024, height=1024, spp=1024, maxDepth=65,
                       shapes=[dict(generator='curly', radius=0.00559955, bsdf=dict(_MARSCHNER_C3, id='hair'))]),
    # C4: models/furball/scene.xml geometry/camera + the Marschner block, maxDepth 32, 2048x2048, 256 spp
    'furball': dict(camera=_CAM_FURBALL, fov=35.0, sun=_SUN_B, width=2048, height=2048, spp=256, maxDepth=32,
                    shapes=[dict(generator='furball', radius=0.00216667, bsdf=dict(_MARSCHNER_C3, id='hair'))]),
    # models/straight-hair/scene.xml as shipped (roughplastic on the fibers), reduced to 512x512 / 16 spp / depth 8
    'straight-hair-default': dict(camera=_CAM_STRAIGHT, fov=35.0, sun=_SUN_A, width=512, height=512, spp=16, maxDepth=8,
                                  shapes=[dict(generator='straight', radius=0.00566563, bsdf=dict(_ROUGHPLASTIC, id='hair'))]),
    # models/straight-hair/scene_dielectric.xml:31-38 (the fork's `marschnerdielectric`) and scene_thindielectric.xml:31-37
    # (same block with type `thindielectric`, no diffuse colour), reduced to 512x512 / 16 spp; maxDepth 65 as in the files
    'straight-hair-dielectric': dict(camera=_CAM_STRAIGHT, fov=35.0, sun=_SUN_A, width=512, height=512, spp=16, maxDepth=65,
                                     shapes=[dict(generator='straight', radius=0.00566563,
                                                  bsdf=dict(type='marschnerdielectric', id='hair', intIOR=1.55, extIOR=1.0, exponent=5.0, specularTransmittance=_HAIR_RGB,
                                                            specularReflectance=_HAIR_RGB, diffuseReflectance=_HAIR_RGB))]),
    'straight-hair-thindielectric': dict(camera=_CAM_STRAIGHT, fov=35.0, sun=_SUN_A, width=512, height=512, spp=16, maxDepth=65,
                                         shapes=[dict(generator='straight', radius=0.00566563,
                                                      bsdf=dict(type='thindielectric', id='hair', intIOR=1.55, extIOR=1.0, specularTransmittance=_HAIR_RGB,
                                                                specularReflectance=_HAIR_RGB))]),
    # T1: the straight-hair fibers on a head -- an ellipsoid mesh with smooth vertex normals under the scalp points and a
    # ground quad with face normals (both `diffuse`, the quad inside `twosided`), i.e. fibers and triangles in one BVH
    'hair-on-head': dict(camera=_CAM_STRAIGHT, fov=35.0, sun=_SUN_A, width=512, height=512, spp=16, maxDepth=8,
                         shapes=[dict(generator='straight', radius=0.00566563, bsdf=dict(_KKAY, id='hair')),
                                 dict(mesh='ellipsoid', center=(0.15, 13.0, 0.5), radii=(3.55, 4.15, 3.95), res=48,
                                      bsdf=dict(type='diffuse', id='skin', reflectance=(0.55, 0.38, 0.30))),
                                 dict(mesh='quad', center=(0.15, -1.0, 0.0), half=(14.0, 14.0),
                                      bsdf=dict(type='twosided', id='ground', reflectance=(0.4, 0.4, 0.4)))]),
}


# ------------------------------------------------------------------------------------------------ mesh generators
def gen_ellipsoid(center, radii, res=48):
    """UV-sphere scaled to an ellipsoid: (xyz (n,3) f32, idx (m,3) u32, unit vertex normals (n,3) f32)."""
    lat = np.linspace(0.0, np.pi, res + 1)[1:-1]
    lon = np.linspace(0.0, 2 * np.pi, 2 * res, endpoint=False)
    d = [np.array([[0.0, 1.0, 0.0]])]
    for th in lat:
        d.append(np.stack([np.sin(th) * np.cos(lon), np.full_like(lon, np.cos(th)), np.sin(th) * np.sin(lon)], axis=1))
    d.append(np.array([[0.0, -1.0, 0.0]]))
    d = np.concatenate(d)
    radii = np.asarray(radii, np.float64); center = np.asarray(center, np.float64)
    xyz = center + d * radii
    nrm = d / radii
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    L = 2 * res
    idx = []
    ring = lambda r: 1 + r * L
    for j in range(L):
        idx.append((0, ring(0) + (j + 1) % L, ring(0) + j))
    for r in range(len(lat) - 1):
        for j in range(L):
            a, b = ring(r) + j, ring(r) + (j + 1) % L
            c, e = ring(r + 1) + j, ring(r + 1) + (j + 1) % L
            idx.append((a, b, e)); idx.append((a, e, c))
    last = len(d) - 1
    for j in range(L):
        idx.append((last, ring(len(lat) - 1) + j, ring(len(lat) - 1) + (j + 1) % L))
    return xyz.astype(np.float32), np.asarray(idx, np.uint32), nrm.astype(np.float32)


def gen_quad(center, half):
    """Horizontal quad (two triangles, counter-clockwise seen from +y), no vertex normals (face normals)."""
    cx, cy, cz = center; hx, hz = half
    xyz = np.array([[cx - hx, cy, cz - hz], [cx - hx, cy, cz + hz], [cx + hx, cy, cz + hz], [cx + hx, cy, cz - hz]], np.float32)
    return xyz, np.array([[0, 1, 2], [0, 2, 3]], np.uint32), None


def generate_mesh(shape_desc):
    if shape_desc['mesh'] == 'ellipsoid':
        return gen_ellipsoid(shape_desc['center'], shape_desc['radii'], shape_desc.get('res', 48))
    if shape_desc['mesh'] == 'quad':
        return gen_quad(shape_desc['center'], shape_desc['half'])
    raise ValueError('unknown mesh generator %r' % shape_desc['mesh'])


def add_shapes(target, sc, scale=1.0):
    """Adds the BSDFs and shapes of a scene description to `target` (a cudapath.Context or its oracle twin: same method names)."""
    for sh in sc['shapes']:
        b = dict(sh['bsdf']); btype = b.pop('type'); b.pop('id', None)
        bid = target.add_bsdf(btype, **b)
        if 'mesh' in sh:
            xyz, idx, nrm = generate_mesh(sh)
            target.add_mesh(xyz, idx, bid, normals=nrm)
        else:
            xyz, starts = generate(sh, scale)
            target.add_hair(xyz, starts, sh['radius'], bid)


# ------------------------------------------------------------------------------------------------ fiber generators
def _strand_arrays(points):
    """points: (strands, verts, 3) float32 -> flat xyz (n,3), starts_fiber (n,) uint8"""
    s, v, _ = points.shape
    xyz = np.ascontiguousarray(points.reshape(s * v, 3), dtype=np.float32)
    starts = np.zeros(s * v, np.uint8)
    starts[::v] = 1
    return xyz, starts


def _scalp_points(rng, n, center, radii, min_y):
    """uniform-ish points on the upper part of an ellipsoid (rejection on height)"""
    out = np.zeros((0, 3))
    while len(out) < n:
        d = rng.normal(size=(2 * n, 3))
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        d = d[d[:, 1] > min_y]
        out = np.concatenate([out, d])
    d = out[:n]
    return center + d * radii, d


def gen_straight(strands=50000, segments=25, seed=1):
    """Straight-ish hair hanging from a scalp ellipsoid in front of the straight-hair camera (target ~ (0.15, 9.7, 0))."""
    rng = np.random.default_rng(seed)
    root, nrm = _scalp_points(rng, strands, np.array([0.15, 13.0, 0.5]), np.array([3.6, 4.2, 4.0]), 0.15)
    t = np.linspace(0.0, 1.0, segments + 1)[None, :, None]
    length = rng.uniform(9.0, 13.0, size=(strands, 1, 1))
    out_dir = nrm.copy(); out_dir[:, 1] = 0.0
    out_dir /= np.maximum(np.linalg.norm(out_dir, axis=1, keepdims=True), 1e-3)
    flare = rng.uniform(0.5, 1.6, size=(strands, 1, 1))
    # leave the scalp along the normal, then fall under gravity; a gentle per-strand wave keeps every joint above the
    # loader's 1-degree merge threshold so that the segment count survives HairShape's vertex merge
    phase = rng.uniform(0, 2 * np.pi, size=(strands, 1, 1))
    freq = rng.uniform(5.0, 9.0, size=(strands, 1, 1))
    amp = rng.uniform(0.05, 0.12, size=(strands, 1, 1))
    side = np.cross(out_dir, np.array([0.0, 1.0, 0.0]))
    p = root[:, None, :] + nrm[:, None, :] * (0.6 * (1 - np.exp(-4 * t))) \
        + out_dir[:, None, :] * (flare * t ** 1.5) \
        + np.array([0.0, -1.0, 0.0])[None, None, :] * (length * t ** 1.3) \
        + side[:, None, :] * (amp * np.sin(freq * t * 2 * np.pi + phase)) \
        + out_dir[:, None, :] * (amp * np.cos(freq * t * 2 * np.pi + phase))
    return _strand_arrays(p.astype(np.float32))


def gen_curly(strands=50000, segments=68, seed=2):
    """Helical curls hanging from the same scalp (models/curly-hair camera = straight-hair camera)."""
    rng = np.random.default_rng(seed)
    root, nrm = _scalp_points(rng, strands, np.array([0.15, 13.0, 0.5]), np.array([3.6, 4.2, 4.0]), 0.15)
    t = np.linspace(0.0, 1.0, segments + 1)[None, :, None]
    length = rng.uniform(7.0, 11.0, size=(strands, 1, 1))
    out_dir = nrm.copy(); out_dir[:, 1] = 0.0
    out_dir /= np.maximum(np.linalg.norm(out_dir, axis=1, keepdims=True), 1e-3)
    side = np.cross(out_dir, np.array([0.0, 1.0, 0.0]))
    turns = rng.uniform(5.0, 9.0, size=(strands, 1, 1))
    rad = rng.uniform(0.18, 0.42, size=(strands, 1, 1)) * (0.3 + 0.7 * t)
    phase = rng.uniform(0, 2 * np.pi, size=(strands, 1, 1))
    ang = turns * 2 * np.pi * t + phase
    p = root[:, None, :] + nrm[:, None, :] * (0.5 * (1 - np.exp(-4 * t))) \
        + out_dir[:, None, :] * (rng.uniform(0.8, 2.2, size=(strands, 1, 1)) * t ** 1.4) \
        + np.array([0.0, -1.0, 0.0])[None, None, :] * (length * t ** 1.2) \
        + side[:, None, :] * (rad * np.cos(ang)) + out_dir[:, None, :] * (rad * np.sin(ang))
    return _strand_arrays(p.astype(np.float32))


def gen_curl(group=0, strands=10000, segments=100, seed=3):
    """One large curl (a lock of hair wound into a descending spiral); four colour groups sit side by side in front of the
    hair-curl camera (target ~ (-0.05, 5.9, 0), visible width ~ 10.8)."""
    rng = np.random.default_rng(seed + 17 * group)
    t = np.linspace(0.0, 1.0, segments + 1)[None, :, None]
    cx = -3.6 + 2.4 * group
    turns = 3.0
    ang = turns * 2 * np.pi * t
    R = 0.85 * (1.0 - 0.25 * t)
    center = np.concatenate([cx + R * np.cos(ang), 10.3 - 8.6 * t, R * np.sin(ang)], axis=2)      # (1, v, 3)
    # offsets inside the lock cross-section (disc of radius 0.32), rotating slowly along the lock
    r = 0.32 * np.sqrt(rng.uniform(size=(strands, 1, 1)))
    a0 = rng.uniform(0, 2 * np.pi, size=(strands, 1, 1))
    twist = a0 + 1.5 * 2 * np.pi * t
    radial = np.concatenate([np.cos(ang), np.zeros_like(ang), np.sin(ang)], axis=2)
    up = np.array([0.0, 1.0, 0.0])[None, None, :]
    jitter = 0.012 * np.sin(rng.uniform(20, 40, size=(strands, 1, 1)) * t * 2 * np.pi + a0)
    p = center + radial * (r * np.cos(twist) + jitter) + up * (r * np.sin(twist))
    return _strand_arrays(p.astype(np.float32))


def gen_furball(strands=200000, segments=8, seed=4):
    """Short fibers on a sphere in front of the furball camera (target ~ (0.03, 12.3, -0.3), visible width ~ 9.6)."""
    rng = np.random.default_rng(seed)
    d = rng.normal(size=(strands, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    center = np.array([0.03, 12.3, -0.3])
    root = center + 2.6 * d
    t = np.linspace(0.0, 1.0, segments + 1)[None, :, None]
    length = rng.uniform(0.9, 1.5, size=(strands, 1, 1))
    bend = rng.normal(size=(strands, 3)); bend -= d * np.sum(bend * d, axis=1, keepdims=True)
    bend /= np.maximum(np.linalg.norm(bend, axis=1, keepdims=True), 1e-3)
    curl = rng.uniform(0.25, 0.6, size=(strands, 1, 1))
    p = root[:, None, :] + d[:, None, :] * (length * t) + bend[:, None, :] * (curl * t ** 2) \
        + np.array([0.0, -1.0, 0.0])[None, None, :] * (0.25 * t ** 2)
    return _strand_arrays(p.astype(np.float32))


GENERATORS = {'straight': gen_straight, 'curly': gen_curly, 'curl': gen_curl, 'furball': gen_furball}


def generate(shape_desc, scale=1.0):
    """scale < 1 reduces the strand count (parity tests use small scenes; bench uses scale=1)."""
    kw = {}
    g = shape_desc['generator']
    if 'group' in shape_desc:
        kw['group'] = shape_desc['group']
    fn = GENERATORS[g]
    default_strands = fn.__defaults__[0 if g != 'curl' else 1]
    kw['strands'] = max(8, int(round(default_strands * scale)))
    return fn(**kw)


# ------------------------------------------------------------------------------------------------ file emitters
def write_mitshair(path, xyz, starts):
    """BINARY_HAIR: 11-byte magic, u32 vertex count, then xyz float32 triples; +inf before a triple starts a fiber."""
    n = len(starts)
    rec = np.zeros((n, 4), np.float32)
    rec[:, 0] = np.inf
    rec[:, 1:] = xyz
    flat = rec.reshape(-1)
    keep = np.ones(n * 4, bool)
    keep[0::4] = starts.astype(bool)
    keep[0] = False          # the first vertex starts a fiber implicitly (hair.cpp:662)
    with open(path, 'wb') as f:
        f.write(b'BINARY_HAIR')
        f.write(struct.pack('<I', n))
        f.write(flat[keep].astype('<f4').tobytes())


def write_obj(path, xyz, idx, normals=None):
    """Wavefront OBJ with `v` / `vn` / `f a//a` records (%.9g round-trips fp32 exactly)."""
    with open(path, 'w') as f:
        for p in xyz:
            f.write('v %.9g %.9g %.9g\n' % tuple(float(c) for c in p))
        if normals is not None:
            for n in normals:
                f.write('vn %.9g %.9g %.9g\n' % tuple(float(c) for c in n))
        for t in idx:
            a, b, c = (int(i) + 1 for i in t)
            f.write(('f %d//%d %d//%d %d//%d\n' % (a, a, b, b, c, c)) if normals is not None else ('f %d %d %d\n' % (a, b, c)))


def _fmt(v):
    return ', '.join(repr(float(x)) for x in v)


def scene_xml(name, overrides=None):
    """XML text in the reference's scene format (version 0.6.0) for one of SCENES."""
    sc = dict(SCENES[name]); sc.update(overrides or {})
    lines = ['<?xml version="1.0" encoding="utf-8"?>', '<scene version="0.6.0">',
             '\t<integrator type="path">', '\t\t<integer name="maxDepth" value="%d"/>' % sc['maxDepth'],
             '\t\t<boolean name="strictNormals" value="true"/>', '\t</integrator>',
             '\t<sensor type="perspective">', '\t\t<float name="fov" value="%r"/>' % sc['fov'],
             '\t\t<transform name="toWorld">', '\t\t\t<matrix value="%s"/>' % ' '.join(repr(float(x)) for x in sc['camera']), '\t\t</transform>',
             '\t\t<sampler type="independent">', '\t\t\t<integer name="sampleCount" value="%d"/>' % sc['spp'], '\t\t</sampler>',
             '\t\t<film type="ldrfilm">', '\t\t\t<integer name="width" value="%d"/>' % sc['width'], '\t\t\t<integer name="height" value="%d"/>' % sc['height'],
             '\t\t\t<string name="pixelFormat" value="rgb"/>', '\t\t\t<rfilter type="tent"/>', '\t\t</film>', '\t</sensor>']
    for i, sh in enumerate(sc['shapes']):
        b = sh['bsdf']
        if b['type'] == 'twosided':
            lines += ['\t<bsdf type="twosided" id="%s">' % b['id'], '\t\t<bsdf type="diffuse">',
                      '\t\t\t<rgb name="reflectance" value="%s"/>' % _fmt(b['reflectance']), '\t\t</bsdf>', '\t</bsdf>']
            continue
        lines.append('\t<bsdf type="%s" id="%s">' % (b['type'], b['id']))
        for k, v in b.items():
            if k in ('type', 'id'):
                continue
            if isinstance(v, bool):
                lines.append('\t\t<boolean name="%s" value="%s"/>' % (k, 'true' if v else 'false'))
            elif isinstance(v, str):
                lines.append('\t\t<string name="%s" value="%s"/>' % (k, v))
            elif isinstance(v, (tuple, list)):
                lines.append('\t\t<rgb name="%s" value="%s"/>' % (k, _fmt(v)))
            else:
                lines.append('\t\t<float name="%s" value="%r"/>' % (k, float(v)))
        lines.append('\t</bsdf>')
    for i, sh in enumerate(sc['shapes']):
        if 'mesh' in sh:
            lines += ['\t<shape type="obj">', '\t\t<string name="filename" value="models/%s.obj"/>' % sh['bsdf']['id']]
            if sh['mesh'] == 'quad':
                lines.append('\t\t<boolean name="faceNormals" value="true"/>')
            lines += ['\t\t<ref id="%s"/>' % sh['bsdf']['id'], '\t</shape>']
            continue
        lines += ['\t<shape type="hair">', '\t\t<float name="radius" value="%r"/>' % sh['radius'],
                  '\t\t<string name="filename" value="models/%s.mitshair"/>' % sh['bsdf']['id'], '\t\t<ref id="%s"/>' % sh['bsdf']['id'], '\t</shape>']
    lines += ['\t<emitter type="sunsky">', '\t\t<float name="turbidity" value="%r"/>' % _SUNSKY['turbidity'],
              '\t\t<vector name="sunDirection" x="%r" y="%r" z="%r"/>' % tuple(sc['sun']),
              '\t\t<float name="skyScale" value="%r"/>' % _SUNSKY['skyScale'], '\t\t<float name="sunScale" value="%r"/>' % _SUNSKY['sunScale'],
              '\t\t<float name="sunRadiusScale" value="%r"/>' % _SUNSKY['sunRadiusScale'], '\t</emitter>', '</scene>']
    return '\n'.join(lines) + '\n'


def write_scene(name, directory, scale=1.0, overrides=None):
    """Writes <directory>/scene.xml and <directory>/models/<id>.mitshair; returns the xml path."""
    os.makedirs(os.path.join(directory, 'models'), exist_ok=True)
    sc = SCENES[name]
    for sh in sc['shapes']:
        if 'mesh' in sh:
            xyz, idx, nrm = generate_mesh(sh)
            write_obj(os.path.join(directory, 'models', sh['bsdf']['id'] + '.obj'), xyz, idx, nrm)
            continue
        xyz, starts = generate(sh, scale)
        write_mitshair(os.path.join(directory, 'models', sh['bsdf']['id'] + '.mitshair'), xyz, starts)
    path = os.path.join(directory, 'scene.xml')
    with open(path, 'w') as f:
        f.write(scene_xml(name, overrides))
    return path


def sunsky_params(name):
    return dict(_SUNSKY, sunDirection=SCENES[name]['sun'], albedo=(0.2, 0.2, 0.2), resolution=512)

"""ctypes wrapper around the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY -- the checker, never the product."""
import ctypes
import os
import subprocess
import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(REPO, 'oracle')
ORACLE_LIB = os.path.join(ORACLE_DIR, 'liboracle.so')
REF_LIB = os.path.join(ORACLE_DIR, '_ref', 'libref_pieces.so')
DATA_DIR = os.path.join(REPO, 'refdata')
REF_PATH_LIB = os.path.join(REPO, 'oracle', '_ref', 'libref_path.so')

_lib = None


def build():
    subprocess.check_call(['make', '-s', '-C', ORACLE_DIR])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_LIB):
            build()
        L = ctypes.CDLL(ORACLE_LIB)
        L.orc_scene_create.restype = ctypes.c_void_p
        L.orc_last_error.restype = ctypes.c_char_p
        L.orc_accel_description.restype = ctypes.c_char_p
        L.orc_hair_file_load.restype = ctypes.c_void_p
        L.orc_hair_file_load_reduced.restype = ctypes.c_void_p
        L.orc_hair_file_vertex_count.restype = ctypes.c_uint32
        L.orc_hair_file_segment_count.restype = ctypes.c_uint32
        L.orc_hair_file_radius.restype = ctypes.c_float
        L.orc_fresnel_diffuse_reflectance.restype = ctypes.c_float
        _lib = L
    return _lib


def accel_description():
    """What accelerates the ray queries of the loaded oracle build (the timing build differs from the checker)."""
    return lib().orc_accel_description().decode()


def have_ref():
    return os.path.exists(REF_LIB)


def ref_lib():
    return ctypes.CDLL(REF_LIB)


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def check(rc):
    if rc < 0:
        raise RuntimeError(lib().orc_last_error().decode())
    return rc


IDENT = np.eye(4, dtype=np.float32)
DISTR = {'beckmann': 0, 'ggx': 1, 'phong': 2}
FILTERS = {'tent': 0, 'box': 1, 'gaussian': 2}


def load_hair_file(path, radius=0.025, angleThreshold=1.0, toWorld=None, reduction=0.0):
    L = lib()
    tw = f32(IDENT if toWorld is None else toWorld).reshape(16)
    h = L.orc_hair_file_load_reduced(path.encode(), ctypes.c_float(radius), ctypes.c_float(angleThreshold), p(tw), ctypes.c_float(reduction))
    if not h:
        raise RuntimeError(L.orc_last_error().decode())
    h = ctypes.c_void_p(h)
    n = L.orc_hair_file_vertex_count(h)
    xyz = np.zeros((n, 3), np.float32); st = np.zeros(n, np.uint8)
    L.orc_hair_file_copy(h, p(xyz), p(st))
    r = L.orc_hair_file_radius(h); nseg = L.orc_hair_file_segment_count(h)
    L.orc_hair_file_free(h)
    return xyz, st, float(r), int(nseg)


def matrix_invert(m):
    """Matrix<4,4,float>::invert as the oracle restates it (o_math.h); None when singular."""
    m = f32(m).reshape(16); out = np.zeros(16, np.float32)
    ok = lib().orc_matrix_invert(p(m), p(out))
    return out.reshape(4, 4) if ok == 1 else None


def load_rgbe(path):
    w = ctypes.c_int(); h = ctypes.c_int()
    check(lib().orc_load_rgbe(str(path).encode(), None, ctypes.byref(w), ctypes.byref(h)))
    out = np.zeros((h.value, w.value, 3), np.float32)
    check(lib().orc_load_rgbe(str(path).encode(), p(out), ctypes.byref(w), ctypes.byref(h)))
    return out


def bake_sunsky(turbidity=3.0, albedo=0.2, sunDirection=(0, 1, 0), skyScale=1.0, sunScale=1.0, sunRadiusScale=1.0, resolution=512):
    out = np.zeros((resolution // 2, resolution, 3), np.float32)
    check(lib().orc_bake_sunsky(REF_LIB.encode(), ctypes.c_float(turbidity), ctypes.c_float(albedo), p(f32(sunDirection)), ctypes.c_float(skyScale),
                                ctypes.c_float(sunScale), ctypes.c_float(sunRadiusScale), int(resolution), p(out)))
    return out


class Scene:
    def __init__(self):
        self.L = lib()
        self.h = ctypes.c_void_p(self.L.orc_scene_create())
        self.width = self.height = 0

    def __del__(self):
        try:
            self.L.orc_scene_destroy(self.h)
        except Exception:
            pass

    def add_bsdf(self, type, **props):
        if type == 'kajiyakay':
            d = f32(np.broadcast_to(props.get('diffuseReflectance', 0.5), 3)); s = f32(np.broadcast_to(props.get('specularReflectance', 0.2), 3))
            return check(self.L.orc_add_bsdf_kajiyakay(self.h, p(d), p(s), ctypes.c_float(props.get('exponent', 30.0))))
        if type == 'roughplastic':
            d = f32(np.broadcast_to(props.get('diffuseReflectance', 0.5), 3)); s = f32(np.broadcast_to(props.get('specularReflectance', 1.0), 3))
            return check(self.L.orc_add_bsdf_roughplastic(self.h, ctypes.c_float(props.get('intIOR', 1.49)), ctypes.c_float(props.get('extIOR', 1.000277)), p(d), p(s),
                                                          ctypes.c_float(props.get('alpha', 0.1)), DISTR[props.get('distribution', 'beckmann')],
                                                          1 if props.get('sampleVisible', True) else 0, 1 if props.get('nonlinear', False) else 0, DATA_DIR.encode()))
        if type == 'marschner_fixed':
            return check(self.L.orc_add_bsdf_marschner_fixed(self.h, ctypes.c_float(props.get('intIOR', 1.55)), ctypes.c_float(props.get('extIOR', 1.000277))))
        if type == 'marschner_full':
            sa = f32(np.broadcast_to(props.get('sigmaA', 0.22), 3))
            return check(self.L.orc_add_bsdf_marschner_full(self.h, ctypes.c_float(props.get('intIOR', 1.55)), ctypes.c_float(props.get('extIOR', 1.000277)), p(sa),
                                                            ctypes.c_float(props.get('betaR', 0.1)), ctypes.c_float(props.get('scaleAngleRad', -0.1)), int(props.get('lobes', 7))))
        if type == 'thindielectric':
            r = f32(np.broadcast_to(props.get('specularReflectance', 1.0), 3)); t = f32(np.broadcast_to(props.get('specularTransmittance', 1.0), 3))
            return check(self.L.orc_add_bsdf_thindielectric(self.h, ctypes.c_float(props.get('intIOR', 1.5046)), ctypes.c_float(props.get('extIOR', 1.000277)), p(r), p(t)))
        if type == 'marschnerdielectric':
            d = f32(np.broadcast_to(props.get('diffuseReflectance', 0.5), 3))
            r = f32(np.broadcast_to(props.get('specularReflectance', 0.1), 3)); t = f32(np.broadcast_to(props.get('specularTransmittance', 0.1), 3))
            return check(self.L.orc_add_bsdf_marschnerdielectric(self.h, ctypes.c_float(props.get('intIOR', 1.501)), ctypes.c_float(props.get('extIOR', 1.000277)), p(d), p(r), p(t),
                                                                 ctypes.c_float(props.get('exponent', 30.0))))
        if type in ('diffuse', 'twosided'):
            r = f32(np.broadcast_to(props.get('reflectance', 0.5), 3))
            return check(self.L.orc_add_bsdf_diffuse(self.h, p(r), 1 if (type == 'twosided' or props.get('twoSided', False)) else 0))
        if type == 'mirror':
            return check(self.L.orc_add_bsdf_mirror(self.h, p(f32(np.broadcast_to(props.get('specularReflectance', 1.0), 3)))))
        if type == 'plastic':
            d = f32(np.broadcast_to(props.get('diffuseReflectance', 0.5), 3)); s = f32(np.broadcast_to(props.get('specularReflectance', 1.0), 3))
            return check(self.L.orc_add_bsdf_plastic(self.h, ctypes.c_float(props.get('intIOR', 1.49)), ctypes.c_float(props.get('extIOR', 1.000277)), p(d), p(s),
                                                     1 if props.get('nonlinear', False) else 0))
        d = f32(np.broadcast_to(props.get('diffuseReflectance', 0.5), 3)); s = f32(np.broadcast_to(props.get('specularReflectance', 0.5), 3))
        return check(self.L.orc_add_bsdf_marschner(self.h, ctypes.c_float(props.get('intIOR', 1.5046)), ctypes.c_float(props.get('extIOR', 1.000277)), p(d), p(s),
                                                   ctypes.c_float(props.get('alpha', 0.1)), DISTR[props.get('distribution', 'beckmann')],
                                                   1 if props.get('nonlinear', False) else 0, DATA_DIR.encode()))

    def set_sampler(self, kind='philox', scramble=0):
        check(self.L.orc_set_sampler(self.h, {'philox': 0, 'sobol': 1}[kind], ctypes.c_uint64(scramble), DATA_DIR.encode()))

    def sobol_sequence(self, px, py, first_sample, n_samples, pattern):
        pat = np.ascontiguousarray(pattern, np.int32); out = np.zeros(n_samples * int(pat.sum()), np.float32)
        check(self.L.orc_sobol_sequence(self.h, int(px), int(py), ctypes.c_uint32(first_sample), ctypes.c_uint32(n_samples), len(pat), p(pat), p(out)))
        return out.reshape(n_samples, -1)

    def set_checkerboard(self, bsdf, color0=0.4, color1=0.2, uoffset=0.0, voffset=0.0, uscale=1.0, vscale=1.0):
        c0 = f32(np.broadcast_to(color0, 3)); c1 = f32(np.broadcast_to(color1, 3))
        check(self.L.orc_bsdf_set_checkerboard(self.h, int(bsdf), p(c0), p(c1), ctypes.c_float(uoffset), ctypes.c_float(voffset), ctypes.c_float(uscale), ctypes.c_float(vscale)))

    def set_twosided(self, bsdf):
        check(self.L.orc_bsdf_set_twosided(self.h, int(bsdf)))

    def plastic_constants(self, bsdf):
        out = np.zeros(4, np.float32); check(self.L.orc_plastic_constants(self.h, int(bsdf), p(out)))
        return dict(fdrInt=out[0], fdrExt=out[1], specW=out[2], invEta2=out[3])

    def add_rectangle(self, toWorld=None, flipNormals=False, bsdf=0):
        tw = f32(IDENT if toWorld is None else toWorld).reshape(16)
        return check(self.L.orc_add_rectangle(self.h, p(tw), 1 if flipNormals else 0, int(bsdf)))

    def bsdf_eval_uv(self, bsdf, wi, wo, uv, discrete=False):
        wi = f32(wi).reshape(-1, 3); wo = f32(wo).reshape(-1, 3); uv = f32(uv).reshape(-1, 2); n = len(wi)
        ev = np.zeros((n, 3), np.float32); pdf = np.zeros(n, np.float32)
        check(self.L.orc_bsdf_eval_batch_uv(self.h, int(bsdf), ctypes.c_uint64(n), p(wi), p(wo), p(uv), 1 if discrete else 0, p(ev), p(pdf)))
        return ev, pdf

    def bsdf_sample_uv(self, bsdf, wi, sample, uv):
        wi = f32(wi).reshape(-1, 3); sample = f32(sample).reshape(-1, 2); uv = f32(uv).reshape(-1, 2); n = len(wi)
        wo = np.zeros((n, 3), np.float32); wt = np.zeros((n, 3), np.float32); pdf = np.zeros(n, np.float32); ty = np.zeros(n, np.int32)
        check(self.L.orc_bsdf_sample_batch_uv(self.h, int(bsdf), ctypes.c_uint64(n), p(wi), p(sample), p(uv), p(wo), p(wt), p(pdf), p(ty)))
        return wo, wt, pdf, ty

    def intersect_uv(self, o, d, mint, maxt):
        o = f32(o).reshape(-1, 3); d = f32(d).reshape(-1, 3); n = len(o)
        mint = f32(np.broadcast_to(mint, n)); maxt = f32(np.broadcast_to(maxt, n))
        out = np.zeros((n, 5), np.float32)
        check(self.L.orc_intersect_uv_batch(self.h, ctypes.c_uint64(n), p(o), p(d), p(mint), p(maxt), p(out)))
        return out[:, :2].copy(), out[:, 2:].copy()

    def add_hair(self, xyz, starts, radius, bsdf):
        xyz = f32(xyz).reshape(-1, 3); st = np.ascontiguousarray(starts, dtype=np.uint8)
        return check(self.L.orc_add_hair(self.h, p(xyz), p(st), ctypes.c_uint32(len(st)), ctypes.c_float(radius), int(bsdf)))

    def add_mesh(self, xyz, indices, bsdf, normals=None, uvs=None):
        xyz = f32(xyz).reshape(-1, 3); idx = np.ascontiguousarray(indices, dtype=np.uint32).reshape(-1, 3)
        nrm = None if normals is None else f32(normals).reshape(-1, 3)
        uv = None if uvs is None else f32(uvs).reshape(-1, 2)
        return check(self.L.orc_add_mesh_uv(self.h, p(xyz), None if nrm is None else p(nrm), None if uv is None else p(uv), ctypes.c_uint32(len(xyz)), p(idx),
                                            ctypes.c_uint32(len(idx)), int(bsdf)))

    def set_envmap(self, rgb, toWorld=None, scale=1.0):
        rgb = f32(rgb); h, w = rgb.shape[:2]
        tw = f32(IDENT if toWorld is None else toWorld).reshape(16)
        check(self.L.orc_set_envmap(self.h, p(rgb), w, h, p(tw), ctypes.c_float(scale)))
        self.env_w, self.env_h = w, h

    def set_camera(self, toWorld, fov=35.0, nearClip=1e-2, farClip=1e4, width=768, height=576):
        tw = f32(toWorld).reshape(16)
        check(self.L.orc_set_camera(self.h, p(tw), ctypes.c_float(fov), ctypes.c_float(nearClip), ctypes.c_float(farClip), int(width), int(height)))
        self.width, self.height = int(width), int(height)

    def set_film(self, rfilter='tent', param=0.0, has_alpha=False):
        check(self.L.orc_set_film(self.h, FILTERS[rfilter], ctypes.c_float(param), 1 if has_alpha else 0))

    def set_integrator(self, maxDepth=-1, rrDepth=5, strictNormals=False, hideEmitters=False):
        check(self.L.orc_set_integrator(self.h, int(maxDepth), int(rrDepth), 1 if strictNormals else 0, 1 if hideEmitters else 0))

    def build(self):
        check(self.L.orc_finalize(self.h))

    def scene_bounds(self):
        a = np.zeros(6, np.float32); b = np.zeros(4, np.float32)
        self.L.orc_scene_bounds(self.h, p(a), p(b))
        return a, b

    def segment_bounds(self, shape, nseg):
        out = np.zeros((nseg, 6), np.float32)
        self.L.orc_segment_bounds(self.h, int(shape), p(out))
        return out

    def bsdf_eval(self, bsdf, wi, wo, discrete=False):
        wi = f32(wi).reshape(-1, 3); wo = f32(wo).reshape(-1, 3); n = len(wi)
        ev = np.zeros((n, 3), np.float32); pdf = np.zeros(n, np.float32)
        check((self.L.orc_bsdf_eval_batch_discrete if discrete else self.L.orc_bsdf_eval_batch)(self.h, int(bsdf), ctypes.c_uint64(n), p(wi), p(wo), p(ev), p(pdf)))
        return ev, pdf

    def bsdf_sample(self, bsdf, wi, sample, extra=None):
        wi = f32(wi).reshape(-1, 3); sample = f32(sample).reshape(-1, 2); n = len(wi)
        ex = None if extra is None else f32(extra).reshape(-1, 4)
        wo = np.zeros((n, 3), np.float32); wt = np.zeros((n, 3), np.float32); pdf = np.zeros(n, np.float32); ty = np.zeros(n, np.int32)
        check(self.L.orc_bsdf_sample_batch(self.h, int(bsdf), ctypes.c_uint64(n), p(wi), p(sample), None if ex is None else p(ex), p(wo), p(wt), p(pdf), p(ty)))
        return wo, wt, pdf, ty

    def marschner_tables(self, bsdf):
        tab = np.zeros((3, 64, 64, 3), np.float32); pdf = np.zeros((3, 64, 64), np.float32); cdf = np.zeros((3, 64, 65), np.float32)
        sums = np.zeros((3, 64), np.float32); rt = np.zeros(1024, np.float32); consts = np.zeros(4, np.float32)
        check(self.L.orc_marschner_tables(self.h, int(bsdf), p(tab), p(pdf), p(cdf), p(sums), p(rt), p(consts)))
        return dict(tables=tab, pdfs=pdf, cdfs=cdf, sums=sums, rt=rt[:int(consts[3])].copy(), Fdr=float(consts[0]), specW=float(consts[1]), eta=float(consts[2]))

    def intersect(self, o, d, mint, maxt, mode=0):
        """mode: 0 BVH closest, 1 BVH any, 2 brute closest, 3 brute any"""
        o = f32(o).reshape(-1, 3); d = f32(d).reshape(-1, 3); n = len(o)
        mint = f32(np.broadcast_to(mint, n)); maxt = f32(np.broadcast_to(maxt, n))
        sh = np.zeros(n, np.int32); pr = np.zeros(n, np.uint32); t = np.zeros(n, np.float32)
        check(self.L.orc_intersect_batch(self.h, ctypes.c_uint64(n), p(o), p(d), p(mint), p(maxt), int(mode), p(sh), p(pr), p(t)))
        return sh, pr, t

    def segment_intersect(self, shape, o, d, iv, mint, maxt):
        o = f32(o).reshape(-1, 3); d = f32(d).reshape(-1, 3); n = len(o)
        iv = np.ascontiguousarray(iv, np.uint32); mint = f32(np.broadcast_to(mint, n)); maxt = f32(np.broadcast_to(maxt, n))
        hit = np.zeros(n, np.int32); t = np.zeros(n, np.float32); pt = np.zeros((n, 3), np.float32)
        check(self.L.orc_segment_intersect_batch(self.h, int(shape), ctypes.c_uint64(n), p(o), p(d), p(iv), p(mint), p(maxt), p(hit), p(t), p(pt)))
        return hit, t, pt

    def segment_records(self, shape, iv, pts):
        iv = np.ascontiguousarray(iv, np.uint32); pts = f32(pts).reshape(-1, 3); n = len(iv)
        out = np.zeros((n, 12), np.float32)
        check(self.L.orc_segment_record_batch(self.h, int(shape), ctypes.c_uint64(n), p(iv), p(pts), p(out)))
        return out

    def intersect_full(self, o, d, mint, maxt):
        o = f32(o).reshape(-1, 3); d = f32(d).reshape(-1, 3); n = len(o)
        mint = f32(np.broadcast_to(mint, n)); maxt = f32(np.broadcast_to(maxt, n))
        sh = np.zeros(n, np.int32); pr = np.zeros(n, np.uint32); t = np.zeros(n, np.float32); rec = np.zeros((n, 15), np.float32)
        check(self.L.orc_intersect_full_batch(self.h, ctypes.c_uint64(n), p(o), p(d), p(mint), p(maxt), p(sh), p(pr), p(t), p(rec)))
        return sh, pr, t, rec

    def candidates(self, o, d, mint, maxt, max_out=64):
        sh = np.zeros(max_out, np.int32); pr = np.zeros(max_out, np.uint32); t = np.zeros(max_out, np.float32)
        n = self.L.orc_intersect_candidates(self.h, p(f32(o)), p(f32(d)), ctypes.c_float(mint), ctypes.c_float(maxt), max_out, p(sh), p(pr), p(t))
        return sh[:n], pr[:n], t[:n]

    def camera_rays(self, pxy):
        pxy = f32(pxy).reshape(-1, 2); n = len(pxy)
        o = np.zeros((n, 3), np.float32); d = np.zeros((n, 3), np.float32); mm = np.zeros((n, 2), np.float32)
        self.L.orc_camera_rays(self.h, ctypes.c_uint64(n), p(pxy), p(o), p(d), p(mm))
        return o, d, mm[:, 0].copy(), mm[:, 1].copy()

    def env_eval(self, d):
        d = f32(d).reshape(-1, 3); n = len(d)
        rgb = np.zeros((n, 3), np.float32); pdf = np.zeros(n, np.float32)
        self.L.orc_env_eval_batch(self.h, ctypes.c_uint64(n), p(d), p(rgb), p(pdf))
        return rgb, pdf

    def env_eval_filtered(self, d, rx, ry):
        d = f32(d).reshape(-1, 3); rx = f32(rx).reshape(-1, 3); ry = f32(ry).reshape(-1, 3); n = len(d)
        rgb = np.zeros((n, 3), np.float32)
        self.L.orc_env_eval_filtered_batch(self.h, ctypes.c_uint64(n), p(d), p(rx), p(ry), p(rgb))
        return rgb

    def env_mip_levels(self):
        out = []
        w = ctypes.c_int(); h = ctypes.c_int()
        n = self.L.orc_env_mip_level(self.h, 0, ctypes.byref(w), ctypes.byref(h), None)
        for l in range(n):
            self.L.orc_env_mip_level(self.h, l, ctypes.byref(w), ctypes.byref(h), None)
            a = np.zeros((h.value, w.value, 3), np.float32)
            self.L.orc_env_mip_level(self.h, l, ctypes.byref(w), ctypes.byref(h), p(a))
            out.append(a)
        return out

    def env_sample(self, ref, sample):
        ref = f32(ref).reshape(-1, 3); sample = f32(sample).reshape(-1, 2); n = len(ref)
        d = np.zeros((n, 3), np.float32); v = np.zeros((n, 3), np.float32); pd = np.zeros((n, 2), np.float32)
        self.L.orc_env_sample_batch(self.h, ctypes.c_uint64(n), p(ref), p(sample), p(d), p(v), p(pd))
        return d, v, pd[:, 0].copy(), pd[:, 1].copy()

    def env_tables(self):
        w, h = self.env_w, self.env_h
        rows = np.zeros(h + 1, np.float32); cols = np.zeros((h, w + 1), np.float32); rw = np.zeros(h, np.float32); nrm = ctypes.c_float(0)
        self.L.orc_env_tables(self.h, p(rows), p(cols), p(rw), ctypes.byref(nrm), None)
        return rows, cols, rw, nrm.value

    def filter_table(self):
        t = np.zeros(32, np.float32)
        self.L.orc_filter_table(self.h, p(t))
        return t

    def render(self, spp, seed=0, sample_begin=0, sample_end=None, threads=None):
        out = np.zeros((self.height, self.width, 5), np.float32); st = np.zeros(6, np.uint64)
        check(self.L.orc_render(self.h, ctypes.c_uint32(spp), ctypes.c_uint64(seed), ctypes.c_uint32(sample_begin),
                                ctypes.c_uint32(spp if sample_end is None else sample_end), int(threads or os.cpu_count() or 1), p(out), p(st)))
        self.last_stats = dict(rays=int(st[0]), shadow_rays=int(st[1]), paths=int(st[2]), path_length=int(st[3]), dropped=int(st[4]), unsupported=int(st[5]))
        return out

    def render_samples(self, xy, samp, spp, seed=0):
        xy = np.ascontiguousarray(xy, dtype=np.uint32).reshape(-1, 2); samp = np.ascontiguousarray(samp, dtype=np.uint32); n = len(samp)
        li = np.zeros((n, 3), np.float32); pos = np.zeros((n, 2), np.float32)
        check(self.L.orc_render_samples(self.h, ctypes.c_uint64(n), p(xy), p(samp), ctypes.c_uint32(spp), ctypes.c_uint64(seed), p(li), p(pos)))
        return li, pos

    def render_samples_ref_li(self, xy, samp, spp, seed=0):
        """The reference's own MIPathTracer::Li (path.cpp compiled unmodified, oracle/_ref/libref_path.so) on this scene's components."""
        xy = np.ascontiguousarray(xy, dtype=np.uint32).reshape(-1, 2); samp = np.ascontiguousarray(samp, dtype=np.uint32); n = len(samp)
        li = np.zeros((n, 3), np.float32); alpha = np.zeros(n, np.float32); depth = np.zeros(n, np.int32)
        check(self.L.orc_render_samples_ref_li(self.h, REF_PATH_LIB.encode(), ctypes.c_uint64(n), p(xy), p(samp), ctypes.c_uint32(spp), ctypes.c_uint64(seed), p(li), p(alpha), p(depth)))
        return li, alpha, depth

    def camera_differentials(self):
        out = np.zeros(6, np.float32)
        self.L.orc_camera_differentials(self.h, p(out))
        return out

    def splat(self, pos, rgb, alpha):
        pos = f32(pos).reshape(-1, 2); rgb = f32(rgb).reshape(-1, 3); alpha = f32(alpha).reshape(-1); n = len(pos)
        out = np.zeros((self.height, self.width, 5), np.float32)
        self.L.orc_splat_batch(self.h, ctypes.c_uint64(n), p(pos), p(rgb), p(alpha), p(out))
        return out


def scene_from_description(name, scale=1.0, overrides=None, envmap=None):
    """Oracle twin of cudapath.scene_from_description (same flattened inputs)."""
    import cudapath
    scenes = cudapath.scenes
    sc = dict(scenes.SCENES[name]); sc.update(overrides or {})
    s = Scene()
    scenes.add_shapes(s, sc, scale)
    if envmap is None:
        sp = scenes.sunsky_params(name)
        envmap = bake_sunsky(sp['turbidity'], sp['albedo'][0], sp['sunDirection'], sp['skyScale'], sp['sunScale'], sp['sunRadiusScale'], sp['resolution'])
    s.set_envmap(envmap)
    s.set_camera(np.array(sc['camera'], np.float32).reshape(4, 4), sc['fov'], width=sc['width'], height=sc['height'])
    s.set_film('tent')
    s.set_integrator(maxDepth=sc['maxDepth'], rrDepth=5, strictNormals=True)
    s.build()
    return s

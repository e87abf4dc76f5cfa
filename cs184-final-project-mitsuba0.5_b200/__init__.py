"""cudapath -- B200-native hair path-tracing hot path behind Mitsuba 0.5's scene format (host-side Python mirror).

The product is the C-ABI shared library `libcudapath.so` (include/cudapath.h); this module is a thin ctypes mirror of it that
keeps the reference's plugin vocabulary (`kajiyakay`, `marschner`, `hair`, `perspective`, `path`, `sunsky`).  There is no CPU
fallback: if the CUDA library is missing the import of `lib()` fails loudly.
"""
import ctypes
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REPO = os.path.dirname(_HERE)
CLI_PATH = os.path.join(_HERE, 'cudapath_render')      # native command-line front end (csrc/cp_cli.cpp)
LIB_PATH = os.environ.get('CUDAPATH_LIB', os.path.join(_HERE, 'libcudapath.so'))   # CUDAPATH_LIB: alternative builds for tuning experiments
DEFAULT_DATA_DIR = os.environ.get('CUDAPATH_DATA_DIR', os.path.join(_REPO, 'refdata'))

_lib = None


class CudapathError(RuntimeError):
    """Mirrors the reference's Log(EError, ...) -> std::runtime_error (src/libcore/logger.cpp:100,147)."""


class Stats(ctypes.Structure):
    _fields_ = [(n, ctypes.c_uint64) for n in ('paths', 'rays', 'shadow_rays', 'kernel_launches', 'bounces', 'nodes_visited', 'prims_tested',
                                               'shadow_nodes_visited', 'shadow_prims_tested',
                                               'unsupported_filtered_lookups', 'dropped_samples', 'segments', 'bvh_nodes', 'bvh_references', 'triangles')] + \
               [(n, ctypes.c_double) for n in ('build_ms', 'render_ms', 'trace_ms', 'shade_ms', 'sort_ms', 'raygen_ms', 'splat_ms')] + \
               [(n, ctypes.c_uint64) for n in ('trace_launches', 'shade_launches', 'sort_launches', 'shadow_rays_traced', 'full_tests', 'shadow_full_tests', 'host_waits')]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


def lib():
    """Loads libcudapath.so (built by `__graft_entry__.build()` / csrc/Makefile).  Never falls back to anything else."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CudapathError('CUDA extension %s is missing -- run `python -c "import __graft_entry__ as g; g.build()"`' % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        L.cudapath_last_error.restype = ctypes.c_char_p
        L.cudapath_hair_file_radius.restype = ctypes.c_float
        L.cudapath_hair_file_vertex_count.restype = ctypes.c_uint32
        L.cudapath_mesh_file_vertex_count.restype = ctypes.c_uint32
        L.cudapath_mesh_file_triangle_count.restype = ctypes.c_uint32
        if hasattr(L, 'cudapath_last_reduce_ms'):
            L.cudapath_last_reduce_ms.restype = ctypes.c_double
        _lib = L
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _check(rc):
    if rc < 0:
        raise CudapathError(lib().cudapath_last_error().decode())
    return rc


_IDENTITY = np.eye(4, dtype=np.float32)
_DISTRIBUTIONS = {'beckmann': 0, 'ggx': 1, 'phong': 2, 'as': 2}
_FILTERS = {'tent': 0, 'box': 1, 'gaussian': 2}
_IOR = {'air': 1.000277, 'bk7': 1.5046, 'vacuum': 1.0, 'water': 1.3330, 'polypropylene': 1.49, 'amber': 1.55, 'benzene': 1.501}    # subset of src/bsdfs/ior.h


def load_hair_file(filename, radius=0.025, angleThreshold=1.0, reduction=0.0, toWorld=None):
    """HairShape(props) file loader (src/shapes/hair.cpp:609-785) -> (xyz (n,3) f32, starts_fiber (n,) u8, world radius)."""
    L = lib()
    h = ctypes.c_void_p()
    tw = _f32(_IDENTITY if toWorld is None else toWorld).reshape(16)
    _check(L.cudapath_hair_file_load(filename.encode(), ctypes.c_float(radius), ctypes.c_float(angleThreshold), ctypes.c_float(reduction), _p(tw), ctypes.byref(h)))
    n = L.cudapath_hair_file_vertex_count(h)
    xyz = np.zeros((n, 3), np.float32); starts = np.zeros(n, np.uint8)
    L.cudapath_hair_file_copy(h, _p(xyz), _p(starts))
    r = L.cudapath_hair_file_radius(h)
    L.cudapath_hair_file_free(h)
    return xyz, starts, float(r)


def load_obj_file(filename, toWorld=None, faceNormals=False, flipNormals=False, texcoords=False):
    """WavefrontOBJ(props) + TriMesh::computeNormals (src/shapes/obj.cpp:186-349, src/librender/trimesh.cpp:608-672) ->
    (xyz (n,3) f32, indices (m,3) u32, normals (n,3) f32 or None); with texcoords=True also the `vt` coordinates (n,2) or None."""
    L = lib()
    h = ctypes.c_void_p()
    tw = _f32(_IDENTITY if toWorld is None else toWorld).reshape(16)
    _check(L.cudapath_mesh_file_load(filename.encode(), _p(tw), 1 if faceNormals else 0, 1 if flipNormals else 0, ctypes.byref(h)))
    nv, nt = L.cudapath_mesh_file_vertex_count(h), L.cudapath_mesh_file_triangle_count(h)
    xyz = np.zeros((nv, 3), np.float32); idx = np.zeros((nt, 3), np.uint32)
    nrm = np.zeros((nv, 3), np.float32) if L.cudapath_mesh_file_has_normals(h) else None
    L.cudapath_mesh_file_copy(h, _p(xyz), None if nrm is None else _p(nrm), _p(idx))
    uv = None
    if texcoords and L.cudapath_mesh_file_has_texcoords(h):
        uv = np.zeros((nv, 2), np.float32); L.cudapath_mesh_file_copy_texcoords(h, _p(uv))
    L.cudapath_mesh_file_free(h)
    return (xyz, idx, nrm, uv) if texcoords else (xyz, idx, nrm)


def fresnel_diffuse_reflectance(eta):
    """fresnelDiffuseReflectance(eta, fast=false), src/libcore/util.cpp:814-862 (host only)."""
    out = ctypes.c_float()
    _check(lib().cudapath_fresnel_diffuse_reflectance(ctypes.c_float(eta), ctypes.byref(out)))
    return np.float32(out.value)


def validate_scene_xml(filename, defines=None):
    """Dry run of the scene loader (no GPU): returns the list of objects the file would create; raises CudapathError naming the
    first unsupported plugin / parameter."""
    buf = ctypes.create_string_buffer(1 << 16)
    d = ';'.join('%s=%s' % kv for kv in (defines or {}).items())
    _check(lib().cudapath_validate_scene_xml(filename.encode(), d.encode(), buf, ctypes.c_size_t(len(buf))))
    return buf.value.decode().splitlines()


def bake_sunsky(turbidity=3.0, albedo=(0.2, 0.2, 0.2), sunDirection=(0, 1, 0), skyScale=1.0, sunScale=1.0, sunRadiusScale=1.0, resolution=512,
                data_dir=None):
    """SunSkyEmitter bake (src/emitters/sunsky.cpp:100-216) -> (resolution/2, resolution, 3) fp32 lat-long map."""
    out = np.zeros((resolution // 2, resolution, 3), np.float32)
    _check(lib().cudapath_bake_sunsky((data_dir or DEFAULT_DATA_DIR).encode(), ctypes.c_float(turbidity), _p(_f32(albedo)), _p(_f32(sunDirection)),
                                      ctypes.c_float(skyScale), ctypes.c_float(sunScale), ctypes.c_float(sunRadiusScale), int(resolution), _p(out)))
    return out


def env_pyramid(rgb):
    """Host-only: the Lanczos MIP pyramid of a lat-long fp32 bitmap (cudapath_env_pyramid_level), fp32 levels before half quantisation."""
    rgb = _f32(rgb); h, w = rgb.shape[:2]
    out = []
    lw = ctypes.c_int(); lh = ctypes.c_int()
    n = _check(lib().cudapath_env_pyramid_level(_p(rgb), w, h, 0, ctypes.byref(lw), ctypes.byref(lh), None))
    for l in range(n):
        _check(lib().cudapath_env_pyramid_level(_p(rgb), w, h, l, ctypes.byref(lw), ctypes.byref(lh), None))
        a = np.zeros((lh.value, lw.value, 3), np.float32)
        _check(lib().cudapath_env_pyramid_level(_p(rgb), w, h, l, ctypes.byref(lw), ctypes.byref(lh), _p(a)))
        out.append(a)
    return out


def sun_radiance(turbidity=3.0, sunDirection=(0, 1, 0), data_dir=None):
    """computeSunRadiance (src/emitters/sunsky/sunmodel.h:260-371) in linear RGB -> (3,) fp32."""
    out = np.zeros(3, np.float32)
    _check(lib().cudapath_sun_radiance((data_dir or DEFAULT_DATA_DIR).encode(), ctypes.c_float(turbidity), _p(_f32(sunDirection)), _p(out)))
    return out


def develop(film):
    """Film::develop normalisation (src/libcore/fmtconv.cpp:955-1056): (h,w,5) accumulated film -> (h,w,3) RGB."""
    film = _f32(film)
    h, w = film.shape[:2]
    out = np.zeros((h, w, 3), np.float32)
    _check(lib().cudapath_develop(_p(film), w, h, _p(out)))
    return out


def develop_ldr(film, gamma=-1.0, exposure=0.0):
    """LDRFilm::develop, `gamma` tonemapper (src/films/ldrfilm.cpp:300-321): (h,w,5) accumulated film -> (h,w,3) uint8."""
    film = _f32(film)
    h, w = film.shape[:2]
    out = np.zeros((h, w, 3), np.uint8)
    _check(lib().cudapath_develop_ldr(_p(film), w, h, ctypes.c_float(gamma), ctypes.c_float(exposure), _p(out)))
    return out


class Context:
    """One GPU context = the flattened scene + the wavefront path integrator (replaces Scene + `path` integrator for this path)."""

    def __init__(self, device=0, data_dir=None):
        """device: a CUDA device index, or a list of indices for ONE context spanning several GPUs of the box (cudapath_create_multi:
        the scene is replicated, render() splits the sample range and sums the films with one ncclReduce onto the first device)."""
        self._h = ctypes.c_void_p()
        self._L = lib()
        if isinstance(device, (list, tuple)):
            devs = (ctypes.c_int * len(device))(*[int(d) for d in device])
            _check(self._L.cudapath_create_multi(devs, len(device), ctypes.byref(self._h)))
        else:
            _check(self._L.cudapath_create(int(device), ctypes.byref(self._h)))
        _check(self._L.cudapath_set_data_dir(self._h, (data_dir or DEFAULT_DATA_DIR).encode()))
        self.width = self.height = 0
        self.spp = 0

    def close(self):
        if self._h:
            self._L.cudapath_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- plugins -------------------------------------------------------------------------------------------------
    def add_bsdf(self, type, **props):
        """`<bsdf type=...>`: kajiyakay (src/bsdfs/kajiyakay.cpp:60-73) or marschner (src/bsdfs/marschner_diffuse.cpp:113-160)."""
        if type == 'kajiyakay':
            d = _f32(np.broadcast_to(props.get('diffuseReflectance', 0.5), 3)); s = _f32(np.broadcast_to(props.get('specularReflectance', 0.2), 3))
            return _check(self._L.cudapath_add_bsdf_kajiyakay(self._h, _p(d), _p(s), ctypes.c_float(props.get('exponent', 30.0))))
        if type == 'marschner':
            ior = lambda v: float(_IOR[v.lower()]) if isinstance(v, str) else float(v)
            d = _f32(np.broadcast_to(props.get('diffuseReflectance', 0.5), 3)); s = _f32(np.broadcast_to(props.get('specularReflectance', 0.5), 3))
            distr = props.get('distribution', 'beckmann').lower()
            if distr not in _DISTRIBUTIONS:
                raise CudapathError('Specified an invalid distribution "%s", must be "beckmann", "ggx", or "phong"/"as"!' % distr)
            return _check(self._L.cudapath_add_bsdf_marschner(self._h, ctypes.c_float(ior(props.get('intIOR', 'bk7'))), ctypes.c_float(ior(props.get('extIOR', 'air'))),
                                                              _p(d), _p(s), ctypes.c_float(props.get('alpha', 0.1)), _DISTRIBUTIONS[distr],
                                                              1 if props.get('nonlinear', False) else 0))
        if type == 'roughplastic':
            ior = lambda v: float(_IOR[v.lower()]) if isinstance(v, str) else float(v)
            d = _f32(np.broadcast_to(props.get('diffuseReflectance', 0.5), 3)); s = _f32(np.broadcast_to(props.get('specularReflectance', 1.0), 3))
            distr = props.get('distribution', 'beckmann').lower()
            if distr not in _DISTRIBUTIONS:
                raise CudapathError('Specified an invalid distribution "%s", must be "beckmann", "ggx", or "phong"/"as"!' % distr)
            return _check(self._L.cudapath_add_bsdf_roughplastic(self._h, ctypes.c_float(ior(props.get('intIOR', 'polypropylene'))), ctypes.c_float(ior(props.get('extIOR', 'air'))),
                                                                 _p(d), _p(s), ctypes.c_float(props.get('alpha', 0.1)), _DISTRIBUTIONS[distr],
                                                                 1 if props.get('sampleVisible', True) else 0, 1 if props.get('nonlinear', False) else 0))
        if type == 'marschner_fixed':
            # the fork's unbuilt src/bsdfs/marschner.cpp ("fixed" mode: TRT-only eval, real pdf); defaults amber / air
            ior = lambda v: float(_IOR[v.lower()]) if isinstance(v, str) else float(v)
            return _check(self._L.cudapath_add_bsdf_marschner_fixed(self._h, ctypes.c_float(ior(props.get('intIOR', 'amber'))), ctypes.c_float(ior(props.get('extIOR', 'air')))))
        if type == 'marschner_full':
            # the same class with all three lobes in eval() and scene-driven constants (cudapath_add_bsdf_marschner_full)
            ior = lambda v: float(_IOR[v.lower()]) if isinstance(v, str) else float(v)
            sa = _f32(np.broadcast_to(props.get('sigmaA', 0.22), 3))
            return _check(self._L.cudapath_add_bsdf_marschner_full(self._h, ctypes.c_float(ior(props.get('intIOR', 'amber'))), ctypes.c_float(ior(props.get('extIOR', 'air'))), _p(sa),
                                                                   ctypes.c_float(props.get('betaR', 0.1)), ctypes.c_float(props.get('scaleAngleRad', -0.1)), int(props.get('lobes', 7))))
        if type == 'mirror':
            # the fork's src/bsdfs/mirror.cpp: default specularReflectance 1
            return _check(self._L.cudapath_add_bsdf_mirror(self._h, _p(_f32(np.broadcast_to(props.get('specularReflectance', 1.0), 3)))))
        if type == 'plastic':
            # src/bsdfs/plastic.cpp:140-167; defaults polypropylene / air, specular 1, diffuse 0.5
            ior = lambda v: float(_IOR[v.lower()]) if isinstance(v, str) else float(v)
            d = _f32(np.broadcast_to(props.get('diffuseReflectance', 0.5), 3)); s = _f32(np.broadcast_to(props.get('specularReflectance', 1.0), 3))
            return _check(self._L.cudapath_add_bsdf_plastic(self._h, ctypes.c_float(ior(props.get('intIOR', 'polypropylene'))), ctypes.c_float(ior(props.get('extIOR', 'air'))),
                                                            _p(d), _p(s), 1 if props.get('nonlinear', False) else 0))
        if type in ('diffuse', 'twosided'):
            # `diffuse` with a constant reflectance (src/bsdfs/diffuse.cpp:70-103); type 'twosided' = <bsdf type="twosided"><bsdf type="diffuse"/></bsdf>
            r = _f32(np.broadcast_to(props.get('reflectance', 0.5), 3))
            return _check(self._L.cudapath_add_bsdf_diffuse(self._h, _p(r), 1 if (type == 'twosided' or props.get('twoSided', False)) else 0))
        if type == 'thindielectric':
            # src/bsdfs/thindielectric.cpp:73-91; defaults bk7 / air, both colours 1
            ior = lambda v: float(_IOR[v.lower()]) if isinstance(v, str) else float(v)
            r = _f32(np.broadcast_to(props.get('specularReflectance', 1.0), 3)); t = _f32(np.broadcast_to(props.get('specularTransmittance', 1.0), 3))
            return _check(self._L.cudapath_add_bsdf_thindielectric(self._h, ctypes.c_float(ior(props.get('intIOR', 'bk7'))), ctypes.c_float(ior(props.get('extIOR', 'air'))), _p(r), _p(t)))
        if type == 'marschnerdielectric':
            # the fork's src/bsdfs/marschnerdielectric.cpp:128-167; defaults benzene / air, diffuse 0.5, specular 0.1 / 0.1, exponent 30
            ior = lambda v: float(_IOR[v.lower()]) if isinstance(v, str) else float(v)
            d = _f32(np.broadcast_to(props.get('diffuseReflectance', 0.5), 3))
            r = _f32(np.broadcast_to(props.get('specularReflectance', 0.1), 3)); t = _f32(np.broadcast_to(props.get('specularTransmittance', 0.1), 3))
            return _check(self._L.cudapath_add_bsdf_marschnerdielectric(self._h, ctypes.c_float(ior(props.get('intIOR', 'benzene'))), ctypes.c_float(ior(props.get('extIOR', 'air'))),
                                                                        _p(d), _p(r), _p(t), ctypes.c_float(props.get('exponent', 30.0))))
        raise CudapathError('bsdf plugin "%s" is outside the hair hot path (supported: kajiyakay, marschner, marschner_fixed, marschnerdielectric, thindielectric, roughplastic, plastic, mirror, diffuse, twosided)' % type)

    def set_checkerboard(self, bsdf_id, color0=0.4, color1=0.2, uoffset=0.0, voffset=0.0, uscale=1.0, vscale=1.0):
        """`<texture type="checkerboard">` as the reflectance of a `diffuse` / the diffuseReflectance of a `plastic` BSDF (src/textures/checkerboard.cpp)."""
        c0 = _f32(np.broadcast_to(color0, 3)); c1 = _f32(np.broadcast_to(color1, 3))
        _check(self._L.cudapath_bsdf_set_checkerboard(self._h, int(bsdf_id), _p(c0), _p(c1), ctypes.c_float(uoffset), ctypes.c_float(voffset), ctypes.c_float(uscale), ctypes.c_float(vscale)))

    def set_twosided(self, bsdf_id):
        """`<bsdf type="twosided">` around a diffuse / roughplastic / plastic BSDF (src/bsdfs/twosided.cpp)."""
        _check(self._L.cudapath_bsdf_set_twosided(self._h, int(bsdf_id)))

    def add_rectangle(self, bsdf_id, toWorld=None, flipNormals=False):
        """`<shape type="rectangle">` (src/shapes/rectangle.cpp): [-1,1]^2 x {0} under toWorld."""
        tw = _f32(_IDENTITY if toWorld is None else toWorld).reshape(16)
        return _check(self._L.cudapath_add_rectangle(self._h, _p(tw), 1 if flipNormals else 0, int(bsdf_id)))

    def add_mesh(self, xyz, indices, bsdf_id, normals=None, uvs=None):
        """Triangle mesh (TriMesh positions / optional vertex normals / optional texture coordinates / index triples); joins the fibers in the device BVH."""
        xyz = _f32(xyz).reshape(-1, 3); idx = np.ascontiguousarray(indices, dtype=np.uint32).reshape(-1, 3)
        nrm = None if normals is None else _f32(normals).reshape(-1, 3)
        uv = None if uvs is None else _f32(uvs).reshape(-1, 2)
        if nrm is not None and len(nrm) != len(xyz):
            raise CudapathError('normals must have one entry per vertex')
        if uv is not None and len(uv) != len(xyz):
            raise CudapathError('uvs must have one entry per vertex')
        return _check(self._L.cudapath_add_mesh_uv(self._h, _p(xyz), None if nrm is None else _p(nrm), None if uv is None else _p(uv), ctypes.c_uint32(len(xyz)), _p(idx),
                                                   ctypes.c_uint32(len(idx)), int(bsdf_id)))

    def add_hair(self, xyz, starts_fiber, radius, bsdf_id):
        xyz = _f32(xyz).reshape(-1, 3); st = np.ascontiguousarray(starts_fiber, dtype=np.uint8)
        if len(st) != len(xyz):
            raise CudapathError('starts_fiber must have one entry per vertex')
        return _check(self._L.cudapath_add_hair(self._h, _p(xyz), _p(st), ctypes.c_uint32(len(st)), ctypes.c_float(radius), int(bsdf_id)))

    def add_mesh_file(self, filename, bsdf_id, toWorld=None, faceNormals=False, flipNormals=False):
        tw = _f32(_IDENTITY if toWorld is None else toWorld).reshape(16)
        return _check(self._L.cudapath_add_mesh_file(self._h, filename.encode(), _p(tw), 1 if faceNormals else 0, 1 if flipNormals else 0, int(bsdf_id)))

    def add_hair_file(self, filename, bsdf_id, radius=0.025, angleThreshold=1.0, reduction=0.0, toWorld=None):
        tw = _f32(_IDENTITY if toWorld is None else toWorld).reshape(16)
        return _check(self._L.cudapath_add_hair_file(self._h, filename.encode(), ctypes.c_float(radius), ctypes.c_float(angleThreshold), ctypes.c_float(reduction), _p(tw), int(bsdf_id)))

    def set_envmap(self, rgb, toWorld=None, scale=1.0):
        rgb = _f32(rgb); h, w = rgb.shape[:2]
        tw = _f32(_IDENTITY if toWorld is None else toWorld).reshape(16)
        _check(self._L.cudapath_set_envmap(self._h, _p(rgb), w, h, _p(tw), ctypes.c_float(scale)))

    def set_sunsky(self, turbidity=3.0, albedo=(0.2, 0.2, 0.2), sunDirection=(0, 1, 0), skyScale=1.0, sunScale=1.0, sunRadiusScale=1.0, resolution=512):
        _check(self._L.cudapath_set_sunsky(self._h, ctypes.c_float(turbidity), _p(_f32(albedo)), _p(_f32(sunDirection)), ctypes.c_float(skyScale),
                                           ctypes.c_float(sunScale), ctypes.c_float(sunRadiusScale), int(resolution)))

    def set_camera(self, toWorld, fov=35.0, nearClip=1e-2, farClip=1e4, width=768, height=576):
        tw = _f32(toWorld).reshape(16)
        _check(self._L.cudapath_set_camera_perspective(self._h, _p(tw), ctypes.c_float(fov), ctypes.c_float(nearClip), ctypes.c_float(farClip), int(width), int(height)))
        self.width, self.height = int(width), int(height)

    def set_film(self, rfilter='tent', param=0.0, has_alpha=False):
        _check(self._L.cudapath_set_film(self._h, _FILTERS[rfilter], ctypes.c_float(param), 1 if has_alpha else 0))

    def film_output(self):
        """(hdr, gamma, exposure) the scene's film asks for (ldrfilm.cpp:180-181); set by load_xml, defaults False, -1 (sRGB), 0."""
        hdr = ctypes.c_int(); g = ctypes.c_float(); e = ctypes.c_float()
        _check(self._L.cudapath_get_film_output(self._h, ctypes.byref(hdr), ctypes.byref(g), ctypes.byref(e)))
        return bool(hdr.value), float(g.value), float(e.value)

    def set_envmap_file(self, path, toWorld=None, scale=1.0):
        """`<emitter type="envmap"><string name="filename" value="...hdr"/>`: src/emitters/envmap.cpp:100-190 with a Radiance RGBE file."""
        tw = _f32(np.eye(4) if toWorld is None else toWorld).reshape(16)
        _check(self._L.cudapath_set_envmap_file(self._h, str(path).encode(), _p(tw), ctypes.c_float(scale)))

    def set_integrator(self, maxDepth=-1, rrDepth=5, strictNormals=False, hideEmitters=False):
        _check(self._L.cudapath_set_integrator(self._h, int(maxDepth), int(rrDepth), 1 if strictNormals else 0, 1 if hideEmitters else 0))

    def set_sampler(self, kind='philox', scramble=0):
        """'philox' (default: counter-based stream), 'sobol' (the reference's src/samplers/sobol.cpp reproduced number for number) or 'file'
        (the <sampler> of the scene file loaded next: sobol is honoured, anything else refused)."""
        _check(self._L.cudapath_set_sampler(self._h, {'philox': 0, 'sobol': 1, 'file': 2}[kind], ctypes.c_uint64(scramble)))

    def set_options(self, wave_size=0, collect_stats=False, profile_stages=False):
        _check(self._L.cudapath_set_options(self._h, ctypes.c_uint32(wave_size), 1 if collect_stats else 0, 1 if profile_stages else 0))

    def set_pixel_shard(self, index=0, count=1):
        """Render only the pixel blocks owned by shard `index` of `count` (the films of all shards add up to the image; block side 8 / 16 / 32
        pixels, CUDAPATH_SHARD_BLOCK)."""
        _check(self._L.cudapath_set_pixel_shard(self._h, ctypes.c_uint32(index), ctypes.c_uint32(count)))

    def set_math_mode(self, mode):
        """'strict' (every elementary function correctly rounded: bit-identical to the oracle) or 'fast' (default; see cudapath.h)."""
        if mode not in ('strict', 'fast'):
            raise CudapathError("math mode must be 'strict' or 'fast'")
        _check(self._L.cudapath_set_math_mode(self._h, 1 if mode == 'strict' else 0))

    def math_mode(self):
        return 'strict' if self._L.cudapath_get_math_mode(self._h) == 1 else 'fast'

    def set_build_options(self, max_split=16):
        _check(self._L.cudapath_set_build_options(self._h, int(max_split)))

    def set_job_size_hint(self, paths_per_device):
        """width x height x spp / devices of the job this scene is built for: picks the build effort (see cudapath.h); call before build()."""
        _check(self._L.cudapath_set_job_size_hint(self._h, ctypes.c_uint64(int(paths_per_device))))

    def load_xml(self, filename, defines=None):
        """SceneHandler for the hair scenes; `defines` = dict for $name substitution (mitsuba -D).  Returns sampleCount."""
        spp = ctypes.c_uint32(0)
        d = ';'.join('%s=%s' % kv for kv in (defines or {}).items())
        _check(self._L.cudapath_load_scene_xml(self._h, filename.encode(), d.encode(), ctypes.byref(spp)))
        self.spp = spp.value
        return spp.value

    def build(self):
        _check(self._L.cudapath_build(self._h))
        st = self.stats()
        return st

    # ---- render ----------------------------------------------------------------------------------------------------
    def film_shape(self):
        w = ctypes.c_int(0); h = ctypes.c_int(0)
        _check(self._L.cudapath_film_size(self._h, ctypes.byref(w), ctypes.byref(h)))
        self.width, self.height = w.value, h.value
        return (self.height, self.width, 5)

    def render(self, spp, seed=0, sample_begin=0, sample_end=None, out=None):
        """Accumulated film (h,w,5) for sample indices [sample_begin, sample_end) of spp; host buffer in/out (the e2e path).  `out`: a
        caller-owned C-contiguous float32 array of the film's shape to receive it (e.g. the numpy view of a page-locked torch tensor:
        the read-back then runs at PCIe speed instead of through pageable memory); it is overwritten, not accumulated into."""
        if out is None:
            out = np.zeros(self.film_shape(), np.float32)
        elif out.dtype != np.float32 or tuple(out.shape) != tuple(self.film_shape()) or not out.flags['C_CONTIGUOUS']:
            raise CudapathError('render(out=...): a C-contiguous float32 array of shape %s is required' % (tuple(self.film_shape()),))
        _check(self._L.cudapath_render(self._h, ctypes.c_uint32(spp), ctypes.c_uint64(seed), ctypes.c_uint32(sample_begin),
                                       ctypes.c_uint32(spp if sample_end is None else sample_end), _p(out)))
        return out

    def render_into(self, film_dev_ptr, spp, seed=0, sample_begin=0, sample_end=None, stream=0):
        """Accumulates into a caller-owned DEVICE film (e.g. a torch tensor's data_ptr()); used for the NCCL film reduce."""
        _check(self._L.cudapath_render_dev(self._h, ctypes.c_uint32(spp), ctypes.c_uint64(seed), ctypes.c_uint32(sample_begin),
                                           ctypes.c_uint32(spp if sample_end is None else sample_end), ctypes.c_void_p(film_dev_ptr), ctypes.c_void_p(stream)))

    def device_count(self):
        return int(self._L.cudapath_device_count(self._h))

    def last_reduce_ms(self):
        """Device time of the ncclReduce of the last multi-GPU render (0 for one device)."""
        return float(self._L.cudapath_last_reduce_ms(self._h))

    def measure_read_bandwidth(self, nbytes, iterations):
        """GB/s of 16-byte L1-bypassing loads over a resident buffer (L2 bandwidth for a buffer well below the L2 size, HBM for GBs)."""
        out = ctypes.c_double(0)
        _check(self._L.cudapath_measure_read_bandwidth(self._h, ctypes.c_size_t(nbytes), int(iterations), ctypes.byref(out)))
        return out.value

    def cancel(self):
        """Integrator::cancel(): may be called from another thread while render() / render_into() blocks; that call then raises
        CudapathError('render cancelled')."""
        _check(self._L.cudapath_cancel(self._h))

    def set_progress_callback(self, fn):
        """fn(paths_done, paths_total) after every finished wave of a render (None removes it)."""
        if fn is None:
            self._progress_cb = None
            _check(self._L.cudapath_set_progress_callback(self._h, None, None))
            return
        CB = ctypes.CFUNCTYPE(None, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint64)
        self._progress_cb = CB(lambda user, done, total: fn(int(done), int(total)))     # kept alive with the context
        _check(self._L.cudapath_set_progress_callback(self._h, self._progress_cb, None))

    def stats(self):
        s = Stats()
        _check(self._L.cudapath_get_stats(self._h, ctypes.byref(s)))
        return s.as_dict()

    def scene_bounds(self):
        a = np.zeros(6, np.float32); b = np.zeros(4, np.float32)
        _check(self._L.cudapath_scene_bounds(self._h, _p(a), _p(b)))
        return a, b

    # ---- parity hooks (host buffers) --------------------------------------------------------------------------------
    def bsdf_eval(self, bsdf_id, wi, wo, discrete=False):
        """BSDF::eval + BSDF::pdf; measure ESolidAngle, or EDiscrete with discrete=True (delta components of `thindielectric`)."""
        wi = _f32(wi).reshape(-1, 3); wo = _f32(wo).reshape(-1, 3); n = len(wi)
        ev = np.zeros((n, 3), np.float32); pdf = np.zeros(n, np.float32)
        _check((self._L.cudapath_bsdf_eval_batch_discrete if discrete else self._L.cudapath_bsdf_eval_batch)(self._h, int(bsdf_id), ctypes.c_uint64(n), _p(wi), _p(wo), _p(ev), _p(pdf)))
        return ev, pdf

    def bsdf_eval_uv(self, bsdf_id, wi, wo, uv, discrete=False):
        """BSDF::eval + pdf with texture coordinates per tuple (its.uv): the textured kinds."""
        wi = _f32(wi).reshape(-1, 3); wo = _f32(wo).reshape(-1, 3); uv = _f32(uv).reshape(-1, 2); n = len(wi)
        ev = np.zeros((n, 3), np.float32); pdf = np.zeros(n, np.float32)
        _check(self._L.cudapath_bsdf_eval_batch_uv(self._h, int(bsdf_id), ctypes.c_uint64(n), _p(wi), _p(wo), _p(uv), 1 if discrete else 0, _p(ev), _p(pdf)))
        return ev, pdf

    def bsdf_sample_uv(self, bsdf_id, wi, sample, uv):
        wi = _f32(wi).reshape(-1, 3); sample = _f32(sample).reshape(-1, 2); uv = _f32(uv).reshape(-1, 2); n = len(wi)
        wo = np.zeros((n, 3), np.float32); wt = np.zeros((n, 3), np.float32); pdf = np.zeros(n, np.float32); ty = np.zeros(n, np.int32)
        _check(self._L.cudapath_bsdf_sample_batch_uv(self._h, int(bsdf_id), ctypes.c_uint64(n), _p(wi), _p(sample), _p(uv), _p(wo), _p(wt), _p(pdf), _p(ty)))
        return wo, wt, pdf, ty

    def intersect_uv(self, o, d, mint, maxt):
        """Closest hit with the record, its.uv and the geometric normal: (shape, prim, t, record (n,15), uv (n,2), geoN (n,3))."""
        o = _f32(o).reshape(-1, 3); d = _f32(d).reshape(-1, 3); n = len(o)
        mint = _f32(np.broadcast_to(mint, n)); maxt = _f32(np.broadcast_to(maxt, n))
        sh = np.zeros(n, np.int32); pr = np.zeros(n, np.uint32); t = np.zeros(n, np.float32); rec = np.zeros((n, 15), np.float32); uv = np.zeros((n, 5), np.float32)
        _check(self._L.cudapath_intersect_batch_uv(self._h, ctypes.c_uint64(n), _p(o), _p(d), _p(mint), _p(maxt), _p(sh), _p(pr), _p(t), _p(rec), _p(uv)))
        return sh, pr, t, rec, uv[:, :2].copy(), uv[:, 2:].copy()

    def bsdf_eval_world(self, bsdf_id, frames, wi_world, wo_world):
        """BSDF::eval + pdf from world-space directions and per-tuple shading frames (n, 3, 3) = rows s, t, n (Frame::toLocal on the device)."""
        fr = _f32(frames).reshape(-1, 9); wi = _f32(wi_world).reshape(-1, 3); wo = _f32(wo_world).reshape(-1, 3); n = len(wi)
        ev = np.zeros((n, 3), np.float32); pdf = np.zeros(n, np.float32)
        _check(self._L.cudapath_bsdf_eval_batch_world(self._h, int(bsdf_id), ctypes.c_uint64(n), _p(fr), _p(wi), _p(wo), _p(ev), _p(pdf)))
        return ev, pdf

    def bsdf_sample(self, bsdf_id, wi, sample, extra=None):
        wi = _f32(wi).reshape(-1, 3); sample = _f32(sample).reshape(-1, 2); n = len(wi)
        ex = None if extra is None else _f32(extra).reshape(-1, 4)
        wo = np.zeros((n, 3), np.float32); wt = np.zeros((n, 3), np.float32); pdf = np.zeros(n, np.float32); ty = np.zeros(n, np.int32)
        _check(self._L.cudapath_bsdf_sample_batch_ex(self._h, int(bsdf_id), ctypes.c_uint64(n), _p(wi), _p(sample), None if ex is None else _p(ex), _p(wo), _p(wt), _p(pdf), _p(ty)))
        return wo, wt, pdf, ty

    def intersect(self, o, d, mint, maxt, any_hit=False, record=False):
        o = _f32(o).reshape(-1, 3); d = _f32(d).reshape(-1, 3); n = len(o)
        mint = _f32(np.broadcast_to(mint, n)); maxt = _f32(np.broadcast_to(maxt, n))
        sh = np.zeros(n, np.int32); pr = np.zeros(n, np.uint32); t = np.zeros(n, np.float32)
        rec = np.zeros((n, 15), np.float32) if record else None
        _check(self._L.cudapath_intersect_batch(self._h, ctypes.c_uint64(n), _p(o), _p(d), _p(mint), _p(maxt), 1 if any_hit else 0, _p(sh), _p(pr), _p(t),
                                                _p(rec) if record else None))
        return (sh, pr, t, rec) if record else (sh, pr, t)

    def env_eval(self, d):
        d = _f32(d).reshape(-1, 3); n = len(d)
        rgb = np.zeros((n, 3), np.float32); pdf = np.zeros(n, np.float32)
        _check(self._L.cudapath_env_eval_batch(self._h, ctypes.c_uint64(n), _p(d), _p(rgb), _p(pdf)))
        return rgb, pdf

    def env_eval_filtered(self, d, rx, ry):
        """evalEnvironment of a ray with differentials: EWA lookup in the Lanczos MIP pyramid (envmap.cpp:391-407, mipmap.h:629-836)."""
        d = _f32(d).reshape(-1, 3); rx = _f32(rx).reshape(-1, 3); ry = _f32(ry).reshape(-1, 3); n = len(d)
        rgb = np.zeros((n, 3), np.float32)
        _check(self._L.cudapath_env_eval_filtered_batch(self._h, ctypes.c_uint64(n), _p(d), _p(rx), _p(ry), _p(rgb)))
        return rgb

    def env_mip_levels(self):
        """The MIP pyramid of the environment map as stored on the device: list of (h, w, 3) arrays, level 0 first."""
        out = []
        w = ctypes.c_int(); h = ctypes.c_int()
        n = _check(self._L.cudapath_env_mip_level(self._h, 0, ctypes.byref(w), ctypes.byref(h), None))
        for l in range(n):
            _check(self._L.cudapath_env_mip_level(self._h, l, ctypes.byref(w), ctypes.byref(h), None))
            a = np.zeros((h.value, w.value, 3), np.float32)
            _check(self._L.cudapath_env_mip_level(self._h, l, ctypes.byref(w), ctypes.byref(h), _p(a)))
            out.append(a)
        return out

    def env_sample(self, ref, sample):
        ref = _f32(ref).reshape(-1, 3); sample = _f32(sample).reshape(-1, 2); n = len(ref)
        d = np.zeros((n, 3), np.float32); v = np.zeros((n, 3), np.float32); pd = np.zeros((n, 2), np.float32)
        _check(self._L.cudapath_env_sample_batch(self._h, ctypes.c_uint64(n), _p(ref), _p(sample), _p(d), _p(v), _p(pd)))
        return d, v, pd[:, 0].copy(), pd[:, 1].copy()

    def camera_rays(self, pixel_sample):
        ps = _f32(pixel_sample).reshape(-1, 2); n = len(ps)
        o = np.zeros((n, 3), np.float32); d = np.zeros((n, 3), np.float32); mm = np.zeros((n, 2), np.float32)
        _check(self._L.cudapath_camera_rays_batch(self._h, ctypes.c_uint64(n), _p(ps), _p(o), _p(d), _p(mm)))
        return o, d, mm[:, 0].copy(), mm[:, 1].copy()

    def splat(self, pos, rgb, alpha):
        pos = _f32(pos).reshape(-1, 2); rgb = _f32(rgb).reshape(-1, 3); alpha = _f32(alpha).reshape(-1); n = len(pos)
        out = np.zeros(self.film_shape(), np.float32)
        _check(self._L.cudapath_splat_batch(self._h, ctypes.c_uint64(n), _p(pos), _p(rgb), _p(alpha), _p(out)))
        return out

    def marschner_tables(self, bsdf_id):
        tab = np.zeros((3, 64, 64, 3), np.float32); pdf = np.zeros((3, 64, 64), np.float32); cdf = np.zeros((3, 64, 65), np.float32)
        sums = np.zeros((3, 64), np.float32); rt = np.zeros(1024, np.float32); consts = np.zeros(4, np.float32)
        _check(self._L.cudapath_marschner_tables(self._h, int(bsdf_id), _p(tab), _p(pdf), _p(cdf), _p(sums), _p(rt), _p(consts)))
        return dict(tables=tab, pdfs=pdf, cdfs=cdf, sums=sums, rt=rt[:int(consts[3])].copy(), Fdr=float(consts[0]), specW=float(consts[1]), eta=float(consts[2]))

    def env_tables(self, w, h):
        rows = np.zeros(h + 1, np.float32); cols = np.zeros((h, w + 1), np.float32); rw = np.zeros(h, np.float32); nrm = ctypes.c_float(0)
        _check(self._L.cudapath_env_tables(self._h, _p(rows), _p(cols), _p(rw), ctypes.byref(nrm)))
        return rows, cols, rw, nrm.value

    def filter_table(self):
        t = np.zeros(32, np.float32)
        _check(self._L.cudapath_filter_table(self._h, _p(t)))
        return t

    # ---- device-resident stage benchmarks (pointers are raw device addresses, e.g. torch tensors' data_ptr()) ------
    def bsdf_eval_dev(self, bsdf_id, n, wi, wo, out_eval, out_pdf, stream=0):
        _check(self._L.cudapath_bsdf_eval_batch_dev(self._h, int(bsdf_id), ctypes.c_uint64(n), ctypes.c_void_p(wi), ctypes.c_void_p(wo),
                                                    ctypes.c_void_p(out_eval), ctypes.c_void_p(out_pdf), ctypes.c_void_p(stream)))

    def bsdf_sample_dev(self, bsdf_id, n, wi, sample, out_wo, out_weight, out_pdf, out_type, stream=0):
        _check(self._L.cudapath_bsdf_sample_batch_dev(self._h, int(bsdf_id), ctypes.c_uint64(n), ctypes.c_void_p(wi), ctypes.c_void_p(sample), ctypes.c_void_p(out_wo),
                                                      ctypes.c_void_p(out_weight), ctypes.c_void_p(out_pdf), ctypes.c_void_p(out_type), ctypes.c_void_p(stream)))

    def intersect_dev(self, n, o, d, mint, maxt, out_shape, out_prim, out_t, any_hit=False, out_stats=0, stream=0):
        _check(self._L.cudapath_intersect_batch_dev(self._h, ctypes.c_uint64(n), ctypes.c_void_p(o), ctypes.c_void_p(d), ctypes.c_void_p(mint), ctypes.c_void_p(maxt),
                                                    1 if any_hit else 0, ctypes.c_void_p(out_shape), ctypes.c_void_p(out_prim), ctypes.c_void_p(out_t),
                                                    ctypes.c_void_p(out_stats) if out_stats else None, ctypes.c_void_p(stream)))


def load_rgbe(path):
    """Radiance RGBE (.hdr) image as a (h, w, 3) float32 array (Bitmap::readRGBE, src/libcore/bitmap.cpp:3590-3678)."""
    w = ctypes.c_int(); h = ctypes.c_int()
    _check(lib().cudapath_load_rgbe(str(path).encode(), None, ctypes.byref(w), ctypes.byref(h)))
    out = np.zeros((h.value, w.value, 3), np.float32)
    _check(lib().cudapath_load_rgbe(str(path).encode(), _p(out), ctypes.byref(w), ctypes.byref(h)))
    return out


def visible_devices():
    """CUDA devices visible to this process (0 without a driver)."""
    return int(lib().cudapath_visible_devices())


def trim_memory(device=0):
    """Returns the device blocks parked in the library's caching allocator to the driver (cudapath_trim_memory)."""
    _check(lib().cudapath_trim_memory(int(device)))


def scene_from_description(name, device=0, scale=1.0, overrides=None, data_dir=None):
    """Builds a Context for one of scenes.SCENES directly from flattened arrays (the path a librender plugin would take)."""
    from . import scenes
    sc = dict(scenes.SCENES[name]); sc.update(overrides or {})
    ctx = Context(device, data_dir)
    scenes.add_shapes(ctx, sc, scale)
    ctx.set_sunsky(**scenes.sunsky_params(name))
    ctx.set_camera(np.array(sc['camera'], np.float32).reshape(4, 4), sc['fov'], width=sc['width'], height=sc['height'])
    ctx.set_film('tent')
    ctx.set_integrator(maxDepth=sc['maxDepth'], rrDepth=5, strictNormals=True)
    ctx.spp = sc['spp']
    return ctx
